// HiFi-GAN V1 generator, batched and length-masked, as a sequence of fused conv launches.
// Mirrors Generator.forward / ResBlock1.forward (vocoder/hifigan/models.py:111-127, 46-53):
//
//   conv_pre -> 4 x [ lrelu(0.1) -> ConvTranspose1d -> mean_j ResBlock1_j ] -> lrelu(0.01) -> conv_post -> tanh
//
// Data flow per stage (all tensors channel-last fp16 [B, len, C]):
// Every ResBlock1 tensor x feeds a conv as lrelu(x) and, one conv later, a residual add as x. Both uses are served by ONE
// stored tensor, the activated one: the residual add inverts the (invertible) leaky-relu in the epilogue
// (EpiParams::res_inv), so no tensor is written twice and no launch needs an activation pass over its input
// (TTSB_ACT_CHAIN=0 restores the round-1 raw + activated pairs).
//   ups      : in  = act(prev)          -> LX0 = lrelu(X0)
//   resblock j, pair p (dilation d_p), LX_p = LX0 / LXA / LXB:
//     conv1  : in  = LX_p               -> TT = lrelu(conv + b)
//     conv2  : in  = TT, residual = inv_lrelu(LX_p)
//              p<2 : LX_{p+1} = lrelu(v)
//              p==2: MRF accumulate v/3 into XS; on the last resblock emit NXT = lrelu(XS_total)
// Pairs with a feasible ConvPairPlan (C <= 64) run conv1+conv2 as ONE launch (conv_pair.cu):
//     pair   : in = LX_p, out = LX_{p+1}  (p==2: MRF accumulate as above)
// Every epilogue zeroes rows beyond the utterance's own length so that a padded batch sees the
// same zero padding as the reference's per-utterance calls (models/fastpitch/networks.py:340-345).
#include <cstdlib>
#include "model_common.cuh"

using namespace ttsb;

struct ttsb_hifigan {
    ttsb_hifigan_config_t cfg;
    int device = 0;
    int hop = 1;
    int chunk_frames = 32768;   // frames per pass through the generator (workspace ~ 128 KB per frame)
    ConvLayer conv_pre;
    std::vector<ConvLayer> ups;
    std::vector<ConvLayer> c1, c2;  // index ((stage*num_kernels)+j)*3+p
    std::vector<ConvPairPlan> pair; // same index: fused (c1, c2) launch plan, ok = 0 -> two launches
    float* post_w = nullptr;        // [7][32] tap-major
    float post_b = 0.f;
    int post_k = 7;
    int max_elems_per_frame = 0;
    int mel_ld = 128;
};

namespace ttsb {

GlobalRuntime& global_runtime() {
    static GlobalRuntime g;
    static bool init = false;
    if (!init) {
        init = true;
        if (const char* e = getenv("TTSB_CONV_IMPL")) g.impl = (std::string(e) == "simt") ? IMPL_SIMT : IMPL_TC;
        if (const char* e = getenv("TTSB_DESC_MODE")) g.desc_mode = atoi(e);
        if (const char* e = getenv("TTSB_TC_VERSION")) g.tc_version = atoi(e) == 1 ? 1 : 2;
    }
    return g;
}

int get_conv_runtime(size_t simt_elems, ConvRuntime& rt) {
    GlobalRuntime& g = global_runtime();
    const int dev = current_device();
    if (!g.err_flag[dev]) {
        TTSB_CHECK_CUDA(cudaMalloc(&g.err_flag[dev], sizeof(int)));
        TTSB_CHECK_CUDA(cudaMemset(g.err_flag[dev], 0, sizeof(int)));
    }
    if (g.impl == IMPL_SIMT && g.simt_scratch_elems[dev] < simt_elems) {
        // check path only: grows outside any timed region
        TTSB_CHECK_CUDA(cudaDeviceSynchronize());
        if (g.simt_scratch[dev]) cudaFree(g.simt_scratch[dev]);
        g.simt_scratch[dev] = nullptr;
        g.simt_scratch_elems[dev] = 0;
        TTSB_CHECK_CUDA(cudaMalloc(&g.simt_scratch[dev], simt_elems * sizeof(float)));
        g.simt_scratch_elems[dev] = simt_elems;
    }
    rt.impl = g.impl;
    rt.desc_mode = g.desc_mode;
    rt.tc_version = g.desc_mode != 0 ? 1 : g.tc_version;   // descriptor probes exist only in v1
    rt.err_flag = g.err_flag[dev];
    rt.simt_scratch = g.simt_scratch[dev];
    rt.simt_scratch_elems = g.simt_scratch_elems[dev];
    rt.timeline = g.timeline;
    return 0;
}

int upload_f32(const float* h, size_t n, float** d) {
    TTSB_CHECK_CUDA(cudaMalloc(d, n * sizeof(float)));
    TTSB_CHECK_CUDA(cudaMemcpy(*d, h, n * sizeof(float), cudaMemcpyHostToDevice));
    return 0;
}

int make_conv1d_layer(ConvLayer& L, const float* w, const float* bias, int cout, int cin, int k,
                      int dilation, int cin_stored, int n_tile_hint) {
    TTSB_REQUIRE(k % 2 == 1, "odd kernel sizes only (same padding)");
    std::vector<float> wl(static_cast<size_t>(cout) * k * cin);
    for (int co = 0; co < cout; ++co)
        for (int ci = 0; ci < cin; ++ci)
            for (int kk = 0; kk < k; ++kk)
                wl[(static_cast<size_t>(co) * k + kk) * cin + ci] = w[(static_cast<size_t>(co) * cin + ci) * k + kk];
    int off[kMaxTaps];
    for (int kk = 0; kk < k; ++kk) off[kk] = (kk - (k - 1) / 2) * dilation;
    return conv_layer_create(L, cin, cin_stored, cout, k, off, nullptr, 1 << 30, wl.data(), bias, n_tile_hint);
}

int make_convT1d_layer(ConvLayer& L, const float* w, const float* bias, int cin, int cout, int k,
                       int stride) {
    TTSB_REQUIRE(k == 2 * stride && stride % 2 == 0, "transposed conv must have k = 2*stride, even stride");
    // y[m*s + p] = sum_i x[i] w[:, :, (m - i)*s + p + s/2]  (padding s/2):
    //   tap 0: i = m      -> kernel index p + s/2            (all phases)
    //   tap 1: i = m - 1  -> kernel index p + 3s/2           (phases p <  s/2, class 0)
    //          i = m + 1  -> kernel index p - s/2            (phases p >= s/2, class 1)
    const int n_total = stride * cout;
    std::vector<float> wl(static_cast<size_t>(n_total) * 2 * cin);
    std::vector<float> bl(n_total);
    for (int p = 0; p < stride; ++p)
        for (int co = 0; co < cout; ++co) {
            const int n = p * cout + co;
            bl[n] = bias ? bias[co] : 0.f;
            const int k0 = p + stride / 2;
            const int k1 = p < stride / 2 ? p + 3 * stride / 2 : p - stride / 2;
            for (int ci = 0; ci < cin; ++ci) {
                wl[(static_cast<size_t>(n) * 2 + 0) * cin + ci] = w[(static_cast<size_t>(ci) * cout + co) * k + k0];
                wl[(static_cast<size_t>(n) * 2 + 1) * cin + ci] = w[(static_cast<size_t>(ci) * cout + co) * k + k1];
            }
        }
    const int off0[2] = {0, -1}, off1[2] = {0, +1};
    const int half = n_total / 2;
    const int n_tile = half >= 256 ? 256 : half;
    TTSB_REQUIRE(half % n_tile == 0, "phase halves must tile evenly");
    return conv_layer_create(L, cin, cin, n_total, 2, off0, off1, half / n_tile, wl.data(),
                             bias ? bl.data() : nullptr, n_tile);
}

}  // namespace ttsb

extern "C" {

int ttsb_hifigan_create(const ttsb_hifigan_config_t* cfg, const ttsb_tensor_t* weights, int n_weights,
                        int device, ttsb_hifigan_t** out) {
    return guarded_call([&]() -> int {
    TTSB_REQUIRE(cfg && weights && out, "null argument");
    TTSB_DEVICE_GUARD(device);
    TTSB_REQUIRE(cfg->num_kernels >= 1 && cfg->num_kernels <= 8 && cfg->num_upsamples >= 1 && cfg->num_upsamples <= 8,
                 "config ranges");
    TensorTable tab(weights, n_weights);
    // owned until the end: an early return (bad checkpoint) must not leak the handle or its device buffers
    std::unique_ptr<ttsb_hifigan, void (*)(ttsb_hifigan*)> owner(new ttsb_hifigan(), ttsb_hifigan_destroy);
    ttsb_hifigan* h = owner.get();
    h->cfg = *cfg;
    h->device = device;
    if (const char* e = getenv("TTSB_HIFIGAN_CHUNK_FRAMES")) h->chunk_frames = atoi(e) > 0 ? atoi(e) : h->chunk_frames;
    const int C0 = cfg->upsample_initial_channel;
    h->mel_ld = round_up(cfg->num_mels, 64);
    {
        TTSB_GET_TENSOR(w, tab, "conv_pre.weight", 3);
        TTSB_GET_TENSOR(b, tab, "conv_pre.bias", 1);
        TTSB_REQUIRE(w->shape[0] == C0 && w->shape[1] == cfg->num_mels, "conv_pre shape");
        TTSB_PROPAGATE(make_conv1d_layer(h->conv_pre, w->h_data, b->h_data, C0, cfg->num_mels,
                                         static_cast<int>(w->shape[2]), 1, h->mel_ld, 256));
    }
    h->hop = 1;
    h->max_elems_per_frame = C0;
    int cin = C0;
    h->ups.resize(cfg->num_upsamples);
    h->c1.resize(cfg->num_upsamples * cfg->num_kernels * 3);
    h->c2.resize(cfg->num_upsamples * cfg->num_kernels * 3);
    h->pair.resize(cfg->num_upsamples * cfg->num_kernels * 3);
    for (int i = 0; i < cfg->num_upsamples; ++i) {
        const int cout = cin / 2, s = cfg->upsample_rates[i], k = cfg->upsample_kernel_sizes[i];
        const std::string pre = "ups." + std::to_string(i);
        TTSB_GET_TENSOR(w, tab, pre + ".weight", 3);
        TTSB_GET_TENSOR(b, tab, pre + ".bias", 1);
        TTSB_REQUIRE(w->shape[0] == cin && w->shape[1] == cout && w->shape[2] == k, pre + " shape");
        TTSB_PROPAGATE(make_convT1d_layer(h->ups[i], w->h_data, b->h_data, cin, cout, k, s));
        h->hop *= s;
        h->max_elems_per_frame = std::max(h->max_elems_per_frame, h->hop * cout);
        for (int j = 0; j < cfg->num_kernels; ++j) {
            const int rk = cfg->resblock_kernel_sizes[j];
            for (int p = 0; p < 3; ++p) {
                const std::string rb = "resblocks." + std::to_string(i * cfg->num_kernels + j);
                TTSB_GET_TENSOR(w1, tab, rb + ".convs1." + std::to_string(p) + ".weight", 3);
                TTSB_GET_TENSOR(b1, tab, rb + ".convs1." + std::to_string(p) + ".bias", 1);
                TTSB_GET_TENSOR(w2, tab, rb + ".convs2." + std::to_string(p) + ".weight", 3);
                TTSB_GET_TENSOR(b2, tab, rb + ".convs2." + std::to_string(p) + ".bias", 1);
                TTSB_REQUIRE(w1->shape[0] == cout && w1->shape[1] == cout && w1->shape[2] == rk, rb + " shape");
                const int idx = (i * cfg->num_kernels + j) * 3 + p;
                TTSB_PROPAGATE(make_conv1d_layer(h->c1[idx], w1->h_data, b1->h_data, cout, cout, rk,
                                                 cfg->resblock_dilations[j][p], cout, 0));
                TTSB_PROPAGATE(make_conv1d_layer(h->c2[idx], w2->h_data, b2->h_data, cout, cout, rk, 1, cout, 0));
                h->pair[idx] = conv_pair_plan(h->c1[idx], h->c2[idx]);
            }
        }
        cin = cout;
    }
    {
        TTSB_GET_TENSOR(w, tab, "conv_post.weight", 3);
        TTSB_GET_TENSOR(b, tab, "conv_post.bias", 1);
        TTSB_REQUIRE(w->shape[0] == 1 && w->shape[1] == cin && cin == 32 && w->shape[2] == 7,
                     "conv_post must be Conv1d(32,1,7)");
        std::vector<float> wt(7 * 32);
        for (int ci = 0; ci < 32; ++ci)
            for (int k = 0; k < 7; ++k) wt[k * 32 + ci] = w->h_data[ci * 7 + k];
        TTSB_PROPAGATE(upload_f32(wt.data(), wt.size(), &h->post_w));
        h->post_b = b->h_data[0];
    }
    *out = owner.release();
    return 0;
    });
}

void ttsb_hifigan_destroy(ttsb_hifigan_t* h) {
    if (!h) return;
    conv_layer_destroy(h->conv_pre);
    for (auto& l : h->ups) conv_layer_destroy(l);
    for (auto& l : h->c1) conv_layer_destroy(l);
    for (auto& l : h->c2) conv_layer_destroy(l);
    if (h->post_w) cudaFree(h->post_w);
    delete h;
}

int ttsb_hifigan_hop(const ttsb_hifigan_t* h) { return h ? h->hop : 0; }

static int chunk_batch(const ttsb_hifigan_t* h, int B, int T) {
    int bc = h->chunk_frames / (T > 0 ? T : 1);
    if (bc < 1) bc = 1;
    return bc > B ? B : bc;
}

size_t ttsb_hifigan_workspace_bytes(const ttsb_hifigan_t* h, int B, int T) {
    if (!h || B <= 0 || T <= 0) return 0;
    const int bc = chunk_batch(h, B, T);
    Carver c(nullptr);
    c.take<__half>(static_cast<size_t>(bc) * T * h->mel_ld);
    for (int i = 0; i < 8; ++i) c.take<__half>(static_cast<size_t>(bc) * T * h->max_elems_per_frame);
    return c.off + 256;
}

int ttsb_hifigan_forward(ttsb_hifigan_t* h, const float* d_mel_f32, const void* d_mel_cl,
                         const int32_t* d_lens, const int32_t* h_lens, int B, int T, float* d_wav, void* d_workspace,
                         size_t workspace_bytes, void* stream_) {
    return guarded_call([&]() -> int {
    TTSB_REQUIRE(h && d_wav && d_workspace, "null argument");
    TTSB_REQUIRE((d_mel_f32 != nullptr) != (d_mel_cl != nullptr), "exactly one mel input");
    TTSB_REQUIRE(B > 0 && T > 0, "empty batch");
    TTSB_REQUIRE(workspace_bytes >= ttsb_hifigan_workspace_bytes(h, B, T), "workspace too small");
    TTSB_DEVICE_GUARD(h->device);
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const ttsb_hifigan_config_t& cfg = h->cfg;
    const int bc_max = chunk_batch(h, B, T);
    ConvRuntime rt;
    TTSB_PROPAGATE(get_conv_runtime(static_cast<size_t>(bc_max) * T * h->max_elems_per_frame, rt));

    Carver c(d_workspace);
    __half* melp = c.take<__half>(static_cast<size_t>(bc_max) * T * h->mel_ld);
    __half* buf[8];
    for (int i = 0; i < 8; ++i) buf[i] = c.take<__half>(static_cast<size_t>(bc_max) * T * h->max_elems_per_frame);
    __half *NXT[2] = {buf[0], buf[1]}, *X0 = buf[2], *LX0 = buf[3], *TT = buf[4], *XA = buf[5], *LXA = buf[6],
           *XS = buf[7];
    __half* XB = TT;   // a resblock is either fused (X0 -> XA -> XB) or not (LX0/TT/XA/LXA); they run one after another
    __half* LXB = XA;  // activated chain: LX0 -> LXA -> LXB (XA and X0 are unused then)
    const bool use_pair = rt.impl == IMPL_TC && rt.tc_version == 2;
    static const bool act_chain = getenv("TTSB_ACT_CHAIN") ? atoi(getenv("TTSB_ACT_CHAIN")) != 0 : true;
    const float kSlope = 0.1f;   // LRELU_SLOPE (hifigan/models.py:9)

    // With a host copy of the lengths every chunk runs at ITS OWN longest utterance instead of the batch's: in a padded
    // mixed-length batch (BASELINE config 5: 64..256 phonemes) 37 % of the frames are padding, and every layer of the
    // generator would compute (and then zero) them. Chunks take as many consecutive utterances as fit the workspace.
    const bool ragged = h_lens != nullptr && d_lens != nullptr && rt.impl == IMPL_TC && rt.tc_version == 2;
    const long frame_budget = static_cast<long>(bc_max) * T;
    for (int b0 = 0; b0 < B;) {
        int bc = std::min(bc_max, B - b0);
        const int T_full = T;
        int Tc = T_full;
        if (ragged) {
            int longest = 1;
            bc = 0;
            while (b0 + bc < B) {
                const int l = std::max(longest, std::min(std::max(h_lens[b0 + bc], 1), T_full));
                if (bc > 0 && static_cast<long>(bc + 1) * l > frame_budget) break;
                longest = l;
                ++bc;
            }
            Tc = longest;
        }
        const int T = Tc;                     // rows per utterance of every buffer of this chunk
        const int* lens = d_lens ? d_lens + b0 : nullptr;
        const __half* mel_in;
        rt.in_t_stride = 0;
        if (d_mel_f32) {
            prof_mark(PROF_VOC_PACK, stream);
            TTSB_PROPAGATE(launch_pack_mel(d_mel_f32 + static_cast<size_t>(b0) * cfg.num_mels * T_full, lens, bc,
                                           cfg.num_mels, T, melp, h->mel_ld, stream, T_full));
            mel_in = melp;
        } else {
            mel_in = static_cast<const __half*>(d_mel_cl) + static_cast<size_t>(b0) * T_full * h->mel_ld;
            rt.in_t_stride = T_full;          // conv_pre reads the caller's [B, T_full, ld] tensor, T rows per utterance
        }
        int cur = 0;
        {
            EpiParams e;
            e.lens = lens; e.len_mul = 1;
            e.out_act = NXT[cur]; e.ld_act = cfg.upsample_initial_channel; e.act_slope = 0.1f;
            prof_mark(PROF_VOC_PRE, stream);
            TTSB_PROPAGATE(conv_forward(h->conv_pre, rt, mel_in, h->mel_ld, bc, T, e, stream));
            rt.in_t_stride = 0;
        }
        int cin = cfg.upsample_initial_channel;
        int up = 1;
        for (int i = 0; i < cfg.num_upsamples; ++i) {
            const int s = cfg.upsample_rates[i], C = cin / 2;
            bool any_unfused = false;
            for (int j = 0; j < cfg.num_kernels * 3; ++j) any_unfused |= !(use_pair && h->pair[i * cfg.num_kernels * 3 + j].ok);
            {
                EpiParams e;
                e.lens = lens; e.len_mul = up;
                if (act_chain) {
                    e.out_act = LX0; e.ld_act = s * C;
                } else {
                    e.out_raw = X0; e.ld_raw = s * C;
                    if (any_unfused) { e.out_act = LX0; e.ld_act = s * C; }
                }
                e.act_slope = kSlope;
                prof_mark(i < 4 ? PROF_VOC_UPS0 + 2 * i : PROF_NONE, stream);
                TTSB_PROPAGATE(conv_forward(h->ups[i], rt, NXT[cur], cin, bc, T * up, e, stream));
                prof_mark(i < 4 ? PROF_VOC_S0 + 2 * i : PROF_NONE, stream);
            }
            up *= s;
            const int rows = T * up;
            const bool last_stage = i == cfg.num_upsamples - 1;
            for (int j = 0; j < cfg.num_kernels; ++j) {
                for (int p = 0; p < 3; ++p) {
                    const int idx = (i * cfg.num_kernels + j) * 3 + p;
                    // a resblock is fused as a whole or not at all (its three pairs share k and C)
                    const bool fused = use_pair && h->pair[idx - p].ok && h->pair[idx - p + 1].ok && h->pair[idx - p + 2].ok;
                    __half* lx_in = p == 0 ? LX0 : (p == 1 ? LXA : LXB);      // activated chain
                    if (!fused) {
                        EpiParams e;
                        e.lens = lens; e.len_mul = up;
                        e.out_act = TT; e.ld_act = C; e.act_slope = kSlope;
                        TTSB_PROPAGATE(conv_forward(h->c1[idx], rt, act_chain ? lx_in : (p == 0 ? LX0 : LXA), C, bc, rows, e, stream));
                    }
                    EpiParams e;
                    e.lens = lens; e.len_mul = up;
                    if (act_chain) {
                        e.residual = lx_in; e.ld_res = C; e.res_inv = 1.f / kSlope;
                    } else {
                        e.residual = p == 0 ? X0 : XA; e.ld_res = C;
                    }
                    if (p < 2) {
                        if (act_chain) {
                            e.out_act = p == 0 ? LXA : LXB; e.ld_act = C;
                        } else {
                            e.out_raw = fused && p == 1 ? XB : XA; e.ld_raw = C;
                            if (!fused) { e.out_act = LXA; e.ld_act = C; }
                        }
                        e.act_slope = kSlope;
                    } else {
                        e.mrf_buf = XS;
                        e.mrf_scale = 1.f / static_cast<float>(cfg.num_kernels);
                        if (cfg.num_kernels == 1) {
                            e.mrf_mode = MRF_NONE;
                        } else {
                            e.mrf_mode = j == 0 ? MRF_FIRST : (j == cfg.num_kernels - 1 ? MRF_LAST : MRF_ADD);
                        }
                        if (j == cfg.num_kernels - 1) {
                            // F.leaky_relu(x) after the last stage uses the default slope 0.01
                            // (hifigan/models.py:123), the stage-to-stage one uses 0.1 (:114).
                            e.out_act = NXT[cur ^ 1]; e.ld_act = C;
                            e.act_slope = last_stage ? 0.01f : 0.1f;
                        }
                    }
                    if (fused) {
                        const __half* xin = act_chain ? lx_in : (p == 0 ? X0 : (p == 1 ? XA : XB));
                        TTSB_PROPAGATE(conv_pair_forward(h->c1[idx], h->c2[idx], h->pair[idx], rt, xin, bc, rows, kSlope, e, stream,
                                                         act_chain ? 1 : 0));
                    } else {
                        TTSB_PROPAGATE(conv_forward(h->c2[idx], rt, TT, C, bc, rows, e, stream));
                    }
                }
            }
            cur ^= 1;
            cin = C;
        }
        prof_mark(PROF_VOC_POST, stream);
        float* wav0 = d_wav + static_cast<size_t>(b0) * T_full * h->hop;
        TTSB_PROPAGATE(launch_conv_post_tanh(NXT[cur], h->post_w, h->post_b, lens, h->hop, bc, T * h->hop, wav0, stream,
                                             T_full * h->hop));
        if (T < T_full)    // samples beyond the chunk's longest utterance are part of the padded output: zeros
            TTSB_CHECK_CUDA(cudaMemset2DAsync(wav0 + static_cast<size_t>(T) * h->hop, static_cast<size_t>(T_full) * h->hop * sizeof(float),
                                              0, static_cast<size_t>(T_full - T) * h->hop * sizeof(float), bc, stream));
        prof_mark(PROF_NONE, stream);
        b0 += bc;
    }
    return 0;
    });
}

}  // extern "C"
