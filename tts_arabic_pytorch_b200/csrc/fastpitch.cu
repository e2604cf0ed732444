// FastPitch inference (models/fastpitch/fastpitch/model.py:351-409) as fused launches over a
// padded batch, channel-last fp16 activations.
//
// The reference is not batch-invariant (SURVEY.md §7 hard part 3): PositionwiseConvFF and
// TemporalPredictor stack convolutions without masking in between, so the first padded position
// leaks into the last valid one. This implementation keeps exactly the reference's masking points
// (mask after attention+LN, after conv-FF+LN, on predictor inputs/outputs) and computes ALL rows
// of the padded batch in between, which reproduces that behaviour by construction.
#include <cmath>
#include "model_common.cuh"

using namespace ttsb;

namespace {

struct FftLayer {
    ConvLayer qkv, o, ff1, ff2;
    float *ln1_g = nullptr, *ln1_b = nullptr, *ln2_g = nullptr, *ln2_b = nullptr;
};
struct Predictor {
    ConvLayer c0, c1;
    float *n0_g = nullptr, *n0_b = nullptr, *n1_g = nullptr, *n1_b = nullptr, *fc_w = nullptr;
    float fc_b = 0.f;
};
struct ScalarEmb {
    float *w = nullptr, *b = nullptr;  // [D,3], [D]
};

}  // namespace

struct ttsb_fastpitch {
    ttsb_fastpitch_config_t cfg;
    int device = 0;
    float* word_emb = nullptr;
    float* inv_freq_enc = nullptr;
    float* inv_freq_dec = nullptr;
    float* spk_table = nullptr;  // [n_speakers, D] pre-scaled by speaker_emb_weight
    std::vector<FftLayer> enc, dec;
    Predictor dur, pitch, energy;
    ScalarEmb pitch_emb, energy_emb;
    ConvLayer proj;
    int mel_ld = 128;
};

namespace {

int make_linear_layer(ConvLayer& L, const float* w, const float* bias, int n_out, int n_in, int n_out_pad,
                      int n_tile_hint) {
    // nn.Linear weight [out,in] == one-tap conv; rows [n_out, n_out_pad) are zero
    std::vector<float> wl(static_cast<size_t>(n_out_pad) * n_in, 0.f);
    std::vector<float> bl(n_out_pad, 0.f);
    for (int n = 0; n < n_out; ++n) {
        for (int k = 0; k < n_in; ++k) wl[static_cast<size_t>(n) * n_in + k] = w[static_cast<size_t>(n) * n_in + k];
        if (bias) bl[n] = bias[n];
    }
    const int off[1] = {0};
    return conv_layer_create(L, n_in, n_in, n_out_pad, 1, off, nullptr, 1 << 30, wl.data(),
                             bias ? bl.data() : nullptr, n_tile_hint);
}

int load_fft(const TensorTable& tab, const std::string& pre, const ttsb_fastpitch_config_t& cfg, FftLayer& l) {
    const int D = cfg.d_model, dh = cfg.d_head, di = cfg.d_inner, k = cfg.conv_kernel;
    TTSB_GET_TENSOR(qw, tab, pre + ".dec_attn.qkv_net.weight", 2);
    TTSB_GET_TENSOR(qb, tab, pre + ".dec_attn.qkv_net.bias", 1);
    TTSB_GET_TENSOR(ow, tab, pre + ".dec_attn.o_net.weight", 2);
    TTSB_GET_TENSOR(g1, tab, pre + ".dec_attn.layer_norm.weight", 1);
    TTSB_GET_TENSOR(b1, tab, pre + ".dec_attn.layer_norm.bias", 1);
    TTSB_GET_TENSOR(w1, tab, pre + ".pos_ff.CoreNet.0.weight", 3);
    TTSB_GET_TENSOR(c1, tab, pre + ".pos_ff.CoreNet.0.bias", 1);
    TTSB_GET_TENSOR(w2, tab, pre + ".pos_ff.CoreNet.2.weight", 3);
    TTSB_GET_TENSOR(c2, tab, pre + ".pos_ff.CoreNet.2.bias", 1);
    TTSB_GET_TENSOR(g2, tab, pre + ".pos_ff.layer_norm.weight", 1);
    TTSB_GET_TENSOR(b2, tab, pre + ".pos_ff.layer_norm.bias", 1);
    TTSB_REQUIRE(qw->shape[0] == 3 * dh && qw->shape[1] == D && ow->shape[0] == D && ow->shape[1] == dh,
                 pre + " attention shapes (one head of d_head)");
    TTSB_REQUIRE(w1->shape[0] == di && w1->shape[1] == D && w1->shape[2] == k && w2->shape[0] == D &&
                     w2->shape[1] == di && w2->shape[2] == k, pre + " conv-FF shapes");
    TTSB_PROPAGATE(make_linear_layer(l.qkv, qw->h_data, qb->h_data, 3 * dh, D, 3 * dh, 0));
    TTSB_PROPAGATE(make_linear_layer(l.o, ow->h_data, nullptr, D, dh, D, 0));
    TTSB_PROPAGATE(make_conv1d_layer(l.ff1, w1->h_data, c1->h_data, di, D, k, 1, D, 256));
    TTSB_PROPAGATE(make_conv1d_layer(l.ff2, w2->h_data, c2->h_data, D, di, k, 1, di, 0));
    TTSB_PROPAGATE(upload_f32(g1->h_data, D, &l.ln1_g));
    TTSB_PROPAGATE(upload_f32(b1->h_data, D, &l.ln1_b));
    TTSB_PROPAGATE(upload_f32(g2->h_data, D, &l.ln2_g));
    TTSB_PROPAGATE(upload_f32(b2->h_data, D, &l.ln2_b));
    return 0;
}

int load_predictor(const TensorTable& tab, const std::string& pre, const ttsb_fastpitch_config_t& cfg,
                   Predictor& p) {
    const int D = cfg.d_model, F = cfg.pred_filter, k = cfg.pred_kernel;
    TTSB_GET_TENSOR(w0, tab, pre + ".layers.0.conv.weight", 3);
    TTSB_GET_TENSOR(b0, tab, pre + ".layers.0.conv.bias", 1);
    TTSB_GET_TENSOR(g0, tab, pre + ".layers.0.norm.weight", 1);
    TTSB_GET_TENSOR(h0, tab, pre + ".layers.0.norm.bias", 1);
    TTSB_GET_TENSOR(w1, tab, pre + ".layers.1.conv.weight", 3);
    TTSB_GET_TENSOR(b1, tab, pre + ".layers.1.conv.bias", 1);
    TTSB_GET_TENSOR(g1, tab, pre + ".layers.1.norm.weight", 1);
    TTSB_GET_TENSOR(h1, tab, pre + ".layers.1.norm.bias", 1);
    TTSB_GET_TENSOR(fw, tab, pre + ".fc.weight", 2);
    TTSB_GET_TENSOR(fb, tab, pre + ".fc.bias", 1);
    TTSB_REQUIRE(w0->shape[0] == F && w0->shape[1] == D && w0->shape[2] == k && w1->shape[0] == F &&
                     w1->shape[1] == F && fw->shape[0] == 1 && fw->shape[1] == F,
                 pre + " shapes (2 ConvReLUNorm layers, 1 prediction)");
    TTSB_PROPAGATE(make_conv1d_layer(p.c0, w0->h_data, b0->h_data, F, D, k, 1, D, 0));
    TTSB_PROPAGATE(make_conv1d_layer(p.c1, w1->h_data, b1->h_data, F, F, k, 1, F, 0));
    TTSB_PROPAGATE(upload_f32(g0->h_data, F, &p.n0_g));
    TTSB_PROPAGATE(upload_f32(h0->h_data, F, &p.n0_b));
    TTSB_PROPAGATE(upload_f32(g1->h_data, F, &p.n1_g));
    TTSB_PROPAGATE(upload_f32(h1->h_data, F, &p.n1_b));
    TTSB_PROPAGATE(upload_f32(fw->h_data, F, &p.fc_w));
    p.fc_b = fb->h_data[0];
    return 0;
}

int load_scalar_emb(const TensorTable& tab, const std::string& pre, int D, ScalarEmb& e) {
    TTSB_GET_TENSOR(w, tab, pre + ".weight", 3);
    TTSB_GET_TENSOR(b, tab, pre + ".bias", 1);
    TTSB_REQUIRE(w->shape[0] == D && w->shape[1] == 1 && w->shape[2] == 3, pre + " must be Conv1d(1,D,3)");
    TTSB_PROPAGATE(upload_f32(w->h_data, static_cast<size_t>(D) * 3, &e.w));
    TTSB_PROPAGATE(upload_f32(b->h_data, D, &e.b));
    return 0;
}

void free_fft(FftLayer& l) {
    conv_layer_destroy(l.qkv); conv_layer_destroy(l.o); conv_layer_destroy(l.ff1); conv_layer_destroy(l.ff2);
    cudaFree(l.ln1_g); cudaFree(l.ln1_b); cudaFree(l.ln2_g); cudaFree(l.ln2_b);
}
void free_pred(Predictor& p) {
    conv_layer_destroy(p.c0); conv_layer_destroy(p.c1);
    cudaFree(p.n0_g); cudaFree(p.n0_b); cudaFree(p.n1_g); cudaFree(p.n1_b); cudaFree(p.fc_w);
}

struct State {  // persistent between encode / condition / decode
    int* lens;
    int* dec_lens;
    int* cum;
    int* summary;   // [0] max(dec_lens), [1] input status bits (launch_ids_to_lens)
    __half* x;  // [B,L,D]
};
struct Scratch {
    __half *qkv, *att, *x, *hid, *p1, *vt;
};

State carve_state(const ttsb_fastpitch* h, void* p, int B, int L, size_t* bytes) {
    Carver c(p);
    State s;
    s.lens = c.take<int>(B);
    s.dec_lens = c.take<int>(B);
    s.cum = c.take<int>(static_cast<size_t>(B) * (L + 1));
    s.summary = c.take<int>(4);
    s.x = c.take<__half>(static_cast<size_t>(B) * L * h->cfg.d_model);
    if (bytes) *bytes = c.off + 256;
    return s;
}
Scratch carve_scratch(const ttsb_fastpitch* h, void* p, int B, int R, size_t* bytes) {
    Carver c(p);
    Scratch s;
    const size_t rows = static_cast<size_t>(B) * R;
    s.qkv = c.take<__half>(rows * 3 * h->cfg.d_head);
    s.att = c.take<__half>(rows * h->cfg.d_head);
    s.x = c.take<__half>(rows * h->cfg.d_model);
    s.hid = c.take<__half>(rows * h->cfg.d_inner);
    s.p1 = c.take<__half>(rows * h->cfg.pred_filter);
    s.vt = c.take<__half>(attention_tc_scratch_bytes(B, R) / sizeof(__half));
    if (bytes) *bytes = c.off + 256;
    return s;
}

// One FFT block (transformer.py:172-177): attention + post-LN, mask, conv-FF + post-LN, mask.
int run_fft_layer(const ttsb_fastpitch* h, const FftLayer& l, const ConvRuntime& rt, __half* x, const int* lens,
                  int B, int R, const Scratch& sc, cudaStream_t stream) {
    const int D = h->cfg.d_model, dh = h->cfg.d_head, di = h->cfg.d_inner;
    {
        EpiParams e;
        e.out_raw = sc.qkv; e.ld_raw = 3 * dh;
        prof_mark(PROF_FP_QKV_O, stream);
        TTSB_PROPAGATE(conv_forward(l.qkv, rt, x, D, B, R, e, stream));
    }
    prof_mark(PROF_FP_ATTENTION, stream);
    TTSB_PROPAGATE(launch_attention(sc.qkv, lens, B, R, 1.f / sqrtf(static_cast<float>(dh)), sc.att, stream, sc.vt, rt.err_flag));
    {
        EpiParams e;
        e.residual = x; e.ld_res = D;
        e.ln_g = l.ln1_g; e.ln_b = l.ln1_b;
        e.lens = lens;
        e.out_raw = x; e.ld_raw = D;
        prof_mark(PROF_FP_QKV_O, stream);
        TTSB_PROPAGATE(conv_forward(l.o, rt, sc.att, dh, B, R, e, stream));
        prof_mark(PROF_FP_FFN, stream);
    }
    {
        EpiParams e;
        e.out_act = sc.hid; e.ld_act = di; e.act_slope = 0.f;  // ReLU, deliberately unmasked
        TTSB_PROPAGATE(conv_forward(l.ff1, rt, x, D, B, R, e, stream));
    }
    {
        EpiParams e;
        e.residual = x; e.ld_res = D;
        e.ln_g = l.ln2_g; e.ln_b = l.ln2_b;
        e.lens = lens;
        e.out_raw = x; e.ld_raw = D;
        TTSB_PROPAGATE(conv_forward(l.ff2, rt, sc.hid, di, B, R, e, stream));
    }
    return 0;
}

// TemporalPredictor (model.py:129-133): [conv k3 -> relu -> LN] x 2 -> fc -> * mask
int run_predictor(const ttsb_fastpitch* h, const Predictor& p, const ConvRuntime& rt, const __half* x,
                  const int* lens, int B, int L, const Scratch& sc, float* out, cudaStream_t stream) {
    const int D = h->cfg.d_model, F = h->cfg.pred_filter;
    prof_mark(PROF_FP_PREDICTORS, stream);
    {
        EpiParams e;
        e.pre_ln_relu = 1; e.ln_g = p.n0_g; e.ln_b = p.n0_b;
        e.out_raw = sc.p1; e.ld_raw = F;  // unmasked on purpose (no mask between the two layers)
        TTSB_PROPAGATE(conv_forward(p.c0, rt, x, D, B, L, e, stream));
    }
    {
        EpiParams e;
        e.pre_ln_relu = 1; e.ln_g = p.n1_g; e.ln_b = p.n1_b;
        e.head_w = p.fc_w; e.head_b = p.fc_b; e.head_out = out;
        e.lens = lens;
        TTSB_PROPAGATE(conv_forward(p.c1, rt, sc.p1, F, B, L, e, stream));
    }
    return 0;
}

size_t simt_elems(const ttsb_fastpitch* h, int B, int R) {
    return static_cast<size_t>(B) * R * std::max(h->cfg.d_inner, 512);
}

}  // namespace

extern "C" {

int ttsb_fastpitch_create(const ttsb_fastpitch_config_t* cfg, const ttsb_tensor_t* weights, int n_weights,
                          int device, ttsb_fastpitch_t** out) {
    return guarded_call([&]() -> int {
    TTSB_REQUIRE(cfg && weights && out, "null argument");
    TTSB_REQUIRE(cfg->d_head == 64, "attention kernel is specialised for one head of 64");
    TTSB_REQUIRE(cfg->d_model % 64 == 0 && cfg->d_inner % 256 == 0 && cfg->pred_filter % 64 == 0 &&
                     cfg->d_model <= 512 && cfg->pred_filter <= 512, "channel sizes");
    TTSB_REQUIRE(cfg->conv_kernel == 3 && cfg->pred_kernel == 3, "kernel size 3 expected");
    TTSB_DEVICE_GUARD(device);
    TensorTable tab(weights, n_weights);
    std::unique_ptr<ttsb_fastpitch, void (*)(ttsb_fastpitch*)> owner(new ttsb_fastpitch(), ttsb_fastpitch_destroy);
    ttsb_fastpitch* h = owner.get();
    h->cfg = *cfg;
    h->device = device;
    h->mel_ld = round_up(cfg->n_mel_channels, 64);
    const int D = cfg->d_model;
    {
        TTSB_GET_TENSOR(we, tab, "encoder.word_emb.weight", 2);
        TTSB_REQUIRE(we->shape[0] == cfg->n_symbols && we->shape[1] == D, "word_emb shape");
        TTSB_PROPAGATE(upload_f32(we->h_data, TensorTable::numel(we), &h->word_emb));
        TTSB_GET_TENSOR(fe, tab, "encoder.pos_emb.inv_freq", 1);
        TTSB_GET_TENSOR(fd, tab, "decoder.pos_emb.inv_freq", 1);
        TTSB_REQUIRE(fe->shape[0] == D / 2 && fd->shape[0] == D / 2, "inv_freq shape");
        TTSB_PROPAGATE(upload_f32(fe->h_data, D / 2, &h->inv_freq_enc));
        TTSB_PROPAGATE(upload_f32(fd->h_data, D / 2, &h->inv_freq_dec));
    }
    if (cfg->n_speakers > 1) {
        TTSB_GET_TENSOR(se, tab, "speaker_emb.weight", 2);
        TTSB_REQUIRE(se->shape[0] == cfg->n_speakers && se->shape[1] == D, "speaker_emb shape");
        std::vector<float> t(TensorTable::numel(se));
        for (size_t i = 0; i < t.size(); ++i) t[i] = se->h_data[i] * cfg->speaker_emb_weight;
        TTSB_PROPAGATE(upload_f32(t.data(), t.size(), &h->spk_table));
    }
    h->enc.resize(cfg->n_layers_enc);
    h->dec.resize(cfg->n_layers_dec);
    for (int i = 0; i < cfg->n_layers_enc; ++i)
        TTSB_PROPAGATE(load_fft(tab, "encoder.layers." + std::to_string(i), *cfg, h->enc[i]));
    for (int i = 0; i < cfg->n_layers_dec; ++i)
        TTSB_PROPAGATE(load_fft(tab, "decoder.layers." + std::to_string(i), *cfg, h->dec[i]));
    TTSB_PROPAGATE(load_predictor(tab, "duration_predictor", *cfg, h->dur));
    TTSB_PROPAGATE(load_predictor(tab, "pitch_predictor", *cfg, h->pitch));
    TTSB_PROPAGATE(load_scalar_emb(tab, "pitch_emb", D, h->pitch_emb));
    if (cfg->energy_conditioning) {
        TTSB_PROPAGATE(load_predictor(tab, "energy_predictor", *cfg, h->energy));
        TTSB_PROPAGATE(load_scalar_emb(tab, "energy_emb", D, h->energy_emb));
    }
    {
        TTSB_GET_TENSOR(pw, tab, "proj.weight", 2);
        TTSB_GET_TENSOR(pb, tab, "proj.bias", 1);
        TTSB_REQUIRE(pw->shape[0] == cfg->n_mel_channels && pw->shape[1] == D, "proj shape");
        TTSB_PROPAGATE(make_linear_layer(h->proj, pw->h_data, pb->h_data, cfg->n_mel_channels, D, h->mel_ld, 0));
    }
    *out = owner.release();
    return 0;
    });
}

void ttsb_fastpitch_destroy(ttsb_fastpitch_t* h) {
    if (!h) return;
    for (auto& l : h->enc) free_fft(l);
    for (auto& l : h->dec) free_fft(l);
    free_pred(h->dur); free_pred(h->pitch); free_pred(h->energy);
    cudaFree(h->pitch_emb.w); cudaFree(h->pitch_emb.b); cudaFree(h->energy_emb.w); cudaFree(h->energy_emb.b);
    conv_layer_destroy(h->proj);
    cudaFree(h->word_emb); cudaFree(h->inv_freq_enc); cudaFree(h->inv_freq_dec); cudaFree(h->spk_table);
    delete h;
}

size_t ttsb_fastpitch_state_bytes(const ttsb_fastpitch_t* h, int B, int L) {
    size_t n = 0;
    carve_state(h, nullptr, B, L, &n);
    return n;
}
size_t ttsb_fastpitch_workspace_bytes(const ttsb_fastpitch_t* h, int B, int L, int T) {
    size_t n = 0;
    carve_scratch(h, nullptr, B, std::max(L, T), &n);
    return n;
}

int ttsb_fastpitch_encode(ttsb_fastpitch_t* h, const int64_t* d_ids, int B, int L, int speaker,
                          const int64_t* d_speaker_ids, float* d_log_dur, float* d_pitch, void* d_state,
                          void* d_workspace, size_t workspace_bytes, void* stream_) {
    return guarded_call([&]() -> int {
    TTSB_REQUIRE(h && d_ids && d_log_dur && d_pitch && d_state && d_workspace, "null argument");
    TTSB_REQUIRE(B > 0 && L > 0, "empty batch");
    TTSB_REQUIRE(workspace_bytes >= ttsb_fastpitch_workspace_bytes(h, B, L, 0), "workspace too small");
    TTSB_REQUIRE(speaker < h->cfg.n_speakers, "speaker id out of range");
    TTSB_DEVICE_GUARD(h->device);
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    ConvRuntime rt;
    TTSB_PROPAGATE(get_conv_runtime(simt_elems(h, B, L), rt));
    State st = carve_state(h, d_state, B, L, nullptr);
    Scratch sc = carve_scratch(h, d_workspace, B, L, nullptr);
    const int D = h->cfg.d_model;
    const bool use_spk = h->spk_table != nullptr && (speaker >= 0 || d_speaker_ids != nullptr);
    prof_mark(PROF_FP_EMBED, stream);
    TTSB_CHECK_CUDA(cudaMemsetAsync(st.summary, 0, 4 * sizeof(int), stream));
    TTSB_PROPAGATE(launch_ids_to_lens(d_ids, B, L, h->cfg.n_symbols, use_spk ? d_speaker_ids : nullptr, h->cfg.n_speakers,
                                      st.lens, st.summary + 1, stream));
    TTSB_PROPAGATE(launch_embed(d_ids, h->word_emb, use_spk ? h->spk_table : nullptr, d_speaker_ids, speaker,
                                h->cfg.n_speakers, h->inv_freq_enc, B, L, D, h->cfg.n_symbols, st.x, stream));
    for (const FftLayer& l : h->enc) TTSB_PROPAGATE(run_fft_layer(h, l, rt, st.x, st.lens, B, L, sc, stream));
    TTSB_PROPAGATE(run_predictor(h, h->dur, rt, st.x, st.lens, B, L, sc, d_log_dur, stream));
    TTSB_PROPAGATE(run_predictor(h, h->pitch, rt, st.x, st.lens, B, L, sc, d_pitch, stream));
    prof_mark(PROF_NONE, stream);
    return 0;
    });
}

int ttsb_fastpitch_read_enc_out(ttsb_fastpitch_t* h, int B, int L, const void* d_state, void* d_enc_out, void* stream_) {
    return guarded_call([&]() -> int {
    TTSB_REQUIRE(h && d_state && d_enc_out, "null argument");
    TTSB_DEVICE_GUARD(h->device);
    State st = carve_state(h, const_cast<void*>(d_state), B, L, nullptr);
    TTSB_CHECK_CUDA(cudaMemcpyAsync(d_enc_out, st.x, static_cast<size_t>(B) * L * h->cfg.d_model * sizeof(__half),
                                    cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream_)));
    return 0;
    });
}

int ttsb_fastpitch_condition(ttsb_fastpitch_t* h, int B, int L, const float* d_log_dur,
                             const float* d_pitch_in, const float* d_energy_tgt, const float* d_dur_tgt,
                             float pace, float max_duration, float* d_dur_pred, float* d_energy_pred,
                             int64_t* d_dec_lens, int32_t* d_summary, void* d_state, void* d_workspace,
                             size_t workspace_bytes, void* stream_) {
    return guarded_call([&]() -> int {
    TTSB_REQUIRE(h && d_log_dur && d_pitch_in && d_dur_pred && d_dec_lens && d_state && d_workspace, "null argument");
    TTSB_REQUIRE(workspace_bytes >= ttsb_fastpitch_workspace_bytes(h, B, L, 0), "workspace too small");
    TTSB_REQUIRE(pace > 0.f, "pace must be positive");
    TTSB_DEVICE_GUARD(h->device);
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    ConvRuntime rt;
    TTSB_PROPAGATE(get_conv_runtime(simt_elems(h, B, L), rt));
    State st = carve_state(h, d_state, B, L, nullptr);
    Scratch sc = carve_scratch(h, d_workspace, B, L, nullptr);
    const int D = h->cfg.d_model;
    // enc_out = enc_out + pitch_emb(pitch)   (model.py:382-386); masked copy is what every later
    // consumer sees (energy predictor masks its input, padded tokens get zero frames)
    prof_mark(PROF_FP_GLUE, stream);
    TTSB_PROPAGATE(launch_scalar_embed_add(st.x, d_pitch_in, h->pitch_emb.w, h->pitch_emb.b, st.lens, B, L, D, 1, stream));
    if (h->cfg.energy_conditioning) {
        const float* energy = d_energy_tgt;
        if (!energy) {
            TTSB_REQUIRE(d_energy_pred != nullptr, "energy_pred output required");
            TTSB_PROPAGATE(run_predictor(h, h->energy, rt, st.x, st.lens, B, L, sc, d_energy_pred, stream));
            energy = d_energy_pred;
        }
        prof_mark(PROF_FP_GLUE, stream);
        TTSB_PROPAGATE(launch_scalar_embed_add(st.x, energy, h->energy_emb.w, h->energy_emb.b, st.lens, B, L, D, 1, stream));
    }
    TTSB_CHECK_CUDA(cudaMemsetAsync(st.summary, 0, sizeof(int), stream));     // [0] only: the input status of encode stays
    TTSB_PROPAGATE(launch_durations(d_log_dur, d_dur_tgt, pace, max_duration, B, L, d_dur_pred, st.cum,
                                    st.dec_lens, d_dec_lens, st.summary, stream));
    if (d_summary) {
        TTSB_CHECK_CUDA(cudaMemcpyAsync(d_summary, st.summary, 2 * sizeof(int), cudaMemcpyDeviceToDevice, stream));
        TTSB_CHECK_CUDA(cudaMemcpyAsync(d_summary + 2, st.dec_lens, B * sizeof(int), cudaMemcpyDeviceToDevice, stream));
    }
    prof_mark(PROF_NONE, stream);
    return 0;
    });
}

int ttsb_fastpitch_decode(ttsb_fastpitch_t* h, int B, int L, int T, float* d_mel, void* d_mel_cl,
                          void* d_state, void* d_workspace, size_t workspace_bytes, void* stream_) {
    return guarded_call([&]() -> int {
    TTSB_REQUIRE(h && d_mel && d_state && d_workspace, "null argument");
    TTSB_REQUIRE(B > 0 && L > 0 && T > 0, "empty batch");
    TTSB_REQUIRE(workspace_bytes >= ttsb_fastpitch_workspace_bytes(h, B, L, T), "workspace too small");
    TTSB_DEVICE_GUARD(h->device);
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    ConvRuntime rt;
    TTSB_PROPAGATE(get_conv_runtime(simt_elems(h, B, T), rt));
    State st = carve_state(h, d_state, B, L, nullptr);
    Scratch sc = carve_scratch(h, d_workspace, B, std::max(L, T), nullptr);
    const int D = h->cfg.d_model;
    prof_mark(PROF_FP_REGULATE, stream);
    TTSB_PROPAGATE(launch_regulate(st.x, st.cum, st.dec_lens, h->inv_freq_dec, B, L, T, D, sc.x, stream));
    for (const FftLayer& l : h->dec) TTSB_PROPAGATE(run_fft_layer(h, l, rt, sc.x, st.dec_lens, B, T, sc, stream));
    EpiParams e;
    e.out_f32_t = d_mel; e.n_store = h->cfg.n_mel_channels; e.f32_unmasked = 1;
    if (d_mel_cl) {
        e.lens = st.dec_lens;
        e.out_raw = static_cast<__half*>(d_mel_cl); e.ld_raw = h->mel_ld;
    }
    prof_mark(PROF_FP_PROJ, stream);
    TTSB_PROPAGATE(conv_forward(h->proj, rt, sc.x, D, B, T, e, stream));
    prof_mark(PROF_NONE, stream);
    return 0;
    });
}

}  // extern "C"
