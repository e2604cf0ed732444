// Host-side description of one "row GEMM with taps": the single compute primitive behind every
// dense contraction on the hot path.
//
//   out[b, t, n] = sum_{tap} sum_{ci} in[b, t + off(tap), ci] * W[n, tap, ci]      (+ fused epilogue)
//
// with channel-last fp16 activations in[B, T, ld] (rows outside [0, T) read as zero, i.e. the
// zero padding of nn.Conv1d). Instances:
//   nn.Conv1d (dilated)        hifigan/models.py:26-44,92,107  transformer.py:60-63  model.py:48-50
//   nn.ConvTranspose1d         hifigan/models.py:98-100  -> two-tap polyphase form, N = stride*Cout
//   nn.Linear                  transformer.py:105,108  model.py:127,231      -> one tap
#pragma once
#include <vector>
#include "common.cuh"
#include "epilogue.cuh"

namespace ttsb {

constexpr int kMaxTaps = 16;
constexpr int kTileM = 128;  // output rows (time positions) per CTA

struct ConvLayer {
    // logical shape
    int cin = 0;        // input channels as stored (multiple of chunk_k)
    int n_total = 0;    // output columns (Cout, or stride*Cout for transposed conv)
    int n_taps = 1;
    int tap_off[2][kMaxTaps] = {};  // row offsets; class 1 used by N tiles >= class_split
    int class_split = 1 << 30;
    // tiling
    int chunk_k = 64;   // K elements per smem row (64 -> 128B swizzle, 32 -> 64B swizzle)
    int n_chunks = 0;
    int n_tile = 0;     // columns per CTA
    int n_sub = 1;      // MMA N splits inside a CTA (n_tile / n_sub <= 256)
    int halo_lo = 0, halo_hi = 0;
    int rows_panel = 0;
    int a_slots = 0, b_stages = 0;
    int tmem_cols = 0;
    size_t smem_bytes = 0;
    // persistent kernel (conv_tc2) plan
    int a_slots2 = 0, b_stages2 = 0, acc_bufs = 1, occ2 = 1, tmem_cols2 = 0;
    int rpp = 1;        // row tiles per weight pass
    int resident = 0;   // weights stay in smem
    size_t smem_bytes2 = 0;
    // device data
    __half* w_packed = nullptr;  // [n_tiles][chunk][tap] tiles of n_tile x chunk_k, pre-swizzled
    // 32 -> 32 channel layers only: the same weights as ceil(n_taps / 2) tiles of 32 x 64 (128-byte rows, 128-byte
    // swizzle) holding taps (2g | 2g + 1) side by side along K, zeros beyond the last tap — conv_pair's conv2 when its
    // TT panel stores [t | t + 1] per row (two taps per K = 64 group on the faster 128-byte operand rows)
    __half* w_pair_packed = nullptr;
    float* bias = nullptr;       // [n_total] or null
    int n_tiles() const { return n_total / n_tile; }
};

// `w_logical` is a dense host array [n_total][n_taps][cin_logical] (fp32); channels
// [cin_logical, cin) are zero padding. `bias` may be null.
int conv_layer_create(ConvLayer& L, int cin_logical, int cin_stored, int n_total, int n_taps,
                      const int* tap_off0, const int* tap_off1, int class_split,
                      const float* w_logical, const float* bias, int n_tile_hint);
void conv_layer_destroy(ConvLayer& L);

enum ConvImpl : int { IMPL_TC = 0, IMPL_SIMT = 1 };

struct ConvRuntime {
    int impl = IMPL_TC;
    int tc_version = 2;         // 2 = persistent conv_tc2 kernel, 1 = one-tile-per-CTA conv_tc kernel
    int desc_mode = 0;          // A-descriptor base_offset rule for row-shifted tap views (see conv_tc.cu)
    int* err_flag = nullptr;    // device int, set by kernels on protocol timeouts
    float* simt_scratch = nullptr;
    size_t simt_scratch_elems = 0;
    long long* timeline = nullptr;  // debug timeline buffer (256 CTAs x 64 slots) or null
    int in_t_stride = 0;            // conv_tc2 only: rows between utterances of the INPUT tensor (0 = T), see get_act_tensor_map
};

// in: [B, T, ld_in] fp16 channel-last. epi.T / epi.n_total / epi.bias are filled from the layer.
int conv_forward(const ConvLayer& L, const ConvRuntime& rt, const __half* in, int ld_in, int B,
                 int T, EpiParams epi, cudaStream_t stream);

// shared by the tcgen05 kernels (defined in conv_tc2.cu)
int get_act_tensor_map(const __half* in, int ld_in, int B, int T, int cin, int chunk_k, int rows_panel,
                       const CUtensorMap** out, int t_stride = 0);
int num_sms();

// ------------------------------------------------------------------------------------------------
// Fused ResBlock1 step (conv_pair.cu):  out = x + conv2(lrelu(conv1(lrelu(x)) + b1)) + b2
// (vocoder/hifigan/models.py:46-53: one (c1, c2) iteration of ResBlock1.forward) in ONE launch;
// the intermediate never leaves the SM. L1 = dilated conv, L2 = dilation-1 conv, same C and k.
// ------------------------------------------------------------------------------------------------
struct ConvPairPlan {
    int ok = 0;
    int C = 0, m_out = 0, h2 = 0, dil = 1, rows_panel = 0, tt_rows = 0;
    int x_slots = 0, tt_slots = 0, w2_resident = 0, b_stages = 0, tmem_cols = 0;
    int tt_pair = 0;   // C = 32: TT rows hold [t | t + 1] (128 bytes), conv2 contracts two taps per K = 64 group (conv_pair.cu)
    size_t smem_bytes = 0;
    const __half* ident = nullptr;   // device: C x C identity in the packed weight-tile layout (shared per C, never freed)
};
// ok = 0 when the pair does not fit the fused kernel (caller falls back to two conv_forward launches)
ConvPairPlan conv_pair_plan(const ConvLayer& L1, const ConvLayer& L2);
// x: [B, T, C] input, also the residual: raw (in_act = 0: lrelu is applied in shared memory) or stored activated as
// lrelu(x, slope) (in_act = 1: no transform, the residual add inverts the activation). epi: bias/T/n_total/residual are
// filled here; the caller sets lens/len_mul, out_raw / out_act or mrf_*, act_slope.
int conv_pair_forward(const ConvLayer& L1, const ConvLayer& L2, const ConvPairPlan& plan, const ConvRuntime& rt,
                      const __half* x, int B, int T, float slope, EpiParams epi, cudaStream_t stream, int in_act = 0);

}  // namespace ttsb
