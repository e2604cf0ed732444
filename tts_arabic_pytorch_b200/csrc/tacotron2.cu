// Tacotron2 (multi-speaker) inference: Tacotron2MS.infer (models/tacotron2/tacotron2_ms.py:278-332) and the
// torchaudio 2.11 blocks it is built from (torchaudio/models/tacotron2.py: _Encoder.forward :396-418,
// _Prenet :273-285, _Attention :203-255, _LocationLayer :150-168, _Decoder.decode :611-684,
// _Decoder.infer :779-866, _Postnet.forward :330-346).
//
//   encoder   embedding -> 3 x [Conv1d k5 + BatchNorm(eval, folded) + ReLU] (tcgen05 conv kernel)
//             -> BiLSTM over each utterance's own length: input projections as one GEMM, the
//             recurrence as a per-(utterance, direction) sequential kernel
//   decoder   autoregressive; per step: prenet (injected / generated dropout masks — the reference
//             keeps p=0.5 dropout ON at inference), attention LSTM cell, location-sensitive attention,
//             decoder LSTM cell, mel + gate projection, stop bookkeeping. fp32 state, fp16 weights,
//             fp32 accumulation. The per-step host sync of the reference (torch.all(finished),
//             torchaudio:850) is replaced by a device flag the caller polls every N steps.
//   postnet   5 x [Conv1d k5 + BatchNorm(folded) (+tanh)] (tcgen05 conv kernel) + residual
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include "model_common.cuh"

using namespace ttsb;

namespace {

// ------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// State vectors that OTHER CTAs wrote earlier in the same launch (the persistent decoder below) must not come from this
// SM's L1, which is not coherent: they are read with ld.global.cg (L2 only). Weights stay on the cached read-only path.
__device__ __forceinline__ float4 ld_state4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }

// dot(w_row[0:n] (fp16), x[0:n] (fp32)) by one warp; n % 8 == 0. kState: x is such a state vector in global memory.
template <bool kState = false>
__device__ __forceinline__ float warp_dot_h(const __half* __restrict__ w, const float* __restrict__ x, int n, int lane) {
    // Four weight loads are issued before the first multiply-add of a pass: the weights come from L2 (~700 cycles), and a
    // one-load-per-iteration loop made every row cost (iterations x L2 latency) — the decoder's per-step time was that
    // latency chain, not bandwidth or arithmetic (tools/t2_phases.py, round 2). Same per-lane summation order.
    float acc = 0.f;
    for (int i0 = lane * 8; i0 < n; i0 += 4 * 256) {
        uint4 wv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int i = i0 + j * 256;
            wv[j] = i < n ? __ldg(reinterpret_cast<const uint4*>(w + i)) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int i = i0 + j * 256;
            if (i < n) {
                float f[8];
                unpack8(wv[j], f);
                const float4 a = kState ? ld_state4(x + i) : *reinterpret_cast<const float4*>(x + i);
                const float4 b = kState ? ld_state4(x + i + 4) : *reinterpret_cast<const float4*>(x + i + 4);
                acc += f[0] * a.x + f[1] * a.y + f[2] * a.z + f[3] * a.w + f[4] * b.x + f[5] * b.y + f[6] * b.z + f[7] * b.w;
            }
        }
    }
    return warp_sum(acc);
}

// ------------------------------------------------------------------------------------------------
// encoder pieces
// ------------------------------------------------------------------------------------------------
// token / speaker ids are clamped into their tables; out-of-range values raise bits in *status (1 token, 4 speaker), which
// the caller reads with the decoder's first host sync (nn.Embedding raises IndexError in the reference)
__global__ void t2_embed_kernel(const int64_t* __restrict__ ids, const float* __restrict__ emb, int D, int n_symbol,
                                int* __restrict__ status, __half* __restrict__ out) {
    const int row = blockIdx.x;
    int64_t id = ids[row];
    if (id < 0 || id >= n_symbol) {
        if (threadIdx.x == 0) atomicOr(status, 1);
        id = id < 0 ? 0 : n_symbol - 1;
    }
    for (int j = threadIdx.x; j < D; j += blockDim.x) out[static_cast<size_t>(row) * D + j] = __float2half(emb[id * D + j]);
}

// One CTA per (utterance, direction): h/c in smem, W_hh rows streamed from L2 each step.
// xproj: [B, L, 2*4H] fp32 (dir-major halves, torch gate order i,f,g,o), includes b_ih + b_hh.
// out: [B, L, 2H] fp32 (forward | backward), zero beyond each utterance's length.
__global__ void __launch_bounds__(1024) t2_bilstm_kernel(const float* __restrict__ xproj, const __half* __restrict__ whh,
                                                         const int* __restrict__ lens, int L, int H,
                                                         float* __restrict__ out) {
    extern __shared__ float sm[];
    float* h = sm;            // [H]
    float* c = sm + H;        // [H]
    float* gates = sm + 2 * H;  // [4H]
    const int b = blockIdx.x, dir = blockIdx.y;
    const int len = lens[b];
    const int G = 4 * H;
    const __half* w = whh + static_cast<size_t>(dir) * G * H;
    for (int j = threadIdx.x; j < H; j += blockDim.x) { h[j] = 0.f; c[j] = 0.f; }
    for (int t = threadIdx.x; t < L; t += blockDim.x)
        if (t >= len)
            for (int j = 0; j < H; ++j) out[(static_cast<size_t>(b) * L + t) * 2 * H + dir * H + j] = 0.f;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    for (int s = 0; s < len; ++s) {
        const int t = dir == 0 ? s : len - 1 - s;
        const float* xp = xproj + (static_cast<size_t>(b) * L + t) * 2 * G + dir * G;
        for (int g = warp; g < G; g += nwarp) {
            const float d = warp_dot_h(w + static_cast<size_t>(g) * H, h, H, lane);
            if (lane == 0) gates[g] = d + xp[g];
        }
        __syncthreads();
        for (int j = threadIdx.x; j < H; j += blockDim.x) {
            const float ig = sigmoidf_(gates[j]), fg = sigmoidf_(gates[H + j]);
            const float gg = tanhf(gates[2 * H + j]), og = sigmoidf_(gates[3 * H + j]);
            const float cn = fg * c[j] + ig * gg;
            const float hn = og * tanhf(cn);
            c[j] = cn;
            h[j] = hn;
            out[(static_cast<size_t>(b) * L + t) * 2 * H + dir * H + j] = hn;
        }
        __syncthreads();
    }
}

// memory = [enc_out (fp32, E) | speaker embedding (S)] and processed_memory = memory @ Wm^T
__global__ void t2_memory_kernel(const float* __restrict__ enc, const float* __restrict__ spk_emb,
                                 const int64_t* __restrict__ spk_ids, const __half* __restrict__ wm, int L, int E,
                                 int S, int A, int num_speakers, int* __restrict__ status, float* __restrict__ memory,
                                 float* __restrict__ pmem) {
    extern __shared__ float row[];  // [E+S]
    const int r = blockIdx.x;       // b*L + l
    const int b = r / L;
    const int M = E + S;
    int64_t sp = S > 0 ? spk_ids[b] : 0;
    if (S > 0 && (sp < 0 || sp >= num_speakers)) {
        if (threadIdx.x == 0) atomicOr(status, 4);
        sp = sp < 0 ? 0 : num_speakers - 1;
    }
    for (int j = threadIdx.x; j < M; j += blockDim.x) {
        const float v = j < E ? enc[static_cast<size_t>(r) * E + j] : spk_emb[sp * S + (j - E)];
        row[j] = v;
        memory[static_cast<size_t>(r) * M + j] = v;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    for (int a = warp; a < A; a += nwarp) {
        const float d = warp_dot_h(wm + static_cast<size_t>(a) * M, row, M, lane);
        if (lane == 0) pmem[static_cast<size_t>(r) * A + a] = d;
    }
}

// ------------------------------------------------------------------------------------------------
// decoder step kernels
// ------------------------------------------------------------------------------------------------
// prenet: frame [B,80] -> relu(W0 .) * m0 -> relu(W1 .) * m1 -> x [B,P]; masks are 0/1 bytes, scale 2
__device__ __forceinline__ void t2_prenet_body(float* sm, int b, const float* __restrict__ frame,
                                               const __half* __restrict__ w0, const __half* __restrict__ w1,
                                               const uint8_t* __restrict__ m0, const uint8_t* __restrict__ m1, int n_mel,
                                               int P, float* __restrict__ x) {
    float* f = sm;            // [n_mel]
    float* h = sm + n_mel;    // [P]
    for (int j = threadIdx.x; j < n_mel; j += blockDim.x) f[j] = frame[b * n_mel + j];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    for (int o = warp; o < P; o += nwarp) {
        const float d = warp_dot_h(w0 + static_cast<size_t>(o) * n_mel, f, n_mel, lane);
        if (lane == 0) h[o] = fmaxf(d, 0.f) * (m0[b * P + o] ? 2.f : 0.f);
    }
    __syncthreads();
    for (int o = warp; o < P; o += nwarp) {
        const float d = warp_dot_h(w1 + static_cast<size_t>(o) * P, h, P, lane);
        if (lane == 0) x[b * P + o] = fmaxf(d, 0.f) * (m1[b * P + o] ? 2.f : 0.f);
    }
}
__global__ void t2_prenet_kernel(const float* __restrict__ frame, const __half* __restrict__ w0,
                                 const __half* __restrict__ w1, const uint8_t* __restrict__ m0,
                                 const uint8_t* __restrict__ m1, int n_mel, int P, float* __restrict__ x) {
    extern __shared__ float sm[];
    t2_prenet_body(sm, blockIdx.x, frame, w0, w1, m0, m1, n_mel, P, x);
}

// LSTMCell: gates = W_ih [x1|x2] + W_hh h + bias; block = 8 warps = 2 hidden units x 4 gates, all B
// utterances (B <= 64 via a loop over groups of 8). h_in / h_out ping-pong, c in place.
__device__ __forceinline__ void t2_lstm_cell_body(float (*g_s)[4][64], int vb, const float* __restrict__ x1, int n1,
                                                  const float* __restrict__ x2, int n2, const float* __restrict__ h_in,
                                                  float* __restrict__ h_out, float* __restrict__ c,
                                                  const __half* __restrict__ w_ih, const __half* __restrict__ w_hh,
                                                  const float* __restrict__ bias, int H, int B) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ju = warp >> 2, gate = warp & 3;
    const int j = vb * 2 + ju;
    const int grow = gate * H + j;
    const int n_in = n1 + n2;
    const __half* wi = w_ih + static_cast<size_t>(grow) * n_in;
    const __half* wh = w_hh + static_cast<size_t>(grow) * H;
    for (int b0 = 0; b0 < B; b0 += 8) {
        float acc[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[q] = 0.f;
        for (int i = lane * 8; i < n_in + H; i += 256) {
            float f[8];
            const float* src;
            int stride, off;
            if (i < n1) { load8h(wi + i, f); src = x1; stride = n1; off = i; }
            else if (i < n_in) { load8h(wi + i, f); src = x2; stride = n2; off = i - n1; }
            else { load8h(wh + (i - n_in), f); src = h_in; stride = H; off = i - n_in; }
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                if (b0 + q < B) {
                    const float4 a = ld_state4(src + static_cast<size_t>(b0 + q) * stride + off);
                    const float4 bb = ld_state4(src + static_cast<size_t>(b0 + q) * stride + off + 4);
                    acc[q] += f[0] * a.x + f[1] * a.y + f[2] * a.z + f[3] * a.w + f[4] * bb.x + f[5] * bb.y + f[6] * bb.z + f[7] * bb.w;
                }
            }
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const float v = warp_sum(acc[q]);
            if (lane == 0 && b0 + q < B) g_s[ju][gate][b0 + q] = v + bias[grow];
        }
    }
    __syncthreads();
    if (threadIdx.x < 2 * B && threadIdx.x < 128) {
        const int u = threadIdx.x / B, b = threadIdx.x % B;
        if (u < 2) {
            const int jj = vb * 2 + u;
            const float ig = sigmoidf_(g_s[u][0][b]), fg = sigmoidf_(g_s[u][1][b]);
            const float gg = tanhf(g_s[u][2][b]), og = sigmoidf_(g_s[u][3][b]);
            const float cn = fg * c[static_cast<size_t>(b) * H + jj] + ig * gg;
            c[static_cast<size_t>(b) * H + jj] = cn;
            h_out[static_cast<size_t>(b) * H + jj] = og * tanhf(cn);
        }
    }
}
__global__ void __launch_bounds__(256) t2_lstm_cell_kernel(const float* __restrict__ x1, int n1,
                                                           const float* __restrict__ x2, int n2,
                                                           const float* __restrict__ h_in, float* __restrict__ h_out,
                                                           float* __restrict__ c, const __half* __restrict__ w_ih,
                                                           const __half* __restrict__ w_hh, const float* __restrict__ bias,
                                                           int H, int B) {
    __shared__ float g_s[2][4][64];
    t2_lstm_cell_body(g_s, blockIdx.x, x1, n1, x2, n2, h_in, h_out, c, w_ih, w_hh, bias, H, B);
}

// location-sensitive attention for one utterance per block (torchaudio:203-255)
__device__ __forceinline__ void t2_attention_body(float* sm, int b, const float* __restrict__ ah,
                                                  const float* __restrict__ memory, const float* __restrict__ pmem,
                                                  const int* __restrict__ lens, const __half* __restrict__ wq,
                                                  const float* wloc_conv, const float* wloc_dense,
                                                  const float* v, float* __restrict__ aw, float* __restrict__ awc,
                                                  float* __restrict__ ctx, float* __restrict__ align_out, int L, int H, int M,
                                                  int A, int NF, int KL, const float* __restrict__ q_in = nullptr,
                                                  const float* att_tables = nullptr) {
    // q_in (persistent decoder): the query projection W_q . ah[b] was computed by ALL CTAs in the phase before (one dot
    // product per warp instead of A / nwarp sequential L2 round trips here); same warp_dot_h, same values
    float* q = sm;                 // [A]
    float* w_prev = q + A;         // [L + KL - 1] padded previous weights
    float* w_cum = w_prev + L + KL; // [L + KL - 1]
    float* e = w_cum + L + KL;     // [L]
    float* red = e + L;            // [32]
    const int len = lens[b];
    const int pad = (KL - 1) / 2;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    // processed memory of THIS utterance [L, A]; att_tables (persistent decoder): shared-memory copies of
    // [location conv | location dense | v | this utterance's processed memory], staged once per launch — read through
    // global memory these small tables were evicted from L1 by every LSTM phase, and the energy loop is a chain of
    // dependent loads (31 taps, then 32 filters per output): ~50 k cycles per step of L2 latency (tools/t2_phases.py)
    const float* pmem_b = pmem + static_cast<size_t>(b) * L * A;
    if (att_tables != nullptr) {
        wloc_conv = att_tables;
        wloc_dense = att_tables + 2 * KL * NF;
        v = wloc_dense + NF * A;
        pmem_b = v + A;
    }
    // the query rows all multiply the same attention-LSTM output: stage it once (it was written by other CTAs: ld.cg)
    float* ahs = sm + ((A + 2 * (L + KL) + L + 32 + 3) & ~3);   // [H], 16-byte aligned for the float4 accesses
    if (q_in != nullptr) {
        for (int a = threadIdx.x; a < A; a += blockDim.x) q[a] = __ldcg(q_in + static_cast<size_t>(b) * A + a);
    } else {
        for (int k = threadIdx.x * 4; k < H; k += blockDim.x * 4)
            *reinterpret_cast<float4*>(ahs + k) = ld_state4(ah + static_cast<size_t>(b) * H + k);
        __syncthreads();
        for (int a = warp; a < A; a += nwarp) {
            const float d = warp_dot_h(wq + static_cast<size_t>(a) * H, ahs, H, lane);
            if (lane == 0) q[a] = d;
        }
    }
    for (int i = threadIdx.x; i < L + KL - 1; i += blockDim.x) {
        const int l = i - pad;
        w_prev[i] = (l >= 0 && l < L) ? aw[b * L + l] : 0.f;
        w_cum[i] = (l >= 0 && l < L) ? awc[b * L + l] : 0.f;
    }
    __syncthreads();
    // energies. Persistent decoder with the tables in shared memory (NF = 32 filters, A = 128): a warp works on FOUR
    // positions at once — the 31-tap location conv and the 32-filter dense sum are chains of dependent multiply-adds (one
    // shared-memory load + one shuffle per link), and one position at a time left the warp waiting on every link (~10 k
    // cycles per position: tools/t2_phases.py). Every (position, a) chain keeps the order of the loop below: same values.
    if (att_tables != nullptr && NF == 32 && A == 128) {
        for (int l0 = warp; l0 < L; l0 += 4 * nwarp) {
            int lp[4];
            float f[4], sacc[4][4];
#pragma unroll
            for (int p = 0; p < 4; ++p) { lp[p] = min(l0 + p * nwarp, L - 1); f[p] = 0.f; }
            for (int k = 0; k < KL; ++k) {
                const float c0 = wloc_conv[k * NF + lane], c1 = wloc_conv[(KL + k) * NF + lane];
#pragma unroll
                for (int p = 0; p < 4; ++p) f[p] += c0 * w_prev[lp[p] + k] + c1 * w_cum[lp[p] + k];
            }
#pragma unroll
            for (int p = 0; p < 4; ++p)
#pragma unroll
                for (int ai = 0; ai < 4; ++ai) sacc[p][ai] = q[lane + 32 * ai] + pmem_b[static_cast<size_t>(lp[p]) * A + lane + 32 * ai];
            for (int nf = 0; nf < 32; ++nf) {
                float wd[4], fv[4];
#pragma unroll
                for (int ai = 0; ai < 4; ++ai) wd[ai] = wloc_dense[nf * A + lane + 32 * ai];
#pragma unroll
                for (int p = 0; p < 4; ++p) fv[p] = __shfl_sync(0xffffffffu, f[p], nf);
#pragma unroll
                for (int p = 0; p < 4; ++p)
#pragma unroll
                    for (int ai = 0; ai < 4; ++ai) sacc[p][ai] += wd[ai] * fv[p];
            }
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                const int l = l0 + p * nwarp;
                float part = 0.f;
#pragma unroll
                for (int ai = 0; ai < 4; ++ai) part += v[lane + 32 * ai] * tanhf(sacc[p][ai]);
                part = warp_sum(part);
                if (lane == 0 && l < L) e[l] = l < len ? part : -INFINITY;
            }
        }
    } else
    for (int l = warp; l < L; l += nwarp) {
        float part = 0.f;
        if (l < len) {
            // location features f[nf] = sum_k conv[nf,0,k]*prev[l+k-pad] + conv[nf,1,k]*cum[l+k-pad]; lane = filter
            float f = 0.f;
            if (lane < NF) {
                // wloc_conv is stored TRANSPOSED ([2][KL][NF]): the lanes (filters) of a warp read one line per tap
                for (int k = 0; k < KL; ++k)
                    f += wloc_conv[k * NF + lane] * w_prev[l + k] + wloc_conv[(KL + k) * NF + lane] * w_cum[l + k];
            }
            for (int a = lane; a < A; a += 32) {
                float s = q[a] + pmem_b[static_cast<size_t>(l) * A + a];
                // wloc_dense is stored TRANSPOSED ([NF][A]): the lanes of a warp (consecutive a) read one line per filter
                for (int nf = 0; nf < NF; ++nf) s += wloc_dense[nf * A + a] * __shfl_sync(0xffffffffu, f, nf);
                part += v[a] * tanhf(s);
            }
        }
        part = warp_sum(part);
        if (lane == 0) e[l] = l < len ? part : -INFINITY;
    }
    __syncthreads();
    // softmax over l < len
    float mx = -INFINITY;
    for (int l = threadIdx.x; l < L; l += blockDim.x) mx = fmaxf(mx, e[l]);
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    mx = red[0];
    for (int i = 1; i < nwarp; ++i) mx = fmaxf(mx, red[i]);
    __syncthreads();
    float sum = 0.f;
    for (int l = threadIdx.x; l < L; l += blockDim.x) {
        const float p = l < len ? __expf(e[l] - mx) : 0.f;
        e[l] = p;
        sum += p;
    }
    sum = warp_sum(sum);
    if (lane == 0) red[warp] = sum;
    __syncthreads();
    sum = 0.f;
    for (int i = 0; i < nwarp; ++i) sum += red[i];
    const float inv = 1.f / sum;
    for (int l = threadIdx.x; l < L; l += blockDim.x) {
        const float p = e[l] * inv;
        e[l] = p;
        aw[b * L + l] = p;
        awc[b * L + l] += p;
        align_out[b * L + l] = p;
    }
    __syncthreads();
    if (q_in != nullptr && M <= static_cast<int>(blockDim.x)) {
        // persistent decoder: the memory rows of 16 positions at a time are staged in shared memory by the whole CTA (one
        // L2 round trip per 16 positions instead of one per 4), then every thread adds its column in position order — the
        // same sequence of fused multiply-adds as the loop below
        float* tile = ahs;                                    // [16][M] (the staged query input is not used in this mode)
        const int m = threadIdx.x;
        float acc = 0.f;
        for (int l0 = 0; l0 < len; l0 += 16) {
            const int nl = min(16, len - l0);
            __syncthreads();
            for (int i = threadIdx.x * 4; i < nl * M; i += blockDim.x * 4)
                *reinterpret_cast<float4*>(tile + i) =
                    __ldg(reinterpret_cast<const float4*>(memory + (static_cast<size_t>(b) * L + l0) * M + i));
            __syncthreads();
            if (m < M)
                for (int l = 0; l < nl; ++l) acc += e[l0 + l] * tile[l * M + m];
        }
        if (m < M) ctx[static_cast<size_t>(b) * M + m] = acc;
        return;
    }
    for (int m = threadIdx.x; m < M; m += blockDim.x) {
        // loads of four positions are issued before their multiply-adds (same summation order, four L2 round trips in
        // flight instead of one per position)
        float acc = 0.f;
        const float* mp = memory + static_cast<size_t>(b) * L * M + m;
        int l = 0;
        for (; l + 4 <= len; l += 4) {
            const float v0 = mp[static_cast<size_t>(l) * M], v1 = mp[static_cast<size_t>(l + 1) * M];
            const float v2 = mp[static_cast<size_t>(l + 2) * M], v3 = mp[static_cast<size_t>(l + 3) * M];
            acc += e[l] * v0;
            acc += e[l + 1] * v1;
            acc += e[l + 2] * v2;
            acc += e[l + 3] * v3;
        }
        for (; l < len; ++l) acc += e[l] * mp[static_cast<size_t>(l) * M];
        ctx[static_cast<size_t>(b) * M + m] = acc;
    }
}
__global__ void __launch_bounds__(256) t2_attention_kernel(const float* __restrict__ ah, const float* __restrict__ memory,
                                                           const float* __restrict__ pmem, const int* __restrict__ lens,
                                                           const __half* __restrict__ wq, const float* __restrict__ wloc_conv,
                                                           const float* __restrict__ wloc_dense, const float* __restrict__ v,
                                                           float* __restrict__ aw, float* __restrict__ awc,
                                                           float* __restrict__ ctx, float* __restrict__ align_out,
                                                           int L, int H, int M, int A, int NF, int KL) {
    extern __shared__ float sm[];
    t2_attention_body(sm, blockIdx.x, ah, memory, pmem, lens, wq, wloc_conv, wloc_dense, v, aw, awc, ctx, align_out, L, H, M, A,
                      NF, KL);
}

// mel frame + gate: rows 0..n_mel-1 of W are linear_projection, row n_mel is gate_layer; one warp per row
__global__ void __launch_bounds__(256) t2_project_kernel(const float* __restrict__ dh, const float* __restrict__ ctx,
                                                         const __half* __restrict__ w, const float* __restrict__ bias,
                                                         int H, int M, int n_mel, int B, float* __restrict__ frame,
                                                         float* __restrict__ mel_steps, int step, int max_steps,
                                                         float* __restrict__ gate) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int o = blockIdx.x * 8 + warp;
    if (o > n_mel) return;
    const __half* wr = w + static_cast<size_t>(o) * (H + M);
    for (int b = 0; b < B; ++b) {
        const float d = warp_dot_h<true>(wr, dh + static_cast<size_t>(b) * H, H, lane) +
                        warp_dot_h<true>(wr + H, ctx + static_cast<size_t>(b) * M, M, lane) + bias[o];
        if (lane == 0) {
            if (o < n_mel) {
                frame[b * n_mel + o] = d;
                mel_steps[(static_cast<size_t>(b) * max_steps + step) * n_mel + o] = d;
            } else {
                gate[b] = d;
            }
        }
    }
}

// mel_lens[~finished] += 1; finished |= sigmoid(gate) > thr; all_finished flag (torchaudio:846-852)
__global__ void t2_bookkeep_kernel(const float* __restrict__ gate, int* __restrict__ finished, int* __restrict__ mel_lens,
                                   int* __restrict__ done_step, int B, float thr, int step) {
    __shared__ int all;
    if (threadIdx.x == 0) all = 1;
    __syncthreads();
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        if (!finished[b]) mel_lens[b] += 1;
        if (sigmoidf_(gate[b]) > thr) finished[b] = 1;
        if (!finished[b]) atomicAnd(&all, 0);
    }
    __syncthreads();
    if (threadIdx.x == 0 && all && *done_step < 0) *done_step = step;
}

// ------------------------------------------------------------------------------------------------
// Persistent decoder: ONE cooperative launch runs a whole chunk of autoregressive steps (torchaudio _Decoder.infer
// :779-866 / decode :611-684). The six per-step launches above cost ~330 us per step at batch 8 on a B200 — host launch
// rate, not arithmetic (round-2 bench, config 4) — against a weight-bandwidth floor of ~5 us. Here the same phase bodies
// run inside a step loop, separated by grid-wide barriers:
//     A  attention LSTM cell      all CTAs (2 hidden units x 4 gates per virtual block, all utterances)
//     B  attention                one CTA per utterance
//     C  decoder LSTM cell        all CTAs
//     D  projection + stop bookkeeping + the NEXT step's prenet      one CTA per utterance
// i.e. four barriers per step. State other CTAs wrote is read with ld.global.cg (see ld_state4). The stop decision
// (torch.all(finished), torchaudio:850) is taken on the device after phase D; the host looks at it once per chunk.
// ------------------------------------------------------------------------------------------------
struct T2PersistArgs {
    int B, L, H, M, P, A, NF, KL, n_mel, max_steps, step0, n_steps, early_stop;
    float gate_threshold;
    const uint8_t* masks;            // [n_steps, 2, B, P]
    const int* lens;
    const float *memory, *pmem;
    float *ah[2], *ac, *dh[2], *dc, *aw, *awc, *ctx, *frame, *x, *gate, *frames, *align;
    int att_off;                     // float offset of the staged attention tables in dynamic shared memory (0 = not staged)
    float *q_buf, *x1;               // [B, A] query projections, [B, P] first prenet layer: written by all CTAs, read after a grid barrier
    int *finished, *mel_lens, *done_step;
    unsigned* bar;                   // grid barrier counter (zeroed before the launch)
    long long* timeline;             // debug (tools/t2_phases.py): CTA 0 and CTA B-1... accumulate cycles per phase, or null
    int* err_flag;
    const __half *pre_w0, *pre_w1, *arnn_wih, *arnn_whh, *drnn_wih, *drnn_whh, *w_query, *w_proj;
    const float *arnn_b, *drnn_b, *loc_conv, *loc_dense, *att_v, *b_proj;
};

// LSTMCell for the persistent decoder: a 512-thread CTA owns 8 hidden units (warp = 2 * unit + gate pair: every warp
// computes TWO gate rows per pass over the staged vectors, which halves the shared-memory reads that bound this phase —
// 16 warps x 86 KB instead of 32 x 86 KB per cell and step) and first stages
// the input vectors [x1 | x2 | h] of up to 8 utterances in shared memory — every gate row of the CTA multiplies the same
// vectors, and reading them per warp through ld.global.cg (no L1 in a persistent kernel, see ld_state4) cost ~250 MB of
// L2 traffic per cell and step. Same per-lane accumulation order as t2_lstm_cell_body: identical results.
__device__ __forceinline__ void t2_lstm_cell_staged(float* st, float (*g_s)[4][64], int cta, const float* __restrict__ x1,
                                                    int n1, const float* __restrict__ x2, int n2,
                                                    const float* __restrict__ h_in, float* __restrict__ h_out,
                                                    float* __restrict__ c, const __half* __restrict__ w_ih,
                                                    const __half* __restrict__ w_hh, const float* __restrict__ bias, int H,
                                                    int B) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int u = warp >> 1, g0 = (warp & 1) * 2;          // gates g0 and g0 + 1 of hidden unit u
    const int j = cta * 8 + u;
    const int n_in = n1 + n2, n_tot = n_in + H;
    const __half* wi[2] = {w_ih + static_cast<size_t>(g0 * H + j) * n_in, w_ih + static_cast<size_t>((g0 + 1) * H + j) * n_in};
    const __half* wh[2] = {w_hh + static_cast<size_t>(g0 * H + j) * H, w_hh + static_cast<size_t>((g0 + 1) * H + j) * H};
    for (int b0 = 0; b0 < B; b0 += 8) {
        const int nb = min(8, B - b0);
        const int quads = n_tot >> 2;
        for (int idx = threadIdx.x; idx < nb * quads; idx += blockDim.x) {
            const int q = idx / quads, k = (idx - q * quads) << 2;
            const float* src = k < n1 ? x1 + static_cast<size_t>(b0 + q) * n1 + k
                                      : (k < n_in ? x2 + static_cast<size_t>(b0 + q) * n2 + (k - n1)
                                                  : h_in + static_cast<size_t>(b0 + q) * H + (k - n_in));
            *reinterpret_cast<float4*>(st + q * n_tot + k) = ld_state4(src);
        }
        __syncthreads();
        float acc[2][8];
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int q = 0; q < 8; ++q) acc[r][q] = 0.f;
        for (int i0 = lane * 8; i0 < n_tot; i0 += 2 * 256) {
            // four weight loads (two rows x two column blocks) in flight per pass: see warp_dot_h
            uint4 wv[2][2];
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
                const int i = i0 + jj * 256;
#pragma unroll
                for (int r = 0; r < 2; ++r)
                    wv[jj][r] = i < n_tot ? __ldg(reinterpret_cast<const uint4*>(i < n_in ? wi[r] + i : wh[r] + (i - n_in)))
                                          : make_uint4(0, 0, 0, 0);
            }
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
                const int i = i0 + jj * 256;
                if (i < n_tot) {
                    float f[2][8];
                    unpack8(wv[jj][0], f[0]);
                    unpack8(wv[jj][1], f[1]);
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        if (q < nb) {
                            const float4 a = *reinterpret_cast<const float4*>(st + q * n_tot + i);
                            const float4 bb = *reinterpret_cast<const float4*>(st + q * n_tot + i + 4);
#pragma unroll
                            for (int r = 0; r < 2; ++r)
                                acc[r][q] += f[r][0] * a.x + f[r][1] * a.y + f[r][2] * a.z + f[r][3] * a.w + f[r][4] * bb.x +
                                             f[r][5] * bb.y + f[r][6] * bb.z + f[r][7] * bb.w;
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float v = warp_sum(acc[r][q]);
                if (lane == 0 && q < nb) g_s[u][g0 + r][b0 + q] = v + bias[(g0 + r) * H + j];
            }
        __syncthreads();
    }
    if (threadIdx.x < 8 * B) {
        const int uu = threadIdx.x / B, b = threadIdx.x % B;
        const int jj = cta * 8 + uu;
        const float ig = sigmoidf_(g_s[uu][0][b]), fg = sigmoidf_(g_s[uu][1][b]);
        const float gg = tanhf(g_s[uu][2][b]), og = sigmoidf_(g_s[uu][3][b]);
        const float cn = fg * c[static_cast<size_t>(b) * H + jj] + ig * gg;
        c[static_cast<size_t>(b) * H + jj] = cn;
        h_out[static_cast<size_t>(b) * H + jj] = og * tanhf(cn);
    }
    __syncthreads();
}

__device__ __forceinline__ void t2_grid_barrier(unsigned* bar, unsigned target, int* err_flag) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();                       // this CTA's writes are visible before its arrival is
        atomicAdd(bar, 1u);
        const long long t0 = clock64();
        unsigned seen;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(bar) : "memory");
            if (seen < target && clock64() - t0 > 4000000000ll) {      // a protocol bug must not hang the box
                if (err_flag) atomicExch(err_flag, 501);
                __trap();
            }
        } while (seen < target);
    }
    __syncthreads();
}

__global__ void __launch_bounds__(512, 1) t2_decoder_persistent_kernel(const T2PersistArgs a) {
    extern __shared__ float sm[];
    float (*g_s)[4][64] = reinterpret_cast<float (*)[4][64]>(sm);           // [8 units][4 gates][64 utterances]
    float* st = sm + 8 * 4 * 64;                                             // staged LSTM inputs, [8][n_in + H]
    const int G = gridDim.x;
    unsigned n_bar = 0;
    const int lstm_blocks = a.H / 8;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    const int gw = blockIdx.x * nwarp + warp, GW = G * nwarp;      // this warp among all warps of the grid
    // attention tables of this CTA's utterance, resident in shared memory for the whole launch (t2_attention_body)
    const float* att_tables = nullptr;
    if (a.att_off > 0 && a.B <= G && static_cast<int>(blockIdx.x) < a.B) {
        float* t = sm + a.att_off;
        const int n_conv = 2 * a.KL * a.NF, n_dense = a.NF * a.A, n_pm = a.L * a.A;
        for (int i = threadIdx.x; i < n_conv; i += blockDim.x) t[i] = a.loc_conv[i];
        for (int i = threadIdx.x; i < n_dense; i += blockDim.x) t[n_conv + i] = a.loc_dense[i];
        for (int i = threadIdx.x; i < a.A; i += blockDim.x) t[n_conv + n_dense + i] = a.att_v[i];
        for (int i = threadIdx.x; i < n_pm; i += blockDim.x)
            t[n_conv + n_dense + a.A + i] = a.pmem[static_cast<size_t>(blockIdx.x) * n_pm + i];
        att_tables = t;
        __syncthreads();
    }
    // prologue: the first step's prenet (later steps get theirs at the end of the previous step's phase D)
    for (int b = blockIdx.x; b < a.B; b += G)
        t2_prenet_body(sm, b, a.frame, a.pre_w0, a.pre_w1, a.masks, a.masks + static_cast<size_t>(a.B) * a.P, a.n_mel, a.P, a.x);
    t2_grid_barrier(a.bar, ++n_bar * G, a.err_flag);
    // debug: cycles of CTA 0 in [A, barrier, B, barrier, C, barrier, D, barrier], summed over the chunk's steps
    long long ph[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const bool tl = a.timeline != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
    long long t_prev = tl ? clock64() : 0;
    auto lap = [&](int k) {
        if (tl) { const long long t = clock64(); ph[k] += t - t_prev; t_prev = t; }
    };
    for (int i = 0; i < a.n_steps; ++i) {
        const int step = a.step0 + i;
        const int cur = step & 1, nxt = cur ^ 1;
        // A: attention LSTM cell on [prenet out | previous context]
        for (int vb = blockIdx.x; vb < lstm_blocks; vb += G)
            t2_lstm_cell_staged(st, g_s, vb, a.x, a.P, a.ctx, a.M, a.ah[cur], a.ah[nxt], a.ac, a.arnn_wih, a.arnn_whh, a.arnn_b, a.H, a.B);
        lap(0);
        t2_grid_barrier(a.bar, ++n_bar * G, a.err_flag);
        lap(1);
        // B1: query projections of all utterances, one dot product per warp of the whole grid (they were A / 16 sequential
        // L2 round trips inside each utterance's CTA: tools/t2_phases.py)
        for (int idx = gw; idx < a.B * a.A; idx += GW) {
            const int b = idx / a.A, row = idx - b * a.A;
            const float d = warp_dot_h<true>(a.w_query + static_cast<size_t>(row) * a.H, a.ah[nxt] + static_cast<size_t>(b) * a.H, a.H, lane);
            if (lane == 0) a.q_buf[idx] = d;
        }
        t2_grid_barrier(a.bar, ++n_bar * G, a.err_flag);
        // B2: location-sensitive attention, one utterance per CTA
        for (int b = blockIdx.x; b < a.B; b += G) {
            t2_attention_body(sm, b, a.ah[nxt], a.memory, a.pmem, a.lens, a.w_query, a.loc_conv, a.loc_dense, a.att_v, a.aw, a.awc,
                              a.ctx, a.align + static_cast<size_t>(step) * a.B * a.L, a.L, a.H, a.M, a.A, a.NF, a.KL, a.q_buf, att_tables);
            __syncthreads();
        }
        lap(2);
        t2_grid_barrier(a.bar, ++n_bar * G, a.err_flag);
        lap(3);
        // C: decoder LSTM cell on [attention hidden | context]
        for (int vb = blockIdx.x; vb < lstm_blocks; vb += G)
            t2_lstm_cell_staged(st, g_s, vb, a.ah[nxt], a.H, a.ctx, a.M, a.dh[cur], a.dh[nxt], a.dc, a.drnn_wih, a.drnn_whh, a.drnn_b, a.H, a.B);
        lap(4);
        t2_grid_barrier(a.bar, ++n_bar * G, a.err_flag);
        lap(5);
        // D1: mel frame + gate, one (utterance, output row) dot product per warp of the whole grid; the warp that owns a gate
        // row does the stop bookkeeping. D2 / D3: the two prenet layers of the NEXT step, again one row per warp, a grid
        // barrier between the layers. (One CTA per utterance ran these ~600 dot products as 40 sequential L2 round trips
        // per warp: 29 % of the step.) Same warp_dot_h calls on the same values as the per-utterance bodies.
        for (int idx = gw; idx < a.B * (a.n_mel + 1); idx += GW) {
            const int b = idx / (a.n_mel + 1), o = idx - b * (a.n_mel + 1);
            const __half* wr = a.w_proj + static_cast<size_t>(o) * (a.H + a.M);
            const float d = warp_dot_h<true>(wr, a.dh[nxt] + static_cast<size_t>(b) * a.H, a.H, lane) +
                            warp_dot_h<true>(wr + a.H, a.ctx + static_cast<size_t>(b) * a.M, a.M, lane) + a.b_proj[o];
            if (lane == 0) {
                if (o < a.n_mel) {
                    a.frame[b * a.n_mel + o] = d;
                    a.frames[(static_cast<size_t>(b) * a.max_steps + step) * a.n_mel + o] = d;
                } else {
                    a.gate[b] = d;
                    // mel_lens[~finished] += 1; finished |= sigmoid(gate) > thr   (torchaudio:846-849)
                    if (!a.finished[b]) a.mel_lens[b] += 1;
                    if (sigmoidf_(d) > a.gate_threshold) a.finished[b] = 1;
                }
            }
        }
        if (i + 1 < a.n_steps) {
            const uint8_t* m0 = a.masks + (static_cast<size_t>(i + 1) * 2 + 0) * a.B * a.P;
            const uint8_t* m1 = m0 + static_cast<size_t>(a.B) * a.P;
            t2_grid_barrier(a.bar, ++n_bar * G, a.err_flag);
            for (int idx = gw; idx < a.B * a.P; idx += GW) {
                const int b = idx / a.P, o = idx - b * a.P;
                const float d = warp_dot_h<true>(a.pre_w0 + static_cast<size_t>(o) * a.n_mel, a.frame + static_cast<size_t>(b) * a.n_mel, a.n_mel, lane);
                if (lane == 0) a.x1[idx] = fmaxf(d, 0.f) * (m0[idx] ? 2.f : 0.f);
            }
            t2_grid_barrier(a.bar, ++n_bar * G, a.err_flag);
            for (int idx = gw; idx < a.B * a.P; idx += GW) {
                const int b = idx / a.P, o = idx - b * a.P;
                const float d = warp_dot_h<true>(a.pre_w1 + static_cast<size_t>(o) * a.P, a.x1 + static_cast<size_t>(b) * a.P, a.P, lane);
                if (lane == 0) a.x[idx] = fmaxf(d, 0.f) * (m1[idx] ? 2.f : 0.f);
            }
        }
        lap(6);
        t2_grid_barrier(a.bar, ++n_bar * G, a.err_flag);
        lap(7);
        // every CTA takes the same stop decision from the same flags
        bool all = true;
        for (int b = 0; b < a.B; ++b) all = all && (__ldcg(a.finished + b) != 0);
        if (all) {
            if (blockIdx.x == 0 && threadIdx.x == 0 && *a.done_step < 0) *a.done_step = step;
            if (a.early_stop) break;
        }
    }
    if (tl)
        for (int k = 0; k < 8; ++k) a.timeline[256 * 128 + k] += ph[k];     // past the slots the conv kernels stamp
}

// frames [B, max_steps, n_mel] fp32 -> channel-last fp16 [B, T, ld] (zero padded channels) for the postnet
__global__ void t2_frames_to_cl_kernel(const float* __restrict__ frames, int max_steps, int n_mel, int T, int ld,
                                       __half* __restrict__ out) {
    const int r = blockIdx.x;  // b*T + t
    const int b = r / T, t = r % T;
    for (int j = threadIdx.x; j < ld; j += blockDim.x)
        out[static_cast<size_t>(r) * ld + j] = __float2half(j < n_mel ? frames[(static_cast<size_t>(b) * max_steps + t) * n_mel + j] : 0.f);
}

// final mel = frames + postnet (both fp32), transposed to [B, n_mel, T]; plus the channel-last fp16 copy
__global__ void t2_finalize_kernel(const float* __restrict__ frames, const float* __restrict__ post_t, int max_steps,
                                   int n_mel, int T, int ld, float* __restrict__ mel_out, __half* __restrict__ mel_cl,
                                   const int* __restrict__ mel_lens) {
    const int b = blockIdx.y;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    const int len = mel_lens[b];
    for (int c = 0; c < n_mel; ++c) {
        const float v = frames[(static_cast<size_t>(b) * max_steps + t) * n_mel + c] + post_t[(static_cast<size_t>(b) * n_mel + c) * T + t];
        mel_out[(static_cast<size_t>(b) * n_mel + c) * T + t] = v;
        if (mel_cl) mel_cl[(static_cast<size_t>(b) * T + t) * ld + c] = __float2half(t < len ? v : 0.f);
    }
    if (mel_cl)
        for (int c = n_mel; c < ld; ++c) mel_cl[(static_cast<size_t>(b) * T + t) * ld + c] = __float2half(0.f);
}

// ------------------------------------------------------------------------------------------------
// wrapper post-processing on the device (models/tacotron2/networks.py:44-67, 192-206): per utterance
//   truncate_mel  cut where the attention on the inserted separator first reaches 80 % of its maximum, replicate the
//                 last kept frame three times
//   resize_mel    bicubic time resize to int(len / rate) frames (torch.nn.functional.interpolate, mode='bicubic',
//                 align_corners=False, A = -0.75; the mel-bin axis keeps its size, so the 2-D kernel is 1-D in time)
// One CTA per utterance; the reference does this in a Python loop with a host sync per utterance.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float cubic1(float x, float A) { return ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f; }
__device__ __forceinline__ float cubic2(float x, float A) { return ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A; }

__global__ void __launch_bounds__(256) t2_postprocess_kernel(const float* __restrict__ mel, const int* __restrict__ mel_lens,
                                                             const float* __restrict__ align, const int* __restrict__ cols,
                                                             double inv_rate, int do_resize, int n_mel, int T, int L, int T_out,
                                                             float* __restrict__ out, int* __restrict__ out_lens) {
    __shared__ float red_f[32];
    __shared__ int red_i[32];
    __shared__ int s_keep, s_cut, s_new;
    const int b = blockIdx.x;
    const int n = min(mel_lens[b], T);
    const int col = cols ? cols[b] : -1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    if (col >= 0 && n > 0) {
        const float* ps = align + static_cast<size_t>(b) * T * L + col;
        float mx = -INFINITY;
        for (int i = threadIdx.x; i < n; i += blockDim.x) mx = fmaxf(mx, ps[static_cast<size_t>(i) * L]);
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if (lane == 0) red_f[warp] = mx;
        __syncthreads();
        mx = red_f[0];
        for (int i = 1; i < nwarp; ++i) mx = fmaxf(mx, red_f[i]);
        const float thr = 0.8f * mx;
        int first = n;
        for (int i = threadIdx.x; i < n; i += blockDim.x)
            if (ps[static_cast<size_t>(i) * L] >= thr) { first = i; break; }
        for (int o = 16; o > 0; o >>= 1) first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
        if (lane == 0) red_i[warp] = first;
        __syncthreads();
        if (threadIdx.x == 0) {
            int f = red_i[0];
            for (int i = 1; i < nwarp; ++i) f = min(f, red_i[i]);
            s_keep = max(f, 1);           // frames kept before the three replicated ones (>= 1: replicate needs a frame)
            s_cut = s_keep + 3;
        }
    } else if (threadIdx.x == 0) {
        s_keep = n;
        s_cut = n;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        // Nt_new = int(1 / rate * Nt)  (networks.py:62): the same two IEEE double operations as the Python expression
        int nn = s_cut;
        if (do_resize) nn = static_cast<int>(inv_rate * static_cast<double>(s_cut));
        s_new = max(0, min(nn, T_out));
        out_lens[b] = s_new;
    }
    __syncthreads();
    const int keep = s_keep, cut = s_cut, n_new = s_new;
    const float* mb = mel + static_cast<size_t>(b) * n_mel * T;
    float* ob = out + static_cast<size_t>(b) * n_mel * T_out;
    const bool same = n_new == cut;
    const float scale = n_new > 0 ? static_cast<float>(cut) / static_cast<float>(n_new) : 0.f;
    const float A = -0.75f;
    for (int idx = threadIdx.x; idx < n_mel * T_out; idx += blockDim.x) {
        const int c = idx / T_out, x = idx - c * T_out;
        float v = 0.f;
        if (x < n_new) {
            const float* row = mb + static_cast<size_t>(c) * T;
            auto at = [&](int i) {                                // mel_cut with replicate padding, index clamped like ATen
                i = min(max(i, 0), cut - 1);
                return row[min(i, keep - 1)];
            };
            if (same) {
                v = at(x);
            } else {
                const float src = scale * (static_cast<float>(x) + 0.5f) - 0.5f;
                const float fl = floorf(src);
                const int ix = static_cast<int>(fl);
                const float t = src - fl;
                const float w0 = cubic2(t + 1.f, A), w1 = cubic1(t, A), w2 = cubic1(1.f - t, A), w3 = cubic2(2.f - t, A);
                v = at(ix - 1) * w0 + at(ix) * w1 + at(ix + 1) * w2 + at(ix + 2) * w3;
            }
        }
        ob[idx] = v;
    }
}

int upload_h16(const float* h, size_t n, __half** d) {
    std::vector<__half> tmp(n);
    for (size_t i = 0; i < n; ++i) tmp[i] = __float2half(h[i]);
    TTSB_CHECK_CUDA(cudaMalloc(d, n * sizeof(__half)));
    TTSB_CHECK_CUDA(cudaMemcpy(*d, tmp.data(), n * sizeof(__half), cudaMemcpyHostToDevice));
    return 0;
}

// Conv1d + BatchNorm(eval) folded: w' = w * g/sqrt(var+eps), b' = (b - mean) * g/sqrt(var+eps) + beta
int make_conv_bn_layer(ConvLayer& L, const TensorTable& tab, const std::string& p, int cout, int cin, int k,
                       int cin_stored, int cout_pad) {
    TTSB_GET_TENSOR(w, tab, p + ".0.weight", 3);
    TTSB_GET_TENSOR(b, tab, p + ".0.bias", 1);
    TTSB_GET_TENSOR(g, tab, p + ".1.weight", 1);
    TTSB_GET_TENSOR(be, tab, p + ".1.bias", 1);
    TTSB_GET_TENSOR(mu, tab, p + ".1.running_mean", 1);
    TTSB_GET_TENSOR(var, tab, p + ".1.running_var", 1);
    TTSB_REQUIRE(w->shape[0] == cout && w->shape[1] == cin && w->shape[2] == k, p + " conv shape");
    std::vector<float> wf(static_cast<size_t>(cout_pad) * cin * k, 0.f), bf(cout_pad, 0.f);
    for (int co = 0; co < cout; ++co) {
        const float s = g->h_data[co] / std::sqrt(var->h_data[co] + 1e-5f);
        for (int i = 0; i < cin * k; ++i) wf[static_cast<size_t>(co) * cin * k + i] = w->h_data[static_cast<size_t>(co) * cin * k + i] * s;
        bf[co] = (b->h_data[co] - mu->h_data[co]) * s + be->h_data[co];
    }
    return make_conv1d_layer(L, wf.data(), bf.data(), cout_pad, cin, k, 1, cin_stored, 0);
}

}  // namespace

struct ttsb_tacotron2 {
    int device = 0;
    int n_symbol = 0, E = 512, H = 1024, S = 0, P = 256, A = 128, NF = 32, KL = 31, n_mel = 80, mel_ld = 128;
    int num_speakers = 0;
    int M = 512;  // memory width = E + S
    float* emb = nullptr;
    ConvLayer enc_conv[3];
    ConvLayer enc_xproj;          // 512 -> 2*4*256 (forward | backward gate pre-activations)
    __half* enc_whh = nullptr;    // [2][4*256][256]
    float* spk_emb = nullptr;
    __half* w_mem = nullptr;      // [A, M]
    __half *pre_w0 = nullptr, *pre_w1 = nullptr;
    __half *arnn_wih = nullptr, *arnn_whh = nullptr, *drnn_wih = nullptr, *drnn_whh = nullptr;
    float *arnn_b = nullptr, *drnn_b = nullptr;
    __half* w_query = nullptr;
    float *loc_conv = nullptr, *loc_dense = nullptr, *att_v = nullptr;
    __half* w_proj = nullptr;     // [n_mel + 1, H + M]
    float* b_proj = nullptr;
    ConvLayer post[5];
};

namespace {

struct T2State {   // device state of one batch, carved from the caller's state buffer
    int* lens; int* finished; int* mel_lens; int* done_step; unsigned* bar; int64_t* spk;
    float *memory, *pmem, *ah[2], *ac, *dh[2], *dc, *aw, *awc, *ctx, *frame, *x, *gate, *frames, *align, *q_buf, *x1;
};

T2State carve_t2(const ttsb_tacotron2* h, void* p, int B, int L, int max_steps, size_t* bytes) {
    Carver c(p);
    T2State s;
    s.lens = c.take<int>(B); s.finished = c.take<int>(B); s.mel_lens = c.take<int>(B); s.done_step = c.take<int>(4);
    s.bar = c.take<unsigned>(4);
    s.spk = c.take<int64_t>(B);
    s.memory = c.take<float>(static_cast<size_t>(B) * L * h->M);
    s.pmem = c.take<float>(static_cast<size_t>(B) * L * h->A);
    for (int i = 0; i < 2; ++i) { s.ah[i] = c.take<float>(static_cast<size_t>(B) * h->H); s.dh[i] = c.take<float>(static_cast<size_t>(B) * h->H); }
    s.ac = c.take<float>(static_cast<size_t>(B) * h->H); s.dc = c.take<float>(static_cast<size_t>(B) * h->H);
    s.aw = c.take<float>(static_cast<size_t>(B) * L); s.awc = c.take<float>(static_cast<size_t>(B) * L);
    s.ctx = c.take<float>(static_cast<size_t>(B) * h->M); s.frame = c.take<float>(static_cast<size_t>(B) * h->n_mel);
    s.x = c.take<float>(static_cast<size_t>(B) * h->P); s.gate = c.take<float>(B);
    s.frames = c.take<float>(static_cast<size_t>(B) * max_steps * h->n_mel);
    s.align = c.take<float>(static_cast<size_t>(max_steps) * B * L);
    s.q_buf = c.take<float>(static_cast<size_t>(B) * h->A);
    s.x1 = c.take<float>(static_cast<size_t>(B) * h->P);
    if (bytes) *bytes = c.off + 256;
    return s;
}

}  // namespace

extern "C" {

int ttsb_tacotron2_create(const ttsb_tensor_t* weights, int n_weights, int device, ttsb_tacotron2_t** out) {
    return guarded_call([&]() -> int {
    TTSB_REQUIRE(weights && out, "null argument");
    TTSB_DEVICE_GUARD(device);
    TensorTable tab(weights, n_weights);
    std::unique_ptr<ttsb_tacotron2, void (*)(ttsb_tacotron2*)> owner(new ttsb_tacotron2(), ttsb_tacotron2_destroy);
    ttsb_tacotron2* h = owner.get();
    h->device = device;
    TTSB_GET_TENSOR(emb, tab, "embedding.weight", 2);
    h->n_symbol = static_cast<int>(emb->shape[0]);
    TTSB_REQUIRE(emb->shape[1] == h->E, "embedding dim 512 expected");
    TTSB_PROPAGATE(upload_f32(emb->h_data, TensorTable::numel(emb), &h->emb));
    for (int i = 0; i < 3; ++i)
        TTSB_PROPAGATE(make_conv_bn_layer(h->enc_conv[i], tab, "encoder.convolutions." + std::to_string(i), h->E, h->E, 5, h->E, h->E));
    {
        const int Hh = h->E / 2, G = 4 * Hh;
        std::vector<float> w(static_cast<size_t>(2) * G * h->E), b(2 * G);
        std::vector<float> whh(static_cast<size_t>(2) * G * Hh);
        const char* suf[2] = {"", "_reverse"};
        for (int d = 0; d < 2; ++d) {
            TTSB_GET_TENSOR(wi, tab, std::string("encoder.lstm.weight_ih_l0") + suf[d], 2);
            TTSB_GET_TENSOR(wh, tab, std::string("encoder.lstm.weight_hh_l0") + suf[d], 2);
            TTSB_GET_TENSOR(bi, tab, std::string("encoder.lstm.bias_ih_l0") + suf[d], 1);
            TTSB_GET_TENSOR(bh, tab, std::string("encoder.lstm.bias_hh_l0") + suf[d], 1);
            TTSB_REQUIRE(wi->shape[0] == G && wi->shape[1] == h->E && wh->shape[1] == Hh, "encoder LSTM shapes");
            std::copy(wi->h_data, wi->h_data + static_cast<size_t>(G) * h->E, w.begin() + static_cast<size_t>(d) * G * h->E);
            std::copy(wh->h_data, wh->h_data + static_cast<size_t>(G) * Hh, whh.begin() + static_cast<size_t>(d) * G * Hh);
            for (int i = 0; i < G; ++i) b[d * G + i] = bi->h_data[i] + bh->h_data[i];
        }
        const int off[1] = {0};
        TTSB_PROPAGATE(conv_layer_create(h->enc_xproj, h->E, h->E, 2 * G, 1, off, nullptr, 1 << 30, w.data(), b.data(), 256));
        TTSB_PROPAGATE(upload_h16(whh.data(), whh.size(), &h->enc_whh));
    }
    if (const ttsb_tensor_t* se = tab.find("speaker_embedding.weight")) {
        h->num_speakers = static_cast<int>(se->shape[0]);
        h->S = static_cast<int>(se->shape[1]);
        TTSB_PROPAGATE(upload_f32(se->h_data, TensorTable::numel(se), &h->spk_emb));
    }
    h->M = h->E + h->S;
    TTSB_REQUIRE(h->M % 8 == 0, "memory width must be a multiple of 8");
    const std::string A = "decoder.attention_layer.";
    {
        TTSB_GET_TENSOR(wm, tab, A + "memory_layer.weight", 2);
        TTSB_REQUIRE(wm->shape[0] == h->A && wm->shape[1] == h->M, "memory_layer shape");
        TTSB_PROPAGATE(upload_h16(wm->h_data, TensorTable::numel(wm), &h->w_mem));
        TTSB_GET_TENSOR(wq, tab, A + "query_layer.weight", 2);
        TTSB_REQUIRE(wq->shape[0] == h->A && wq->shape[1] == h->H, "query_layer shape");
        TTSB_PROPAGATE(upload_h16(wq->h_data, TensorTable::numel(wq), &h->w_query));
        TTSB_GET_TENSOR(vv, tab, A + "v.weight", 2);
        TTSB_PROPAGATE(upload_f32(vv->h_data, h->A, &h->att_v));
        TTSB_GET_TENSOR(lc, tab, A + "location_layer.location_conv.weight", 3);
        TTSB_REQUIRE(lc->shape[0] == h->NF && lc->shape[1] == 2 && lc->shape[2] == h->KL, "location conv shape");
        std::vector<float> lct(TensorTable::numel(lc));                     // [NF][2][KL] -> [2][KL][NF] (t2_attention_body)
        for (int nf = 0; nf < h->NF; ++nf)
            for (int c2 = 0; c2 < 2; ++c2)
                for (int k = 0; k < h->KL; ++k)
                    lct[(static_cast<size_t>(c2) * h->KL + k) * h->NF + nf] = lc->h_data[(static_cast<size_t>(nf) * 2 + c2) * h->KL + k];
        TTSB_PROPAGATE(upload_f32(lct.data(), lct.size(), &h->loc_conv));
        TTSB_GET_TENSOR(ld, tab, A + "location_layer.location_dense.weight", 2);
        TTSB_REQUIRE(ld->shape[0] == h->A && ld->shape[1] == h->NF, "location dense shape");
        std::vector<float> ldt(static_cast<size_t>(h->A) * h->NF);          // [A][NF] -> [NF][A] (t2_attention_body)
        for (int a_ = 0; a_ < h->A; ++a_)
            for (int nf = 0; nf < h->NF; ++nf) ldt[static_cast<size_t>(nf) * h->A + a_] = ld->h_data[static_cast<size_t>(a_) * h->NF + nf];
        TTSB_PROPAGATE(upload_f32(ldt.data(), ldt.size(), &h->loc_dense));
    }
    {
        TTSB_GET_TENSOR(p0, tab, "decoder.prenet.layers.0.weight", 2);
        TTSB_GET_TENSOR(p1, tab, "decoder.prenet.layers.1.weight", 2);
        TTSB_REQUIRE(p0->shape[0] == h->P && p0->shape[1] == h->n_mel && p1->shape[0] == h->P, "prenet shapes");
        TTSB_PROPAGATE(upload_h16(p0->h_data, TensorTable::numel(p0), &h->pre_w0));
        TTSB_PROPAGATE(upload_h16(p1->h_data, TensorTable::numel(p1), &h->pre_w1));
    }
    auto load_cell = [&](const std::string& p, int n_in, __half** wih, __half** whh, float** bias) -> int {
        TTSB_GET_TENSOR(wi, tab, p + ".weight_ih", 2);
        TTSB_GET_TENSOR(wh, tab, p + ".weight_hh", 2);
        TTSB_GET_TENSOR(bi, tab, p + ".bias_ih", 1);
        TTSB_GET_TENSOR(bh, tab, p + ".bias_hh", 1);
        TTSB_REQUIRE(wi->shape[0] == 4 * h->H && wi->shape[1] == n_in && wh->shape[1] == h->H, p + " shapes");
        TTSB_PROPAGATE(upload_h16(wi->h_data, TensorTable::numel(wi), wih));
        TTSB_PROPAGATE(upload_h16(wh->h_data, TensorTable::numel(wh), whh));
        std::vector<float> b(4 * h->H);
        for (int i = 0; i < 4 * h->H; ++i) b[i] = bi->h_data[i] + bh->h_data[i];
        return upload_f32(b.data(), b.size(), bias);
    };
    TTSB_PROPAGATE(load_cell("decoder.attention_rnn", h->P + h->M, &h->arnn_wih, &h->arnn_whh, &h->arnn_b));
    TTSB_PROPAGATE(load_cell("decoder.decoder_rnn", h->H + h->M, &h->drnn_wih, &h->drnn_whh, &h->drnn_b));
    {
        TTSB_GET_TENSOR(lp, tab, "decoder.linear_projection.weight", 2);
        TTSB_GET_TENSOR(lb, tab, "decoder.linear_projection.bias", 1);
        TTSB_GET_TENSOR(gw, tab, "decoder.gate_layer.weight", 2);
        TTSB_GET_TENSOR(gb, tab, "decoder.gate_layer.bias", 1);
        const int K = h->H + h->M;
        TTSB_REQUIRE(lp->shape[0] == h->n_mel && lp->shape[1] == K && gw->shape[1] == K, "projection shapes");
        std::vector<float> w(static_cast<size_t>(h->n_mel + 1) * K), b(h->n_mel + 1);
        std::copy(lp->h_data, lp->h_data + static_cast<size_t>(h->n_mel) * K, w.begin());
        std::copy(gw->h_data, gw->h_data + K, w.begin() + static_cast<size_t>(h->n_mel) * K);
        std::copy(lb->h_data, lb->h_data + h->n_mel, b.begin());
        b[h->n_mel] = gb->h_data[0];
        TTSB_PROPAGATE(upload_h16(w.data(), w.size(), &h->w_proj));
        TTSB_PROPAGATE(upload_f32(b.data(), b.size(), &h->b_proj));
    }
    const int dims[6] = {h->n_mel, 512, 512, 512, 512, h->n_mel};
    for (int i = 0; i < 5; ++i)
        TTSB_PROPAGATE(make_conv_bn_layer(h->post[i], tab, "postnet.convolutions." + std::to_string(i), dims[i + 1], dims[i], 5,
                                          i == 0 ? h->mel_ld : 512, i == 4 ? h->mel_ld : 512));
    *out = owner.release();
    return 0;
    });
}

void ttsb_tacotron2_destroy(ttsb_tacotron2_t* h) {
    if (!h) return;
    for (auto& l : h->enc_conv) conv_layer_destroy(l);
    conv_layer_destroy(h->enc_xproj);
    for (auto& l : h->post) conv_layer_destroy(l);
    void* ptrs[] = {h->emb, h->enc_whh, h->spk_emb, h->w_mem, h->pre_w0, h->pre_w1, h->arnn_wih, h->arnn_whh, h->drnn_wih,
                    h->drnn_whh, h->arnn_b, h->drnn_b, h->w_query, h->loc_conv, h->loc_dense, h->att_v, h->w_proj, h->b_proj};
    for (void* p : ptrs) if (p) cudaFree(p);
    delete h;
}

size_t ttsb_tacotron2_state_bytes(const ttsb_tacotron2_t* h, int B, int L, int max_steps) {
    size_t n = 0;
    carve_t2(h, nullptr, B, L, max_steps, &n);
    return n;
}

size_t ttsb_tacotron2_workspace_bytes(const ttsb_tacotron2_t* h, int B, int L, int T) {
    const size_t rows = static_cast<size_t>(B) * std::max(L, T);
    return rows * (512 * 2 * 2 + 2048 * 4 + 512 * 4 + 128 * 2) + static_cast<size_t>(B) * h->n_mel * std::max(T, 1) * 4 + 4096;
}

/* tokens [B,L] int64 (padded), lengths [B] int32, speaker ids [B] int64 -> encoder + decoder state reset */
int ttsb_tacotron2_encode(ttsb_tacotron2_t* h, const int64_t* d_tokens, const int32_t* d_lengths,
                          const int64_t* d_speaker_ids, int B, int L, int max_steps, void* d_state, void* d_workspace,
                          size_t workspace_bytes, void* stream_) {
    return guarded_call([&]() -> int {
    TTSB_REQUIRE(h && d_tokens && d_lengths && d_state && d_workspace, "null argument");
    TTSB_REQUIRE(B > 0 && B <= 64 && L > 0 && max_steps > 0, "batch must be 1..64");
    TTSB_REQUIRE(h->spk_emb == nullptr || d_speaker_ids != nullptr, "speaker ids required for a multi-speaker model");
    TTSB_REQUIRE(workspace_bytes >= ttsb_tacotron2_workspace_bytes(h, B, L, 0), "workspace too small");
    TTSB_DEVICE_GUARD(h->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream_);
    ConvRuntime rt;
    TTSB_PROPAGATE(get_conv_runtime(static_cast<size_t>(B) * L * 2048, rt));
    T2State st = carve_t2(h, d_state, B, L, max_steps, nullptr);
    Carver c(d_workspace);
    const size_t rows = static_cast<size_t>(B) * L;
    __half* xa = c.take<__half>(rows * 512);
    __half* xb = c.take<__half>(rows * 512);
    float* xproj = c.take<float>(rows * 2048);
    float* enc = c.take<float>(rows * 512);

    TTSB_CHECK_CUDA(cudaMemcpyAsync(st.lens, d_lengths, B * sizeof(int), cudaMemcpyDeviceToDevice, s));
    if (d_speaker_ids) TTSB_CHECK_CUDA(cudaMemcpyAsync(st.spk, d_speaker_ids, B * sizeof(int64_t), cudaMemcpyDeviceToDevice, s));
    TTSB_CHECK_CUDA(cudaMemsetAsync(st.done_step, 0xFF, 4 * sizeof(int), s));       // [0] done step = -1
    TTSB_CHECK_CUDA(cudaMemsetAsync(st.done_step + 1, 0, sizeof(int), s));           // [1] input status bits
    t2_embed_kernel<<<B * L, 128, 0, s>>>(d_tokens, h->emb, h->E, h->n_symbol, st.done_step + 1, xa);
    count_launch();
    __half* cur = xa; __half* nxt = xb;
    for (int i = 0; i < 3; ++i) {     // no masking between the convs, exactly like the reference
        EpiParams e;
        e.out_act = nxt; e.ld_act = h->E; e.act_slope = 0.f;
        TTSB_PROPAGATE(conv_forward(h->enc_conv[i], rt, cur, h->E, B, L, e, s));
        std::swap(cur, nxt);
    }
    {
        EpiParams e;
        e.out_f32 = xproj; e.ld_f32 = 2048;
        TTSB_PROPAGATE(conv_forward(h->enc_xproj, rt, cur, h->E, B, L, e, s));
    }
    const int Hh = h->E / 2;
    t2_bilstm_kernel<<<dim3(B, 2), 1024, (2 * Hh + 4 * Hh) * sizeof(float), s>>>(xproj, h->enc_whh, st.lens, L, Hh, enc);
    count_launch();
    t2_memory_kernel<<<B * L, 128, h->M * sizeof(float), s>>>(enc, h->spk_emb, st.spk, h->w_mem, L, h->E, h->S, h->A,
                                                              h->num_speakers, st.done_step + 1, st.memory, st.pmem);
    count_launch();
    // decoder state reset (torchaudio:_initialize_decoder_states, _get_go_frame)
    float* zero_f[] = {st.ah[0], st.ah[1], st.ac, st.dh[0], st.dh[1], st.dc};
    for (float* p : zero_f) TTSB_CHECK_CUDA(cudaMemsetAsync(p, 0, static_cast<size_t>(B) * h->H * sizeof(float), s));
    TTSB_CHECK_CUDA(cudaMemsetAsync(st.aw, 0, static_cast<size_t>(B) * L * sizeof(float), s));
    TTSB_CHECK_CUDA(cudaMemsetAsync(st.awc, 0, static_cast<size_t>(B) * L * sizeof(float), s));
    TTSB_CHECK_CUDA(cudaMemsetAsync(st.ctx, 0, static_cast<size_t>(B) * h->M * sizeof(float), s));
    TTSB_CHECK_CUDA(cudaMemsetAsync(st.frame, 0, static_cast<size_t>(B) * h->n_mel * sizeof(float), s));
    TTSB_CHECK_CUDA(cudaMemsetAsync(st.finished, 0, B * sizeof(int), s));
    TTSB_CHECK_CUDA(cudaMemsetAsync(st.mel_lens, 0, B * sizeof(int), s));
    TTSB_CHECK_CUDA(cudaGetLastError());
    return 0;
    });
}

/* Runs decoder steps [step0, step0 + n_steps). d_masks: [n_steps, 2, B, P] bytes (1 = keep) for the
 * prenet dropout. h_done_step (host, optional): after the call (synchronises the stream) receives the first
 * step index at which every utterance had finished, or -1. */
int ttsb_tacotron2_decode(ttsb_tacotron2_t* h, int B, int L, int max_steps, int step0, int n_steps,
                          const uint8_t* d_masks, float gate_threshold, int early_stop, void* d_state, int* h_done_step,
                          void* stream_) {
    return guarded_call([&]() -> int {
    TTSB_REQUIRE(h && d_masks && d_state, "null argument");
    TTSB_REQUIRE(step0 >= 0 && n_steps > 0 && step0 + n_steps <= max_steps, "step range");
    TTSB_DEVICE_GUARD(h->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream_);
    T2State st = carve_t2(h, d_state, B, L, max_steps, nullptr);
    const int H = h->H, M = h->M, P = h->P;
    const size_t att_smem = (h->A + 2 * (L + h->KL) + L + 32 + 4 + h->H) * sizeof(float);
    static const int want_persistent = getenv("TTSB_T2_PERSISTENT") ? atoi(getenv("TTSB_T2_PERSISTENT")) : 1;
    if (want_persistent) {
        // one cooperative launch for the whole chunk (t2_decoder_persistent_kernel)
        T2PersistArgs a;
        a.B = B; a.L = L; a.H = H; a.M = M; a.P = P; a.A = h->A; a.NF = h->NF; a.KL = h->KL; a.n_mel = h->n_mel;
        a.max_steps = max_steps; a.step0 = step0; a.n_steps = n_steps; a.early_stop = early_stop ? 1 : 0;
        a.gate_threshold = gate_threshold;
        a.masks = d_masks; a.lens = st.lens; a.memory = st.memory; a.pmem = st.pmem;
        a.ah[0] = st.ah[0]; a.ah[1] = st.ah[1]; a.ac = st.ac; a.dh[0] = st.dh[0]; a.dh[1] = st.dh[1]; a.dc = st.dc;
        a.aw = st.aw; a.awc = st.awc; a.ctx = st.ctx; a.frame = st.frame; a.x = st.x; a.gate = st.gate; a.frames = st.frames;
        a.q_buf = st.q_buf; a.x1 = st.x1;
        a.align = st.align; a.finished = st.finished; a.mel_lens = st.mel_lens; a.done_step = st.done_step; a.bar = st.bar;
        ConvRuntime rt;
        TTSB_PROPAGATE(get_conv_runtime(0, rt));
        a.err_flag = rt.err_flag;
        a.timeline = rt.timeline;
        a.pre_w0 = h->pre_w0; a.pre_w1 = h->pre_w1; a.arnn_wih = h->arnn_wih; a.arnn_whh = h->arnn_whh;
        a.drnn_wih = h->drnn_wih; a.drnn_whh = h->drnn_whh; a.w_query = h->w_query; a.w_proj = h->w_proj;
        a.arnn_b = h->arnn_b; a.drnn_b = h->drnn_b; a.loc_conv = h->loc_conv; a.loc_dense = h->loc_dense; a.att_v = h->att_v;
        a.b_proj = h->b_proj;
        const size_t lstm_smem = (8 * 4 * 64 + static_cast<size_t>(8) * (std::max(P, H) + M + H)) * sizeof(float);
        const size_t att_smem_p = (h->A + 2 * (L + h->KL) + L + 32 + 4 + std::max(h->H, 16 * M)) * sizeof(float);   // + the context tile
        const size_t smem = std::max({att_smem_p, static_cast<size_t>(h->n_mel + P) * sizeof(float), lstm_smem,
                                      static_cast<size_t>(H + M) * sizeof(float)});
        // + the attention tables (conv, dense, v, one utterance's processed memory) behind the phases' scratch, room permitting
        const size_t tables = (static_cast<size_t>(2) * h->KL * h->NF + static_cast<size_t>(h->NF) * h->A + h->A +
                               static_cast<size_t>(L) * h->A) * sizeof(float);
        size_t smem_total = smem;
        a.att_off = 0;
        if (((smem + 15) & ~static_cast<size_t>(15)) + tables <= 200 * 1024) {
            a.att_off = static_cast<int>(((smem + 15) & ~static_cast<size_t>(15)) / sizeof(float));
            smem_total = a.att_off * sizeof(float) + tables;
        }
        static PerDeviceOnce configured;
        if (!configured.here()) {
            TTSB_CHECK_CUDA(cudaFuncSetAttribute(t2_decoder_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            configured.here() = true;
        }
        TTSB_REQUIRE(smem <= 200 * 1024 && H % 8 == 0, "persistent decoder: shared-memory plan");
        int per_sm = 0;
        TTSB_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, t2_decoder_persistent_kernel, 512, smem_total));
        TTSB_REQUIRE(per_sm >= 1, "persistent decoder does not fit on an SM");
        // one CTA per SM; one pass over the LSTM's H/8 eight-unit blocks when the device has that many SMs
        const int grid = std::min(num_sms(), std::max(H / 8, B));
        TTSB_CHECK_CUDA(cudaMemsetAsync(st.bar, 0, 4 * sizeof(unsigned), s));
        void* kargs[] = {&a};
        TTSB_CHECK_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(t2_decoder_persistent_kernel), dim3(grid), dim3(512),
                                                    kargs, smem_total, s));
        count_launch();
        if (h_done_step) {
            TTSB_CHECK_CUDA(cudaMemcpyAsync(h_done_step, st.done_step, 2 * sizeof(int), cudaMemcpyDeviceToHost, s));
            TTSB_CHECK_CUDA(cudaStreamSynchronize(s));
        }
        return 0;
    }
    for (int i = 0; i < n_steps; ++i) {
        const int step = step0 + i;
        const int cur = step & 1, nxt = cur ^ 1;
        const uint8_t* m0 = d_masks + (static_cast<size_t>(i) * 2 + 0) * B * P;
        const uint8_t* m1 = d_masks + (static_cast<size_t>(i) * 2 + 1) * B * P;
        t2_prenet_kernel<<<B, 256, (h->n_mel + P) * sizeof(float), s>>>(st.frame, h->pre_w0, h->pre_w1, m0, m1, h->n_mel, P, st.x);
        t2_lstm_cell_kernel<<<H / 2, 256, 0, s>>>(st.x, P, st.ctx, M, st.ah[cur], st.ah[nxt], st.ac, h->arnn_wih, h->arnn_whh,
                                                 h->arnn_b, H, B);
        t2_attention_kernel<<<B, 256, att_smem, s>>>(st.ah[nxt], st.memory, st.pmem, st.lens, h->w_query, h->loc_conv,
                                                    h->loc_dense, h->att_v, st.aw, st.awc, st.ctx,
                                                    st.align + static_cast<size_t>(step) * B * L, L, H, M, h->A, h->NF, h->KL);
        t2_lstm_cell_kernel<<<H / 2, 256, 0, s>>>(st.ah[nxt], H, st.ctx, M, st.dh[cur], st.dh[nxt], st.dc, h->drnn_wih,
                                                 h->drnn_whh, h->drnn_b, H, B);
        t2_project_kernel<<<ceil_div(h->n_mel + 1, 8), 256, 0, s>>>(st.dh[nxt], st.ctx, h->w_proj, h->b_proj, H, M, h->n_mel, B,
                                                                    st.frame, st.frames, step, max_steps, st.gate);
        t2_bookkeep_kernel<<<1, 64, 0, s>>>(st.gate, st.finished, st.mel_lens, st.done_step, B, gate_threshold, step);
        count_launch(6);
    }
    TTSB_CHECK_CUDA(cudaGetLastError());
    if (h_done_step) {
        TTSB_CHECK_CUDA(cudaMemcpyAsync(h_done_step, st.done_step, 2 * sizeof(int), cudaMemcpyDeviceToHost, s));
        TTSB_CHECK_CUDA(cudaStreamSynchronize(s));
    }
    return 0;
    });
}

int ttsb_tacotron2_postprocess(const float* d_mel, const int32_t* d_mel_lens, const float* d_align, const int32_t* d_cols,
                               double rate, int B, int n_mel, int T, int L, int T_out, float* d_out, int32_t* d_out_lens,
                               void* stream_) {
    return guarded_call([&]() -> int {
    TTSB_REQUIRE(d_mel && d_mel_lens && d_out && d_out_lens, "null argument");
    TTSB_REQUIRE(B > 0 && n_mel > 0 && T > 0 && T_out > 0, "empty batch");
    TTSB_REQUIRE(d_cols == nullptr || d_align != nullptr, "truncation needs the alignments");
    TTSB_REQUIRE(rate > 0.0, "speed rate must be positive");
    const int do_resize = rate != 1.0 ? 1 : 0;
    t2_postprocess_kernel<<<B, 256, 0, static_cast<cudaStream_t>(stream_)>>>(d_mel, d_mel_lens, d_align, d_cols, 1.0 / rate,
                                                                           do_resize, n_mel, T, L, T_out, d_out, d_out_lens);
    count_launch();
    TTSB_CHECK_CUDA(cudaGetLastError());
    return 0;
    });
}

/* After decoding T steps: postnet + residual -> d_mel [B, n_mel, T] fp32, d_mel_lengths [B] int32,
 * d_alignments [B, T, L] fp32; optional d_mel_cl [B, T, 128] fp16 (zero beyond each utterance's length). */
int ttsb_tacotron2_finish(ttsb_tacotron2_t* h, int B, int L, int max_steps, int T, float* d_mel, int32_t* d_mel_lengths,
                          float* d_alignments, void* d_mel_cl, void* d_state, void* d_workspace, size_t workspace_bytes,
                          void* stream_) {
    return guarded_call([&]() -> int {
    TTSB_REQUIRE(h && d_mel && d_mel_lengths && d_state && d_workspace, "null argument");
    TTSB_REQUIRE(T > 0 && T <= max_steps, "T out of range");
    TTSB_REQUIRE(workspace_bytes >= ttsb_tacotron2_workspace_bytes(h, B, L, T), "workspace too small");
    TTSB_DEVICE_GUARD(h->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream_);
    ConvRuntime rt;
    TTSB_PROPAGATE(get_conv_runtime(static_cast<size_t>(B) * T * 512, rt));
    T2State st = carve_t2(h, d_state, B, L, max_steps, nullptr);
    Carver c(d_workspace);
    const size_t rows = static_cast<size_t>(B) * T;
    __half* xa = c.take<__half>(rows * 512);
    __half* xb = c.take<__half>(rows * 512);
    __half* mel_in = c.take<__half>(rows * h->mel_ld);
    float* post_t = c.take<float>(static_cast<size_t>(B) * h->n_mel * T);
    t2_frames_to_cl_kernel<<<B * T, 128, 0, s>>>(st.frames, max_steps, h->n_mel, T, h->mel_ld, mel_in);
    count_launch();
    const __half* cur = mel_in;
    int ld = h->mel_ld;
    __half* bufs[2] = {xa, xb};
    for (int i = 0; i < 5; ++i) {
        EpiParams e;
        if (i < 4) {
            e.out_act = bufs[i & 1]; e.ld_act = 512; e.act_tanh = 1;
        } else {
            e.out_f32_t = post_t; e.n_store = h->n_mel;
        }
        TTSB_PROPAGATE(conv_forward(h->post[i], rt, cur, ld, B, T, e, s));
        cur = bufs[i & 1];
        ld = 512;
    }
    TTSB_CHECK_CUDA(cudaMemcpyAsync(d_mel_lengths, st.mel_lens, B * sizeof(int), cudaMemcpyDeviceToDevice, s));
    t2_finalize_kernel<<<dim3(ceil_div(T, 128), B), 128, 0, s>>>(st.frames, post_t, max_steps, h->n_mel, T, h->mel_ld, d_mel,
                                                                static_cast<__half*>(d_mel_cl), st.mel_lens);
    count_launch();
    if (d_alignments) {
        // alignments are stored [step, B, L]; the reference returns [B, T, L] (tacotron2_ms.py:330)
        for (int b = 0; b < B; ++b)
            TTSB_CHECK_CUDA(cudaMemcpy2DAsync(d_alignments + static_cast<size_t>(b) * T * L, L * sizeof(float),
                                              st.align + static_cast<size_t>(b) * L, static_cast<size_t>(B) * L * sizeof(float),
                                              L * sizeof(float), T, cudaMemcpyDeviceToDevice, s));
    }
    TTSB_CHECK_CUDA(cudaGetLastError());
    return 0;
    });
}

}  // extern "C"
