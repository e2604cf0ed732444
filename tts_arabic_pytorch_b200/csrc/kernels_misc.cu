// HBM-bound and small kernels: layout packing, embedding, attention, length regulation,
// conv_post+tanh. See kernels.cuh for the reference op site each one replaces.
#include <cstdlib>
#include <string>
#include "kernels.cuh"
#include "epilogue.cuh"

namespace ttsb {

// ------------------------------------------------------------------------------------------------
// pack_mel: [B,C,T] fp32 -> [B,T,ld] fp16 (transpose through smem, 32x32 tiles)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pack_mel_kernel(const float* __restrict__ mel,
                                                       const int* __restrict__ lens, int C, int T, int t_stride,
                                                       __half* __restrict__ out, int ld) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int len = lens ? lens[b] : T;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 8 rows per pass
    for (int r = ty; r < 32; r += 8) {
        const int c = c0 + r, t = t0 + tx;
        tile[r][tx] = (c < C && t < T && t < len) ? mel[(static_cast<size_t>(b) * C + c) * t_stride + t] : 0.f;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int t = t0 + r, c = c0 + tx;
        if (t < T && c < ld) out[(static_cast<size_t>(b) * T + t) * ld + c] = __float2half(tile[tx][r]);
    }
}

int launch_pack_mel(const float* mel, const int* lens, int B, int C, int T, __half* out, int ld,
                    cudaStream_t s, int t_stride) {
    dim3 grid(ceil_div(T, 32), ceil_div(ld, 32), B);
    pack_mel_kernel<<<grid, 256, 0, s>>>(mel, lens, C, T, t_stride > 0 ? t_stride : T, out, ld);
    count_launch();
    TTSB_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------
// conv_post + tanh: 256 samples per block. The (256+6) x 32 fp16 activations are staged as they are (80-byte row
// pitch: 16-byte reads of consecutive rows hit disjoint banks), the 7 x 32 weights as fp32; a thread reads its
// 7 rows with 4 LDS.128 each and the weights with broadcast LDS.128 — 84 shared loads for 224 FMAs (the first
// version staged fp32 and issued two 4-byte loads per FMA, which made it LSU-bound at 1.3 TB/s).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) conv_post_tanh_kernel(const __half* __restrict__ x,
                                                             const float* __restrict__ w, float bias,
                                                             const int* __restrict__ lens, int len_mul,
                                                             int N, int n_stride, float* __restrict__ wav) {
    constexpr int kPitch = 80;                       // bytes per staged row (64 B of data)
    __shared__ __align__(16) uint8_t sx[262 * kPitch];
    __shared__ __align__(16) float sw[7 * 32];
    const int b = blockIdx.y;
    const int n0 = blockIdx.x * 256;
    if (threadIdx.x < 224) sw[threadIdx.x] = w[threadIdx.x];
    // 262 rows x 32 ch = 262 x 4 uint4
    for (int i = threadIdx.x; i < 262 * 4; i += 256) {
        const int r = i >> 2, q = i & 3;
        const int n = n0 + r - 3;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (n >= 0 && n < N) v = *reinterpret_cast<const uint4*>(x + (static_cast<size_t>(b) * N + n) * 32 + q * 8);
        *reinterpret_cast<uint4*>(sx + r * kPitch + q * 16) = v;
    }
    __syncthreads();
    const int n = n0 + threadIdx.x;
    if (n >= N) return;
    float acc = bias;
#pragma unroll
    for (int k = 0; k < 7; ++k) {
        const uint8_t* row = sx + (threadIdx.x + k) * kPitch;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float f[8];
            unpack8(*reinterpret_cast<const uint4*>(row + q * 16), f);
            const float4 w0 = *reinterpret_cast<const float4*>(sw + k * 32 + q * 8);
            const float4 w1 = *reinterpret_cast<const float4*>(sw + k * 32 + q * 8 + 4);
            acc += f[0] * w0.x + f[1] * w0.y + f[2] * w0.z + f[3] * w0.w + f[4] * w1.x + f[5] * w1.y + f[6] * w1.z + f[7] * w1.w;
        }
    }
    const int len = lens ? lens[b] * len_mul : N;
    wav[static_cast<size_t>(b) * n_stride + n] = n < len ? tanhf(acc) : 0.f;
}

int launch_conv_post_tanh(const __half* x, const float* w, float bias, const int* lens, int len_mul,
                          int B, int N, float* wav, cudaStream_t s, int n_stride) {
    dim3 grid(ceil_div(N, 256), B);
    conv_post_tanh_kernel<<<grid, 256, 0, s>>>(x, w, bias, lens, len_mul, N, n_stride > 0 ? n_stride : N, wav);
    count_launch();
    TTSB_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------
// embedding + positional embedding
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float posemb_val(int pos, int j, int D, const float* inv_freq) {
    // pos_emb = cat[sin(pos*inv_freq), cos(pos*inv_freq)] (transformer.py:42-44)
    const int half = D / 2;
    const float a = static_cast<float>(pos) * inv_freq[j < half ? j : j - half];
    return j < half ? sinf(a) : cosf(a);
}

// ids / speaker ids are clamped into their tables here; out-of-range values are REPORTED by ids_to_lens_kernel through
// the status word the caller reads at its one host sync (nn.Embedding raises IndexError in the reference — a silent
// out-of-bounds gather would be garbage audio or a poisoned context)
__global__ void embed_kernel(const int64_t* __restrict__ ids, const float* __restrict__ emb,
                             const float* __restrict__ spk_table, const int64_t* __restrict__ spk_ids, int spk_scalar,
                             int n_speakers, const float* __restrict__ inv_freq,
                             int L, int D, int n_symbols, __half* __restrict__ out) {
    const int row = blockIdx.x;  // b*L + l
    const int l = row % L;
    const int64_t raw = ids[row];
    const int64_t id = raw < 0 ? 0 : (raw >= n_symbols ? n_symbols - 1 : raw);
    const bool valid = raw != 0;
    const float* cond = nullptr;
    if (spk_table != nullptr) {
        int64_t sp = spk_ids != nullptr ? spk_ids[row / L] : spk_scalar;
        sp = sp < 0 ? 0 : (sp >= n_speakers ? n_speakers - 1 : sp);
        cond = spk_table + sp * D;
    }
    for (int j = threadIdx.x; j < D; j += blockDim.x) {
        float v = emb[id * D + j];
        if (valid) v += posemb_val(l, j, D, inv_freq);
        if (cond) v += cond[j];
        out[static_cast<size_t>(row) * D + j] = __float2half(v);
    }
}

int launch_embed(const int64_t* ids, const float* emb, const float* spk_table, const int64_t* spk_ids, int spk_scalar,
                 int n_speakers, const float* inv_freq, int B, int L, int D, int n_symbols, __half* out, cudaStream_t s) {
    embed_kernel<<<B * L, 128, 0, s>>>(ids, emb, spk_table, spk_ids, spk_scalar, n_speakers, inv_freq, L, D, n_symbols, out);
    count_launch();
    TTSB_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// lens[b] = number of non-pad ids; status bits (OR-ed into *status, which the caller zeroed): 1 an id outside
// [0, n_symbols), 2 padding that is not trailing or an empty utterance, 4 a speaker id outside [0, n_speakers)
__global__ void ids_to_lens_kernel(const int64_t* __restrict__ ids, int L, int n_symbols,
                                   const int64_t* __restrict__ spk_ids, int n_speakers, int* __restrict__ lens,
                                   int* __restrict__ status) {
    const int b = blockIdx.x;
    int cnt = 0, last = -1, bad = 0;
    for (int l = threadIdx.x; l < L; l += 32) {
        const int64_t id = ids[static_cast<size_t>(b) * L + l];
        if (id != 0) { ++cnt; last = l; }
        if (id < 0 || id >= n_symbols) bad |= 1;
    }
    for (int o = 16; o > 0; o >>= 1) {
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        last = max(last, __shfl_xor_sync(0xffffffffu, last, o));
        bad |= __shfl_xor_sync(0xffffffffu, bad, o);
    }
    if (threadIdx.x == 0) {
        lens[b] = cnt;
        if (cnt == 0 || last != cnt - 1) bad |= 2;
        if (spk_ids != nullptr && (spk_ids[b] < 0 || spk_ids[b] >= n_speakers)) bad |= 4;
        if (bad && status != nullptr) atomicOr(status, bad);
    }
}
int launch_ids_to_lens(const int64_t* ids, int B, int L, int n_symbols, const int64_t* spk_ids, int n_speakers, int* lens,
                       int* status, cudaStream_t s) {
    ids_to_lens_kernel<<<B, 32, 0, s>>>(ids, L, n_symbols, spk_ids, n_speakers, lens, status);
    count_launch();
    TTSB_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------
// attention: one CTA = 64 queries of one utterance; keys/values streamed in tiles of 64 with an
// online softmax (no [S,S] score tensor in HBM, unlike transformer.py:131-141).
// 256 threads as 16x16: thread (ty,tx) owns score rows 4ty..4ty+3, cols 4tx..4tx+3 of the tile,
// and output rows 4ty.., head dims 4tx.. .
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) attention_kernel(const __half* __restrict__ qkv,
                                                        const int* __restrict__ lens, int S,
                                                        float scale, __half* __restrict__ out) {
    extern __shared__ float att_smem[];
    float (*sq)[65] = reinterpret_cast<float (*)[65]>(att_smem);
    float (*sk)[65] = sq + 64;
    float (*sv)[65] = sk + 64;
    float (*sp)[65] = sv + 64;
    const int b = blockIdx.y;
    const int q0 = blockIdx.x * 64;
    const int len = lens ? min(lens[b], S) : S;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const __half* base = qkv + static_cast<size_t>(b) * S * 192;

    for (int i = threadIdx.x; i < 64 * 8; i += 256) {
        const int r = i >> 3, c8 = i & 7;
        float f[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (q0 + r < S) load8h(base + static_cast<size_t>(q0 + r) * 192 + c8 * 8, f);
#pragma unroll
        for (int j = 0; j < 8; ++j) sq[r][c8 * 8 + j] = f[j] * scale;
    }
    float m[4], lsum[4], o[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        m[i] = -INFINITY;
        lsum[i] = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
    }
    for (int k0 = 0; k0 < len; k0 += 64) {
        __syncthreads();
        for (int i = threadIdx.x; i < 64 * 8; i += 256) {
            const int r = i >> 3, c8 = i & 7;
            float fk[8] = {0, 0, 0, 0, 0, 0, 0, 0}, fv[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            if (k0 + r < len) {
                load8h(base + static_cast<size_t>(k0 + r) * 192 + 64 + c8 * 8, fk);
                load8h(base + static_cast<size_t>(k0 + r) * 192 + 128 + c8 * 8, fv);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) { sk[r][c8 * 8 + j] = fk[j]; sv[r][c8 * 8 + j] = fv[j]; }
        }
        __syncthreads();
        float sc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) sc[i][j] = 0.f;
        for (int d = 0; d < 64; ++d) {
            float qv[4], kv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { qv[i] = sq[ty * 4 + i][d]; kv[i] = sk[tx * 4 + i][d]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) sc[i][j] += qv[i] * kv[j];
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float mx = -INFINITY;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (k0 + tx * 4 + j >= len) sc[i][j] = -INFINITY;
                mx = fmaxf(mx, sc[i][j]);
            }
            for (int off = 8; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
            const float m_new = fmaxf(m[i], mx);  // finite: every tile has >= 1 valid key
            const float corr = __expf(m[i] - m_new);
            float ps = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float p = __expf(sc[i][j] - m_new);
                sp[ty * 4 + i][tx * 4 + j] = p;
                ps += p;
            }
            for (int off = 8; off > 0; off >>= 1) ps += __shfl_xor_sync(0xffffffffu, ps, off);
            lsum[i] = lsum[i] * corr + ps;
            m[i] = m_new;
#pragma unroll
            for (int j = 0; j < 4; ++j) o[i][j] *= corr;
        }
        __syncthreads();
        for (int k = 0; k < 64; ++k) {
            float pv[4], vv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { pv[i] = sp[ty * 4 + i][k]; vv[i] = sv[k][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) o[i][j] += pv[i] * vv[j];
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int q = q0 + ty * 4 + i;
        if (q >= S) continue;
        const float inv = lsum[i] > 0.f ? 1.f / lsum[i] : 0.f;
        __half2 h01 = __floats2half2_rn(o[i][0] * inv, o[i][1] * inv);
        __half2 h23 = __floats2half2_rn(o[i][2] * inv, o[i][3] * inv);
        uint2 u;
        u.x = *reinterpret_cast<uint32_t*>(&h01);
        u.y = *reinterpret_cast<uint32_t*>(&h23);
        *reinterpret_cast<uint2*>(out + (static_cast<size_t>(b) * S + q) * 64 + tx * 4) = u;
    }
}

// ------------------------------------------------------------------------------------------------
// attention on the legacy tensor path (mma.sync m16n8k16, fp16 in / fp32 accumulate): the same online-softmax
// schedule, 64 queries per CTA as 4 warps x 16 rows, K/V tiles of 64 keys staged with a 144-byte row pitch
// (conflict-free ldmatrix). 2 % of the model's FLOPs with S <= ~1k and d_head = 64: a tcgen05 tile (M = 128,
// TMEM round trip for the softmax) does not pay here; the fp32-FMA kernel above stays as the cross-check
// (TTSB_ATTENTION=simt).
// ------------------------------------------------------------------------------------------------
constexpr int kAttPitch = 72;   // halfs per staged row (64 of data)

__device__ __forceinline__ void ldsm_x4(uint32_t saddr, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t saddr, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr));
}
__device__ __forceinline__ void mma_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(128) attention_mma_kernel(const __half* __restrict__ qkv,
                                                            const int* __restrict__ lens, int S,
                                                            float scale, __half* __restrict__ out) {
    __shared__ __align__(16) __half sq[64 * kAttPitch];
    __shared__ __align__(16) __half sk[64 * kAttPitch];
    __shared__ __align__(16) __half sv[64 * kAttPitch];
    const int b = blockIdx.y;
    const int q0 = blockIdx.x * 64;
    const int len = lens ? min(lens[b], S) : S;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const __half* base = qkv + static_cast<size_t>(b) * S * 192;

    // stage Q (rows beyond S as zeros)
    for (int i = threadIdx.x; i < 64 * 8; i += 128) {
        const int r = i >> 3, c8 = i & 7;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (q0 + r < S) v = *reinterpret_cast<const uint4*>(base + static_cast<size_t>(q0 + r) * 192 + c8 * 8);
        *reinterpret_cast<uint4*>(sq + r * kAttPitch + c8 * 8) = v;
    }
    __syncthreads();
    // Q fragments of this warp's 16 rows, all four k steps of d
    uint32_t qf[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks)
        ldsm_x4(smem_u32(sq + (warp * 16 + (lane & 15)) * kAttPitch + ks * 16 + (lane >> 4) * 8), qf[ks]);

    float o[8][4];
#pragma unroll
    for (int nb = 0; nb < 8; ++nb)
#pragma unroll
        for (int j = 0; j < 4; ++j) o[nb][j] = 0.f;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;   // rows lane/4 and lane/4 + 8
    const float sl2 = scale * 1.4426950408889634f;              // scores in log2 units

    for (int k0 = 0; k0 < len; k0 += 64) {
        __syncthreads();
        for (int i = threadIdx.x; i < 64 * 8; i += 128) {
            const int r = i >> 3, c8 = i & 7;
            uint4 vk = make_uint4(0, 0, 0, 0), vv = make_uint4(0, 0, 0, 0);
            if (k0 + r < len) {
                vk = *reinterpret_cast<const uint4*>(base + static_cast<size_t>(k0 + r) * 192 + 64 + c8 * 8);
                vv = *reinterpret_cast<const uint4*>(base + static_cast<size_t>(k0 + r) * 192 + 128 + c8 * 8);
            }
            *reinterpret_cast<uint4*>(sk + r * kAttPitch + c8 * 8) = vk;
            *reinterpret_cast<uint4*>(sv + r * kAttPitch + c8 * 8) = vv;
        }
        __syncthreads();
        // S = Q K^T: 8 key blocks of 8, 4 k steps over d
        float sc[8][4];
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) {
#pragma unroll
            for (int j = 0; j < 4; ++j) sc[nb][j] = 0.f;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                uint32_t kf[4];   // keys nb*8.., d = h*32 + {0-7, 8-15, 16-23, 24-31}
                ldsm_x4(smem_u32(sk + (nb * 8 + (lane & 7)) * kAttPitch + h * 32 + (lane >> 3) * 8), kf);
                mma_16816(sc[nb], qf[h * 2], kf[0], kf[1]);
                mma_16816(sc[nb], qf[h * 2 + 1], kf[2], kf[3]);
            }
        }
        // online softmax; thread owns cols (lane%4)*2, +1 of each key block for rows lane/4 (c0,c1) and +8 (c2,c3)
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) {
            const int key = k0 + nb * 8 + (lane & 3) * 2;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const bool ok = key + (j & 1) < len;
                sc[nb][j] = ok ? sc[nb][j] * sl2 : -INFINITY;
            }
            mx0 = fmaxf(mx0, fmaxf(sc[nb][0], sc[nb][1]));
            mx1 = fmaxf(mx1, fmaxf(sc[nb][2], sc[nb][3]));
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);   // finite: every tile has >= 1 valid key
        const float c0 = exp2f(m0 - mn0), c1 = exp2f(m1 - mn1);
        m0 = mn0; m1 = mn1;
        float ps0 = 0.f, ps1 = 0.f;
        uint32_t pf[4][4];   // P as A fragments: k step j = key blocks 2j, 2j+1
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) {
            const float p0 = exp2f(sc[nb][0] - mn0), p1 = exp2f(sc[nb][1] - mn0);
            const float p2 = exp2f(sc[nb][2] - mn1), p3 = exp2f(sc[nb][3] - mn1);
            ps0 += p0 + p1;
            ps1 += p2 + p3;
            pf[nb >> 1][(nb & 1) * 2 + 0] = pack_half2(p0, p1);
            pf[nb >> 1][(nb & 1) * 2 + 1] = pack_half2(p2, p3);
        }
        ps0 += __shfl_xor_sync(0xffffffffu, ps0, 1);
        ps0 += __shfl_xor_sync(0xffffffffu, ps0, 2);
        ps1 += __shfl_xor_sync(0xffffffffu, ps1, 1);
        ps1 += __shfl_xor_sync(0xffffffffu, ps1, 2);
        l0 = l0 * c0 + ps0;
        l1 = l1 * c1 + ps1;
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) { o[nb][0] *= c0; o[nb][1] *= c0; o[nb][2] *= c1; o[nb][3] *= c1; }
        // O += P V: 4 k steps of 16 keys, 8 d blocks (two per ldmatrix.trans)
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
            for (int n2 = 0; n2 < 4; ++n2) {
                uint32_t vf[4];   // {keys 0-7, 8-15} x d block 2*n2, then the same for d block 2*n2+1
                ldsm_x4_trans(smem_u32(sv + (ks * 16 + ((lane >> 3) & 1) * 8 + (lane & 7)) * kAttPitch + (n2 * 2 + (lane >> 4)) * 8), vf);
                mma_16816(o[n2 * 2], pf[ks], vf[0], vf[1]);
                mma_16816(o[n2 * 2 + 1], pf[ks], vf[2], vf[3]);
            }
        }
    }
    const float inv0 = l0 > 0.f ? 1.f / l0 : 0.f, inv1 = l1 > 0.f ? 1.f / l1 : 0.f;
    const int r0 = q0 + warp * 16 + (lane >> 2), r1 = r0 + 8;
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
        const int d = nb * 8 + (lane & 3) * 2;
        if (r0 < S) *reinterpret_cast<uint32_t*>(out + (static_cast<size_t>(b) * S + r0) * 64 + d) = pack_half2(o[nb][0] * inv0, o[nb][1] * inv0);
        if (r1 < S) *reinterpret_cast<uint32_t*>(out + (static_cast<size_t>(b) * S + r1) * 64 + d) = pack_half2(o[nb][2] * inv1, o[nb][3] * inv1);
    }
}

static const int kAttSmem = 4 * 64 * 65 * sizeof(float);
int launch_attention(const __half* qkv, const int* lens, int B, int S, float scale, __half* out,
                     cudaStream_t s, __half* vt_scratch, int* err_flag) {
    // TTSB_ATTENTION = tc (default: tcgen05, attention_tc.cu) | mma (mma.sync m16n8k16) | simt (fp32 FMA): the two older
    // generations stay as on-GPU cross-checks
    static const std::string mode = getenv("TTSB_ATTENTION") != nullptr ? std::string(getenv("TTSB_ATTENTION")) : std::string("tc");
    if (mode == "tc" && vt_scratch != nullptr) return launch_attention_tc(qkv, lens, B, S, scale, out, vt_scratch, err_flag, s);
    const bool use_simt = mode == "simt";
    dim3 grid(ceil_div(S, 64), B);
    if (!use_simt) {
        attention_mma_kernel<<<grid, 128, 0, s>>>(qkv, lens, S, scale, out);
        count_launch();
        TTSB_CHECK_CUDA(cudaGetLastError());
        return 0;
    }
    static PerDeviceOnce configured;
    if (!configured.here()) {
        TTSB_CHECK_CUDA(cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             kAttSmem));
        configured.here() = true;
    }
    attention_kernel<<<grid, 256, kAttSmem, s>>>(qkv, lens, S, scale, out);
    count_launch();
    TTSB_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------
// pitch / energy embedding add
// ------------------------------------------------------------------------------------------------
__global__ void scalar_embed_add_kernel(__half* __restrict__ x, const float* __restrict__ p,
                                        const float* __restrict__ w, const float* __restrict__ bias,
                                        const int* __restrict__ lens, int L, int D, int apply_mask) {
    const int row = blockIdx.x;
    const int b = row / L, l = row % L;
    const float pm = l > 0 ? p[row - 1] : 0.f;
    const float p0 = p[row];
    const float pp = l + 1 < L ? p[row + 1] : 0.f;
    const bool keep = !apply_mask || !lens || l < lens[b];
    for (int j = threadIdx.x; j < D; j += blockDim.x) {
        const size_t idx = static_cast<size_t>(row) * D + j;
        float v = __half2float(x[idx]) + bias[j] + w[j * 3 + 0] * pm + w[j * 3 + 1] * p0 + w[j * 3 + 2] * pp;
        x[idx] = __float2half(keep ? v : 0.f);
    }
}
int launch_scalar_embed_add(__half* x, const float* p, const float* w, const float* bias,
                            const int* lens, int B, int L, int D, int apply_mask, cudaStream_t s) {
    scalar_embed_add_kernel<<<B * L, 128, 0, s>>>(x, p, w, bias, lens, L, D, apply_mask);
    count_launch();
    TTSB_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------
// durations -> repeats -> cumulative offsets (one warp per utterance; L is a few hundred)
// ------------------------------------------------------------------------------------------------
__global__ void durations_kernel(const float* __restrict__ log_dur, const float* __restrict__ dur_tgt,
                                 float pace, float max_duration, int L, float* __restrict__ dur_pred,
                                 int* __restrict__ cum, int* __restrict__ dec_lens,
                                 int64_t* __restrict__ dec_lens64, int* __restrict__ max_dec_len) {
    const int b = blockIdx.x;
    const int lane = threadIdx.x;
    int running = 0;
    if (lane == 0) cum[static_cast<size_t>(b) * (L + 1)] = 0;
    for (int l0 = 0; l0 < L; l0 += 32) {
        const int l = l0 + lane;
        int rep = 0;
        if (l < L) {
            const size_t idx = static_cast<size_t>(b) * L + l;
            float d = 0.f;
            if (log_dur) {
                // dur_pred = clamp(exp(log_dur) - 1, 0, max_duration)   (model.py:368)
                d = fminf(fmaxf(expf(log_dur[idx]) - 1.f, 0.f), max_duration);
                if (dur_pred) dur_pred[idx] = d;
            }
            if (dur_tgt) d = dur_tgt[idx];
            // reps = (dur / pace + 0.5).long()   (model.py:72-73; .long() truncates toward zero)
            rep = static_cast<int>(d / pace + 0.5f);
        }
        int incl = rep;
        for (int o = 1; o < 32; o <<= 1) {
            const int n = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += n;
        }
        if (l < L) cum[static_cast<size_t>(b) * (L + 1) + l + 1] = running + incl;
        running += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) {
        dec_lens[b] = running;
        if (dec_lens64) dec_lens64[b] = running;
        if (max_dec_len) atomicMax(max_dec_len, running);
    }
}
int launch_durations(const float* log_dur, const float* dur_tgt, float pace, float max_duration,
                     int B, int L, float* dur_pred, int* cum, int* dec_lens, int64_t* dec_lens64,
                     int* max_dec_len, cudaStream_t s) {
    durations_kernel<<<B, 32, 0, s>>>(log_dur, dur_tgt, pace, max_duration, L, dur_pred, cum, dec_lens,
                                      dec_lens64, max_dec_len);
    count_launch();
    TTSB_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------
// regulate: frame -> token gather (+ decoder positional embedding). One warp per frame row.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) regulate_kernel(const __half* __restrict__ enc,
                                                       const int* __restrict__ cum,
                                                       const int* __restrict__ dec_lens,
                                                       const float* __restrict__ inv_freq, int L, int T,
                                                       int D, __half* __restrict__ out) {
    const int b = blockIdx.y;
    const int t = blockIdx.x * 4 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (t >= T) return;
    __half* orow = out + (static_cast<size_t>(b) * T + t) * D;
    if (t >= dec_lens[b]) {
        for (int j = lane * 8; j < D; j += 256) *reinterpret_cast<uint4*>(orow + j) = make_uint4(0, 0, 0, 0);
        return;
    }
    // token i with cum[i] <= t < cum[i+1]  (model.py:82-83)
    const int* c = cum + static_cast<size_t>(b) * (L + 1);
    int lo = 0, hi = L;  // invariant: c[lo] <= t < c[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (c[mid] <= t) lo = mid; else hi = mid;
    }
    const __half* erow = enc + (static_cast<size_t>(b) * L + lo) * D;
    for (int j = lane * 8; j < D; j += 256) {
        float f[8];
        load8h(erow + j, f);
#pragma unroll
        for (int q = 0; q < 8; ++q) f[q] += posemb_val(t, j + q, D, inv_freq);
        store8h(orow + j, f);
    }
}
int launch_regulate(const __half* enc, const int* cum, const int* dec_lens, const float* inv_freq,
                       int B, int L, int T, int D, __half* out, cudaStream_t s) {
    dim3 grid(ceil_div(T, 4), B);
    regulate_kernel<<<grid, 128, 0, s>>>(enc, cum, dec_lens, inv_freq, L, T, D, out);
    count_launch();
    TTSB_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace ttsb
