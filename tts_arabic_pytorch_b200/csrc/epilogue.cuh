// Row-wise fused epilogue shared by the tcgen05 conv kernel (accumulators in TMEM) and the SIMT
// check kernel (accumulators in a global fp32 scratch). One thread owns one output row
// (one time position of one utterance) and walks its N columns in groups of 32.
//
// What it fuses, per reference op site:
//   bias                       nn.Conv1d / nn.Linear bias
//   residual + LayerNorm       transformer.py:88,156 (post-LN), eps 1e-5
//   relu -> LayerNorm          model.py:54-57 (ConvReLUNorm)
//   head dot                   model.py:132 (TemporalPredictor.fc, 256 -> 1)
//   row mask                   transformer.py:174,176 (`output *= mask`), model.py:132
//   leaky-relu copy            hifigan/models.py:48,50,114 (activation of the *next* conv's input)
//   MRF mean                   hifigan/models.py:116-122 (xs / num_kernels)
//   transposed fp32 store      model.py:406-408 (proj + permute -> [B,80,T])
#pragma once
#include "common.cuh"

namespace ttsb {

enum MrfMode : int { MRF_NONE = 0, MRF_FIRST = 1, MRF_ADD = 2, MRF_LAST = 3 };

struct EpiParams {
    int T = 0;                      // rows per utterance in every row-indexed buffer below
    int n_total = 0;                // logical GEMM width (all N tiles)
    const int* lens = nullptr;      // [B] valid units per utterance, or null = no masking
    int len_mul = 1;                // valid rows = lens[b] * len_mul
    const float* bias = nullptr;    // [n_total]
    const __half* residual = nullptr;
    int ld_res = 0;
    // residual stored ACTIVATED: the buffer holds lrelu(x, s) and the epilogue adds x = min(r, r * res_inv) with
    // res_inv = 1 / s (1 = the buffer holds x itself). lrelu is invertible and both forms carry the same fp16 relative
    // error, so a tensor that feeds a conv as lrelu(x) AND a later residual add as x (every ResBlock1 input,
    // hifigan/models.py:46-53) is kept in HBM once, activated, instead of twice.
    float res_inv = 1.f;
    int pre_ln_relu = 0;
    const float* ln_g = nullptr;    // LayerNorm over n_total columns (needs a single N tile)
    const float* ln_b = nullptr;
    float ln_eps = 1e-5f;
    const float* head_w = nullptr;  // [n_total]; scalar head on the normalised row
    float head_b = 0.f;
    float* head_out = nullptr;      // [B*T]
    int mrf_mode = MRF_NONE;
    __half* mrf_buf = nullptr;      // [B*T, n_total]
    float mrf_scale = 1.f;
    __half* out_raw = nullptr;      // v
    int ld_raw = 0;
    __half* out_act = nullptr;      // lrelu(v, act_slope); slope 0 == relu
    int ld_act = 0;
    float act_slope = 0.f;
    int act_tanh = 0;               // out_act = tanh(v) instead of leaky-relu (Tacotron2 postnet)
    float* out_f32 = nullptr;       // v as fp32, row-major [B*T, ld_f32] (LSTM gate pre-activations)
    int ld_f32 = 0;
    float* out_f32_t = nullptr;     // [B, n_store, T] fp32, transposed store
    int n_store = 0;
    int f32_unmasked = 0;           // out_f32_t receives the value before row masking
    int debug = 0;                  // timing decomposition (TTSB_EPI_DEBUG; results are WRONG when set): 1 no wait for the
                                    // staging tile's previous TMA store, 2 no TMA store, 4 no proxy fence, 8 no epilogue math
};

#ifdef __CUDACC__
struct TmemAcc {
    uint32_t taddr;  // lane base already folded in
    __device__ __forceinline__ void load(int c0, float (&v)[32]) const { tmem_ld32(taddr + c0, v); }
    __device__ __forceinline__ void store(int c0, const float (&v)[32]) const { tmem_st32(taddr + c0, v); }
};
struct GmemAcc {
    float* row;  // this thread's row of the fp32 scratch, or null when the row is out of range
    __device__ __forceinline__ void load(int c0, float (&v)[32]) const {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = row ? row[c0 + j] : 0.f;
    }
    __device__ __forceinline__ void store(int c0, const float (&v)[32]) const {
        if (!row) return;
#pragma unroll
        for (int j = 0; j < 32; ++j) row[c0 + j] = v[j];
    }
};

__device__ __forceinline__ void load8h(const __half* p, float (&f)[8]) {
    uint4 u = *reinterpret_cast<const uint4*>(p);
    float2 a = unpack_half2(u.x), b = unpack_half2(u.y), c = unpack_half2(u.z), d = unpack_half2(u.w);
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ void store8h(__half* p, const float* f) {
    uint4 u;
    u.x = pack_half2(f[0], f[1]); u.y = pack_half2(f[2], f[3]);
    u.z = pack_half2(f[4], f[5]); u.w = pack_half2(f[6], f[7]);
    *reinterpret_cast<uint4*>(p) = u;
}

// ------------------------------------------------------------------------------------------------
// Row I/O of one warp: 32 rows x 32 fp16 columns (64 B per row) per step.
//
// In the accumulator layout lane == row, so a naive 16-byte global access per lane touches 32
// different 128-byte lines per instruction; the v2 timelines (profiles/r01_s11_timeline_v2_elect.txt)
// show the epilogue bound by exactly that line-request rate (5.7k cycles to retire the stores of a
// 128x32 tile). With a per-warp 2 KB staging tile the global side is issued "transposed": lane l of
// step i moves 16-byte unit (i*32 + l) of the row-major block -> 8 rows x 64 B per instruction
// (4 full lines when rows are contiguous, C = 32; 8 half lines otherwise) instead of 32 lines.
// The tile is XOR-swizzled so both the row-owner and the transposed accesses are conflict-free.
// ------------------------------------------------------------------------------------------------
struct Chunk32 {
    uint4 q[4];
};
__device__ __forceinline__ uint32_t wtile_off(int r, int u) { return r * 64 + ((u ^ ((r >> 1) & 3)) << 4); }

// explicit shared-space accesses: through a generic pointer the compiler emits generic LD/ST (ST.E.128)
// for the staging tile, which costs an address-space lookup per access
__device__ __forceinline__ void sts128(uint32_t saddr, const uint4& v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t saddr) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr) : "memory");
    return v;
}

// timing-decomposition switches (EpiParams::debug) are compiled in only with -DTTSB_EPI_DEBUG: the epilogues sit at their
// register budgets and even a dead runtime branch pushes them into spilling
#ifdef TTSB_EPI_DEBUG
constexpr bool kEpiDebug = true;
#else
constexpr bool kEpiDebug = false;
#endif

struct RowIO {
    uint8_t* tile;      // this warp's 2 KB staging tile, or null -> direct per-row access
    int lane;
    int rows_valid;     // rows of this warp inside [0, T): clamp(T - warp_row0, 0, 32)
    int dbg = 0;        // EpiParams::debug

    // A TMA store may still be reading the tile: the lane that issued it (elect.sync = lowest lane) waits for the
    // read to finish before anyone overwrites the tile. Cheap when nothing is pending.
    __device__ __forceinline__ void release_tile() const {
        if (elect_one()) tma_store_wait_read0();
        __syncwarp();
    }
    // write this lane's own row (4 units) of a [32 x 32] block through ONE TMA store: the tile's XOR pattern is the
    // 64-byte TMA swizzle, so the row-owner layout goes out as it is — no LDS / STG / address arithmetic per lane.
    // (c_col, c_row, c_b) = tensor coordinates of the block; rows beyond the tensor are clipped by the TMA unit.
    __device__ __forceinline__ void store_tma(const CUtensorMap* tm, int c_col, int c_row, int c_b, const Chunk32& own,
                                              long long* dbg_released = nullptr) const {
        const uint32_t base = smem_u32(tile);
        if (!(kEpiDebug && (dbg & 1))) release_tile();
        if (dbg_released) *dbg_released = clock64();
#pragma unroll
        for (int u = 0; u < 4; ++u) sts128(base + wtile_off(lane, u), own.q[u]);
        if (!(kEpiDebug && (dbg & 4))) fence_proxy_async();
        __syncwarp();
        if (rows_valid > 0 && !(kEpiDebug && (dbg & 2)) && elect_one()) {
            tma_store_3d(tm, tile, c_col, c_row, c_b);
            tma_store_commit();
        }
    }

    // request a [32 x 32] block whose (row 0, col 0) element is at `blk` (row pitch ld elements)
    __device__ __forceinline__ void request(const __half* blk, long ld, bool on, Chunk32& c) const {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int r, u;
            if (tile) { const int id = i * 32 + lane; r = id >> 2; u = id & 3; } else { r = lane; u = i; }
            c.q[i] = (on && r < rows_valid) ? *reinterpret_cast<const uint4*>(blk + r * ld + u * 8) : make_uint4(0, 0, 0, 0);
        }
    }
    // turn a requested block into this lane's own row (4 units of 8 columns)
    __device__ __forceinline__ void to_row(Chunk32& c, bool tma_in_use = false) const {
        if (!tile) return;
        const uint32_t base = smem_u32(tile);
        if (tma_in_use) release_tile();
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int id = i * 32 + lane;
            sts128(base + wtile_off(id >> 2, id & 3), c.q[i]);
        }
        __syncwarp();
#pragma unroll
        for (int u = 0; u < 4; ++u) c.q[u] = lds128(base + wtile_off(lane, u));
    }
    // write this lane's own row (4 units) of a [32 x 32] block
    __device__ __forceinline__ void store(__half* blk, long ld, const Chunk32& own, bool tma_in_use = false) const {
        if (tile && tma_in_use) release_tile();
        if (!tile) {
            if (lane < rows_valid) {
#pragma unroll
                for (int u = 0; u < 4; ++u) *reinterpret_cast<uint4*>(blk + lane * ld + u * 8) = own.q[u];
            }
            return;
        }
        const uint32_t base = smem_u32(tile);
        __syncwarp();
#pragma unroll
        for (int u = 0; u < 4; ++u) sts128(base + wtile_off(lane, u), own.q[u]);
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int id = i * 32 + lane;
            const int r = id >> 2, u = id & 3;
            const uint4 v = lds128(base + wtile_off(r, u));
            if (r < rows_valid) *reinterpret_cast<uint4*>(blk + r * ld + u * 8) = v;
        }
    }
};

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
    float2 a = unpack_half2(u.x), b = unpack_half2(u.y), c = unpack_half2(u.z), d = unpack_half2(u.w);
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ uint4 pack8(const float* f) {
    uint4 u;
    u.x = pack_half2(f[0], f[1]); u.y = pack_half2(f[2], f[3]);
    u.z = pack_half2(f[4], f[5]); u.w = pack_half2(f[6], f[7]);
    return u;
}
__device__ __forceinline__ void bias8(const float* bias, int n, float (&f)[8]) {
    if (bias == nullptr) {
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = 0.f;
        return;
    }
    const float4 a = __ldg(reinterpret_cast<const float4*>(bias + n));
    const float4 b = __ldg(reinterpret_cast<const float4*>(bias + n) + 1);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

// same from shared memory (a broadcast read): with ~all of the SM's L1 carved out as shared memory the __ldg path
// above misses to L2 — one ~700-cycle round trip per 8 columns, on the epilogue's critical path
// volatile form: stays where it is written relative to the other volatile asm statements (tcgen05.ld, st.shared), so a
// chunk's bias values can be requested BEFORE the accumulator load is waited for instead of right before their use
// (each LDS -> FADD pair there is a ~30-cycle stall of a warp that has no other work; round-2 SASS review)
__device__ __forceinline__ void bias8s_early(uint32_t saddr, float* f) {
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(f[0]), "=f"(f[1]), "=f"(f[2]), "=f"(f[3]) : "r"(saddr));
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(f[4]), "=f"(f[5]), "=f"(f[6]), "=f"(f[7]) : "r"(saddr + 16));
}
__device__ __forceinline__ void bias8s(uint32_t saddr, float (&f)[8]) {
    // not volatile: the bias tile is written once before the role dispatch, so the compiler may hoist / batch these
    asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(f[0]), "=f"(f[1]), "=f"(f[2]), "=f"(f[3]) : "r"(saddr));
    asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(f[4]), "=f"(f[5]), "=f"(f[6]), "=f"(f[7]) : "r"(saddr + 16));
}

// b: utterance, t: row inside the utterance (lane i of the warp owns row t = warp_row0 + i), row_ok:
// t < T (loads from the accumulator are warp-collective, so out-of-range threads still walk the loop
// but never touch memory). n_base: first logical column of this N tile; n_tile: its width (multiple
// of 32). `wait_acc()` blocks until the accumulator is complete; it is called AFTER the first chunk
// of the residual / MRF rows has been requested from HBM. `acc_drained()` is called right after the
// last read of the accumulator so a persistent kernel can hand the TMEM buffer back early.
// `stage`: this warp's 2 KB staging tile for coalesced row I/O (null = direct access).
// kLnMode: -1 = every feature decided at run time (the check kernels), 1 = LayerNorm launches only (residual + LN,
// relu -> LN, predictor head: no fp32 / transposed / tanh outputs), 0 = launches without LayerNorm. The two compile-time
// forms exist because the all-in-one body needs 254 registers AND a stack frame in conv_tc2 (every spill reload is an L2
// round trip on an SM whose L1 is carved out as shared memory).
template <int kLnMode = -1, class Acc, class WaitFn, class DrainFn>
__device__ __forceinline__ void run_epilogue(const EpiParams& e, const Acc& acc, int b, int t,
                                             bool row_ok, int n_base, int n_tile, WaitFn wait_acc,
                                             DrainFn acc_drained, uint8_t* stage = nullptr,
                                             long long* dbg = nullptr, uint32_t params_saddr = 0) {
    // params_saddr != 0: shared-memory copy of this N tile's per-column parameters, n_tile floats each:
    // [bias][ln_g][ln_b][head_w] (see bias8s for why)
    const int lane = threadIdx.x & 31;
    const int warp_row0 = t - lane;
    const long row0 = static_cast<long>(b) * e.T + warp_row0;      // first row of this warp
    const long row = row0 + lane;
    bool in_len = true;
    if (e.lens != nullptr && row_ok) in_len = t < e.lens[b] * e.len_mul;
    const bool do_ln = kLnMode < 0 ? e.ln_g != nullptr : kLnMode == 1;
    constexpr bool kNoLnOutputs = kLnMode == 1;     // LayerNorm launches write raw / activated fp16 rows and the head only
    RowIO io{stage, lane, min(32, max(0, e.T - warp_row0))};
    const bool use_res = e.residual != nullptr;
    const bool use_mrf = !kNoLnOutputs && (e.mrf_mode == MRF_ADD || e.mrf_mode == MRF_LAST);
    const __half* res_blk = e.residual + row0 * e.ld_res + n_base;
    __half* mrf_blk = e.mrf_buf + row0 * e.n_total + n_base;

    // which = 0 bias, 1 ln_g, 2 ln_b, 3 head_w; n = global column
    auto param8 = [&](int which, const float* gptr, int n, float (&f)[8]) {
        if (params_saddr != 0 && gptr != nullptr) bias8s(params_saddr + (which * n_tile + (n - n_base)) * 4, f);
        else bias8(gptr, n, f);
    };
    Chunk32 res_cur, mrf_cur;
    io.request(res_blk, e.ld_res, use_res, res_cur);
    io.request(mrf_blk, e.n_total, use_mrf, mrf_cur);
    wait_acc();
    if (dbg) dbg[0] = clock64();

    float mean = 0.f, rstd = 1.f;
    if (do_ln) {
        // pass 1: materialise z = acc + bias + residual (+relu), keep it in the accumulator store
        float s1 = 0.f, s2 = 0.f;
        for (int c0 = 0; c0 < n_tile; c0 += 32) {
            float v[32];
            Chunk32 res_nxt;
            io.request(res_blk + c0 + 32, e.ld_res, use_res && c0 + 32 < n_tile, res_nxt);
            if (use_res) io.to_row(res_cur);
            __syncwarp();
            acc.load(c0, v);
            if (row_ok) {
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    float r[8], bs[8];
                    unpack8(res_cur.q[g], r);
                    param8(0, e.bias, n_base + c0 + g * 8, bs);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float x = v[g * 8 + j] + bs[j] + r[j];
                        if (e.pre_ln_relu) x = fmaxf(x, 0.f);
                        v[g * 8 + j] = x;
                        s1 += x;
                        s2 += x * x;
                    }
                }
            }
            __syncwarp();
            acc.store(c0, v);
            res_cur = res_nxt;
        }
        mean = s1 / static_cast<float>(n_tile);
        float var = s2 / static_cast<float>(n_tile) - mean * mean;
        rstd = rsqrtf(fmaxf(var, 0.f) + e.ln_eps);
    }

    float head = 0.f;
    for (int c0 = 0; c0 < n_tile; c0 += 32) {
        float v[32];
        Chunk32 res_nxt, mrf_nxt;
        const bool more = c0 + 32 < n_tile;
        io.request(res_blk + c0 + 32, e.ld_res, use_res && !do_ln && more, res_nxt);
        io.request(mrf_blk + c0 + 32, e.n_total, use_mrf && more, mrf_nxt);
        if (use_res && !do_ln) io.to_row(res_cur);
        if (use_mrf) io.to_row(mrf_cur);
        if (dbg && c0 == 0) dbg[1] = clock64();
        __syncwarp();
        acc.load(c0, v);
        if (dbg && c0 == 0) dbg[2] = clock64();
        if (!more) acc_drained();
        if (dbg && c0 == 0) dbg[3] = clock64();
        Chunk32 o_raw, o_act, o_mrf;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const int n = n_base + c0 + g * 8;
            float x[8];
            if (do_ln) {
                float gm[8], bt[8];
                param8(1, e.ln_g, n, gm);
                param8(2, e.ln_b, n, bt);
#pragma unroll
                for (int j = 0; j < 8; ++j) x[j] = (v[g * 8 + j] - mean) * rstd * gm[j] + bt[j];
                if (kLnMode != 0 && e.head_w) {
                    float hw[8];
                    param8(3, e.head_w, n, hw);
#pragma unroll
                    for (int j = 0; j < 8; ++j) head += x[j] * hw[j];
                }
            } else {
                float r[8], bs[8];
                unpack8(res_cur.q[g], r);
                param8(0, e.bias, n, bs);
#pragma unroll
                for (int j = 0; j < 8; ++j) x[j] = v[g * 8 + j] + bs[j] + fminf(r[j], r[j] * e.res_inv);
            }
            if (!kNoLnOutputs && e.out_f32_t && e.f32_unmasked && row_ok) {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (n + j < e.n_store)
                        e.out_f32_t[(static_cast<long>(b) * e.n_store + n + j) * e.T + t] = x[j];
            }
            if (!in_len) {
#pragma unroll
                for (int j = 0; j < 8; ++j) x[j] = 0.f;
            }
            if (!kNoLnOutputs && e.mrf_mode != MRF_NONE) {
                float m[8];
                unpack8(mrf_cur.q[g], m);
#pragma unroll
                for (int j = 0; j < 8; ++j) x[j] = m[j] + x[j] * e.mrf_scale;
                o_mrf.q[g] = pack8(x);
            }
            o_raw.q[g] = pack8(x);
            if (!kNoLnOutputs && e.out_f32_t && !e.f32_unmasked && row_ok) {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (n + j < e.n_store)
                        e.out_f32_t[(static_cast<long>(b) * e.n_store + n + j) * e.T + t] = x[j];
            }
            if (!kNoLnOutputs && e.out_f32 && row_ok) {
                float4* o = reinterpret_cast<float4*>(e.out_f32 + row * e.ld_f32 + n);
                o[0] = make_float4(x[0], x[1], x[2], x[3]);
                o[1] = make_float4(x[4], x[5], x[6], x[7]);
            }
            float a[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] = (!kNoLnOutputs && e.act_tanh) ? tanhf(x[j]) : lrelu(x[j], e.act_slope);
            o_act.q[g] = pack8(a);
        }
        if (dbg && c0 == 0) dbg[4] = clock64();
        if (!kNoLnOutputs && e.mrf_mode != MRF_NONE && e.mrf_mode != MRF_LAST) {
            io.store(mrf_blk + c0, e.n_total, o_mrf);
        } else {
            if (e.out_raw) io.store(e.out_raw + row0 * e.ld_raw + n_base + c0, e.ld_raw, o_raw);
            if (dbg && c0 == 0) dbg[5] = clock64();
            if (e.out_act) io.store(e.out_act + row0 * e.ld_act + n_base + c0, e.ld_act, o_act);
        }
        if (dbg && c0 == 0) dbg[6] = clock64();
        res_cur = res_nxt;
        mrf_cur = mrf_nxt;
    }
    if (kLnMode != 0 && e.head_out && row_ok) e.head_out[row] = in_len ? head + e.head_b : 0.f;
}
// ------------------------------------------------------------------------------------------------
// Lean epilogue: the vocoder's hot subset only (bias, residual, row mask, MRF accumulate, raw and
// leaky-relu outputs). The general epilogue above costs ~1000 executed instructions per 32-column chunk
// because every optional feature is a runtime branch; with ONE epilogue warp per SM sub-partition that
// instruction stream — not memory — bounded the tile period (profiles/r01_s13_epilogue_detail.txt:
// 3400 cycles of "math" per chunk). Here the feature set is fixed at compile time.
// `pre` carries the first residual / MRF chunk requested while the previous tile was being finished.
// ------------------------------------------------------------------------------------------------
template <bool kMrf>
struct LeanPrefetch {
    Chunk32 res;
    int len_rows;  // valid rows of the tile's utterance (lens[b] * len_mul), fetched a tile ahead: it is an L2 round trip
};
// Only the residual rows travel a tile ahead. The MRF rows are requested at the start of their own tile: a second
// prefetched chunk pair pushed the MRF variants over the register budget, and with the L1 carved out as shared
// memory every spill reload is an L2 round trip — the spilling variants ran at HALF the speed of the others
// (profiles/r01_s41_launches_b64.csv).
template <bool kMrf>
__device__ __forceinline__ void lean_prefetch(const EpiParams& e, const RowIO& io, long row0, int n_base, bool on,
                                              LeanPrefetch<kMrf>& p, int b) {
    p.len_rows = (on && e.lens != nullptr) ? __ldg(e.lens + b) * e.len_mul : 0x7fffffff;
    io.request(e.residual + row0 * e.ld_res + n_base, e.ld_res, on && e.residual != nullptr, p.res);
}

// kMrf = false: mrf_mode == MRF_NONE is guaranteed by the caller.
// kLookahead = false (register-starved variants): a chunk's residual / MRF rows are requested when the chunk
// starts, not one chunk ahead.
template <bool kMrf, bool kSmemBias = false, bool kLookahead = true, bool kEarlyBias = true, bool kSmemRes = false,
          bool kActOnly = false, class Acc, class WaitFn, class DrainFn>
__device__ __forceinline__ void run_epilogue_lean(const EpiParams& e, const Acc& acc, int b, int t, int n_base,
                                                  int n_tile, WaitFn wait_acc, DrainFn acc_drained, uint8_t* stage,
                                                  const LeanPrefetch<kMrf>& pre, int t_end = 0x7fffffff,
                                                  uint32_t bias_saddr = 0, const CUtensorMap* tm_raw = nullptr,
                                                  const CUtensorMap* tm_act = nullptr, const CUtensorMap* tm_mrf = nullptr,
                                                  uint8_t* stage_in = nullptr, long long* dbg = nullptr,
                                                  uint32_t res_saddr = 0, uint32_t res_phase = 0) {
    // kSmemRes: the residual rows of this tile are still in shared memory (conv_pair: the x panel that fed conv1, in the
    // UMMA K-major swizzled layout). res_saddr = shared address of this lane's row, res_phase = its 16-byte-unit XOR
    // pattern; 16-byte unit u of the row sits at res_saddr + ((u ^ res_phase) << 4). No global loads, no transposes and
    // none of the 32 prefetch registers of the global path.
    // dbg (timeline tools): clock64() after 0 accumulator seen, then for a steady-state chunk (the second one, or the only
    // one) 1 chunk starts, 2 residual rows in registers and accumulator chunk loaded, 3 math + pack done, 4 staging tile released by the
    // previous TMA store (TMA outputs only), 5 stores issued; 6 after the last chunk
    // stage_in: a second 2 KB tile of this warp for the residual / MRF row transposes. With ONE tile every transpose has
    // to wait until the TMA store issued just before it (the previous chunk's output) has finished reading the tile —
    // a full TMA read latency per chunk on the epilogue's critical path (round-2 timeline: ~2000 cycles per chunk).
    // With two tiles the output tile is next written a whole chunk after its store was issued. null = share `stage`.
    // tm_*: when non-null the matching output leaves through TMA stores of 32 x 32 blocks (RowIO::store_tma)
    // kSmemBias: bias_saddr is the shared-memory address of this N tile's bias floats (else e.bias through __ldg)
    // t_end: exclusive row limit of this tile's stores (conv_pair tiles own fewer than 128 rows)
    const int lane = threadIdx.x & 31;
    const int warp_row0 = t - lane;
    const long row0 = static_cast<long>(b) * e.T + warp_row0;
    const bool in_len = t < pre.len_rows;
    RowIO io{stage, lane, min(32, max(0, min(e.T, t_end) - warp_row0)), e.debug};
    RowIO io_in{stage_in != nullptr ? stage_in : stage, lane, io.rows_valid, e.debug};
    const bool shared_tile = stage_in == nullptr || stage_in == stage;
    const bool use_res = !kSmemRes && e.residual != nullptr;
    const bool use_mrf = kMrf && (e.mrf_mode == MRF_ADD || e.mrf_mode == MRF_LAST);
    const bool mrf_store = kMrf && (e.mrf_mode == MRF_FIRST || e.mrf_mode == MRF_ADD);
    const __half* res_blk = e.residual + row0 * e.ld_res + n_base;
    __half* mrf_blk = e.mrf_buf + row0 * e.n_total + n_base;
    const float mscale = kMrf ? e.mrf_scale : 1.f;
    const float slope = e.act_slope;
    const float rinv = e.res_inv;

    Chunk32 res_cur = pre.res, mrf_cur;
    if (kMrf) io.request(mrf_blk, e.n_total, use_mrf, mrf_cur);
    wait_acc();
    if (dbg) dbg[0] = clock64();
    const int dbg_c0 = n_tile > 32 ? 32 : 0;
    // warp-uniform shortcuts: tiles without masked rows skip the selects, layers without an activated copy
    // (conv_pair steps, MRF accumulation) skip its math
    const bool any_masked = __any_sync(0xffffffffu, !in_len);
    // kActOnly (caller guarantees: no MRF, no raw output, an activated output): the output selection is compile-time
    const bool want_act = kActOnly || (e.out_act != nullptr && !mrf_store);
    for (int c0 = 0; c0 < n_tile; c0 += 32) {
        float v[32];
        Chunk32 res_nxt, mrf_nxt;
        const bool more = c0 + 32 < n_tile;
        const bool tma_out = tm_raw != nullptr || tm_act != nullptr || tm_mrf != nullptr;
        if (dbg && c0 == dbg_c0) dbg[1] = clock64();
        if (!kLookahead && c0 > 0) {
            io.request(res_blk + c0, e.ld_res, use_res, res_cur);
            if (kMrf) io.request(mrf_blk + c0, e.n_total, use_mrf, mrf_cur);
        }
        // The rows requested one chunk (or one tile) ago are consumed BEFORE the next request is issued: ptxas tracks both
        // groups of loads on one scoreboard, and a request issued first made the transposes below wait for the loads that
        // had just left (DEPBAR.LE SB0, 0 — a full L2 round trip per chunk on the critical path; round-2 timelines)
        // (register-tight callers, kEarlyBias = false, keep the request first: the other order spills there)
        if (kLookahead && !kEarlyBias) {
            io.request(res_blk + c0 + 32, e.ld_res, use_res && more, res_nxt);
            if (kMrf) io.request(mrf_blk + c0 + 32, e.n_total, use_mrf && more, mrf_nxt);
        }
        if (use_res) io_in.to_row(res_cur, tma_out && shared_tile);
        if (kMrf && use_mrf) io_in.to_row(mrf_cur, tma_out && shared_tile);
        if (kLookahead && kEarlyBias) {
            io.request(res_blk + c0 + 32, e.ld_res, use_res && more, res_nxt);
            if (kMrf) io.request(mrf_blk + c0 + 32, e.n_total, use_mrf && more, mrf_nxt);
        }
        // bias values travel one 8-column group ahead of their use (bias8s_early): group 0 before the accumulator load is
        // waited for, group g + 1 while group g is computed — 8 registers more than loading at the point of use
        // (kEarlyBias = false where that pushes a 128-register kernel into spilling: conv_pair)
        float bs_nxt[8];
        if (kSmemBias && kEarlyBias) bias8s_early(bias_saddr + c0 * 4, bs_nxt);
        __syncwarp();
        acc.load(c0, v);
        if (dbg && c0 == dbg_c0) dbg[2] = clock64();
        if (!more) acc_drained();
        Chunk32 o_raw, o_act;
        const bool want_raw = !kActOnly && (mrf_store || e.out_raw != nullptr);
        if (kEpiDebug && (e.debug & 8)) {
#pragma unroll
            for (int g = 0; g < 4; ++g) o_raw.q[g] = o_act.q[g] = make_uint4(__float_as_uint(v[g]), 0, 0, 0);
        } else
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            float r[8], m[8], bs[8], x[8];
            if (kSmemRes) unpack8(lds128(res_saddr + ((static_cast<uint32_t>((c0 >> 3) + g) ^ res_phase) << 4)), r);
            else unpack8(res_cur.q[g], r);
            if (kMrf) unpack8(mrf_cur.q[g], m);
            if (kSmemBias && kEarlyBias) {
#pragma unroll
                for (int j = 0; j < 8; ++j) bs[j] = bs_nxt[j];
                if (g < 3) bias8s_early(bias_saddr + (c0 + (g + 1) * 8) * 4, bs_nxt);
            } else if (kSmemBias) {
                bias8s(bias_saddr + (c0 + g * 8) * 4, bs);
            } else {
                bias8(e.bias, n_base + c0 + g * 8, bs);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) x[j] = v[g * 8 + j] + bs[j] + fminf(r[j], r[j] * rinv);   // rinv >= 1: inverse lrelu
            if (any_masked) {
#pragma unroll
                for (int j = 0; j < 8; ++j) x[j] = in_len ? x[j] : 0.f;
            }
            if (kMrf) {
#pragma unroll
                for (int j = 0; j < 8; ++j) x[j] = m[j] + x[j] * mscale;
            }
            if (want_raw) o_raw.q[g] = pack8(x);
            if (want_act) {
                float a[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) a[j] = fmaxf(x[j], x[j] * slope);      // slope in [0, 1): == leaky-relu
                o_act.q[g] = pack8(a);
            }
        }
        long long* dbg_rel = (dbg && c0 == dbg_c0) ? dbg + 4 : nullptr;
        if (dbg && c0 == dbg_c0) dbg[3] = clock64();
        if (mrf_store) {
            if (tm_mrf) io.store_tma(tm_mrf, n_base + c0, warp_row0, b, o_raw, dbg_rel);
            else io.store(mrf_blk + c0, e.n_total, o_raw, tma_out);
        } else {
            if (!kActOnly && e.out_raw) {
                if (tm_raw) io.store_tma(tm_raw, n_base + c0, warp_row0, b, o_raw, dbg_rel);
                else io.store(e.out_raw + row0 * e.ld_raw + n_base + c0, e.ld_raw, o_raw, tma_out);
            }
            if (want_act) {
                if (tm_act) io.store_tma(tm_act, n_base + c0, warp_row0, b, o_act, dbg_rel);
                else io.store(e.out_act + row0 * e.ld_act + n_base + c0, e.ld_act, o_act, tma_out);
            }
        }
        if (dbg && c0 == dbg_c0) dbg[5] = clock64();
        if (kLookahead) {
            res_cur = res_nxt;
            if (kMrf) mrf_cur = mrf_nxt;
        }
    }
    if (dbg) dbg[6] = clock64();
}

// ------------------------------------------------------------------------------------------------
// Act-only lean epilogue: the activated-chain vocoder launches whose ONE output is lrelu(v) through TMA stores —
// conv_pre, the up-samplers and conv1 of every two-launch pair (kRes = false), conv2 of pairs 0 / 1 (kRes = true).
// run_epilogue_lean decides residual / raw / activated / MRF at run time, so a launch without a residual still
// executed the residual's unpack + inverse-lrelu + add and packed two outputs: ~290 floating-point instructions per
// 32-column chunk where this one needs 112 (round-2 SASS count; the math was 700 of a chunk's 1900 cycles, and at two
// epilogue warps per scheduler those cycles are issue slots). kLd2: accumulator chunk c + 1 is in flight while chunk c
// is converted (tcgen05.ld took ~450 cycles under the other CTA's MMAs when it was waited for on the spot).
// Two chunks per loop trip keep both accumulator buffers in registers (n_tile: any multiple of 32).
// ------------------------------------------------------------------------------------------------
template <bool kRes, bool kLd2, class WaitFn, class DrainFn>
__device__ __forceinline__ void run_epilogue_act(const EpiParams& e, uint32_t taddr, int b, int t, int n_base, int n_tile,
                                                 WaitFn wait_acc, DrainFn acc_drained, uint8_t* stage, uint8_t* stage_in,
                                                 const LeanPrefetch<false>& pre, uint32_t bias_saddr,
                                                 const CUtensorMap* tm_act, long long* dbg = nullptr) {
    const int lane = threadIdx.x & 31;
    const int warp_row0 = t - lane;
    const long row0 = static_cast<long>(b) * e.T + warp_row0;
    const bool in_len = t < pre.len_rows;
    RowIO io{stage, lane, min(32, max(0, e.T - warp_row0)), e.debug};
    RowIO io_in{stage_in != nullptr ? stage_in : stage, lane, io.rows_valid, e.debug};
    const bool shared_tile = stage_in == nullptr || stage_in == stage;
    const __half* res_blk = e.residual + row0 * e.ld_res + n_base;
    const float slope = e.act_slope;
    const float rinv = e.res_inv;
    Chunk32 res_cur;
    if (kRes) res_cur = pre.res;
    wait_acc();
    if (dbg) dbg[0] = clock64();
    const bool any_masked = __any_sync(0xffffffffu, !in_len);
    float v[kLd2 ? 2 : 1][32];
    if (kLd2) tmem_ld32_issue(taddr, v[0]);
    auto step = [&](int c0, float (&cur)[32], float (&nxt)[32], bool more) {
        const bool stamp = dbg != nullptr && c0 == 32;
        if (stamp) dbg[1] = clock64();
        Chunk32 res_nxt;
        if (kRes) {
            io_in.to_row(res_cur, shared_tile);
            io.request(res_blk + c0 + 32, e.ld_res, more, res_nxt);
        }
        float bs_nxt[8];
        bias8s_early(bias_saddr + c0 * 4, bs_nxt);
        if (kLd2) {
            tmem_ld_wait(cur);
            if (more) tmem_ld32_issue(taddr + c0 + 32, nxt);
        } else {
            __syncwarp();
            tmem_ld32(taddr + c0, cur);
        }
        if (!more) acc_drained();
        if (stamp) dbg[2] = clock64();
        Chunk32 o;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            float bs[8], a[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) bs[j] = bs_nxt[j];
            if (g < 3) bias8s_early(bias_saddr + (c0 + (g + 1) * 8) * 4, bs_nxt);
            if (kRes) {
                float r[8];
                unpack8(res_cur.q[g], r);
#pragma unroll
                for (int j = 0; j < 8; ++j) a[j] = cur[g * 8 + j] + bs[j] + fminf(r[j], r[j] * rinv);   // rinv >= 1: inverse lrelu
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) a[j] = cur[g * 8 + j] + bs[j];
            }
            if (any_masked) {
#pragma unroll
                for (int j = 0; j < 8; ++j) a[j] = in_len ? a[j] : 0.f;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] = fmaxf(a[j], a[j] * slope);      // slope in [0, 1): == leaky-relu
            o.q[g] = pack8(a);
        }
        if (stamp) dbg[3] = clock64();
        io.store_tma(tm_act, n_base + c0, warp_row0, b, o, stamp ? dbg + 4 : nullptr);
        if (stamp) dbg[5] = clock64();
        if (kRes) res_cur = res_nxt;
    };
    for (int c0 = 0; c0 < n_tile; c0 += 64) {
        const bool two = c0 + 32 < n_tile;          // n_tile is a multiple of 32: an odd chunk count ends on a single step
        step(c0, v[0], v[kLd2 ? 1 : 0], two);
        if (two) step(c0 + 32, v[kLd2 ? 1 : 0], v[0], c0 + 64 < n_tile);
    }
    if (dbg) dbg[6] = clock64();
}

// ------------------------------------------------------------------------------------------------
// TMA-in lean epilogue: conv2 of every two-launch pair (residual add; MRF accumulate). The residual / MRF rows of a
// chunk arrive as 32 x 32 TMA boxes in per-warp shared-memory tiles whose 64-byte swizzle is the row-owner layout
// (wtile_off), `ring` chunks ahead, signalled on per-warp mbarriers. What this replaces (run_epilogue_lean): four
// LDG.128 per lane and chunk through a transposing staging tile — 12 shared-memory instructions and two warp syncs per
// chunk and kind, 16-32 prefetch registers, and above all a scoreboard that ptxas shares between those global loads and
// the bias LDS of the math loop, so the "look-ahead" loads were waited for inside the math of the SAME chunk (round-2
// timelines: 1700-2400 cycles of "math" per chunk where the act-only variant needs 300).
//   kMrfIn:    x = m + v * mrf_scale with m from the MRF buffer (MRF_ADD / MRF_LAST); else x = v * mrf_scale when
//              kStoreMrf (MRF_FIRST), x = v otherwise; v = acc + bias + lrelu^-1(residual), row-masked
//   kStoreMrf: x goes to the MRF buffer as it is (tm_out = its map); else lrelu(x) goes to out_act (tm_out = its map)
// ------------------------------------------------------------------------------------------------
struct TmaInState {
    uint8_t* tiles;   // this warp's input tiles: ring x (residual [, MRF]) x 2 KB, 512-byte aligned
    uint64_t* bars;   // this warp's `ring` mbarriers (arrival count 1)
    int ring;         // chunks of look-ahead = tiles per kind: 1 or 2
    int j;            // chunks consumed so far: slot j % ring, parity (j / ring) & 1
};

template <bool kMrfIn>
__device__ __forceinline__ void tma_in_issue(const TmaInState& st, int slot, const CUtensorMap* tm_res, const CUtensorMap* tm_mrf,
                                             int col, int row, int b) {
    constexpr uint32_t kBytes = kMrfIn ? 4096u : 2048u;
    uint8_t* dst = st.tiles + slot * kBytes;
    mbar_expect_tx(&st.bars[slot], kBytes);
    tma_load_3d(dst, tm_res, &st.bars[slot], col, row, b);
    if (kMrfIn) tma_load_3d(dst + 2048, tm_mrf, &st.bars[slot], col, row, b);
}

template <bool kMrfIn, bool kStoreMrf, bool kLd2, class WaitFn, class DrainFn>
__device__ __forceinline__ void run_epilogue_tma(const EpiParams& e, uint32_t taddr, int b, int t, int n_base, int n_tile,
                                                 WaitFn wait_acc, DrainFn acc_drained, uint8_t* stage, TmaInState& st,
                                                 int len_rows, uint32_t bias_saddr, const CUtensorMap* tm_out,
                                                 const CUtensorMap* tm_res, const CUtensorMap* tm_mrf, bool nvalid, int nb,
                                                 int nrow0, int* err_flag, long long* dbg = nullptr) {
    const int lane = threadIdx.x & 31;
    const int warp_row0 = t - lane;
    const bool in_len = t < len_rows;
    RowIO io{stage, lane, min(32, max(0, e.T - warp_row0)), e.debug};
    const float slope = e.act_slope;
    const float rinv = e.res_inv;
    const float mscale = e.mrf_scale;
    const int n_chunks = n_tile >> 5;
    constexpr uint32_t kBytes = kMrfIn ? 4096u : 2048u;
    wait_acc();
    if (dbg) dbg[0] = clock64();
    const bool any_masked = __any_sync(0xffffffffu, !in_len);
    float v[kLd2 ? 2 : 1][32];
    if (kLd2) tmem_ld32_issue(taddr, v[0]);
    auto step = [&](int ci, float (&cur)[32], float (&nxt)[32], bool more) {
        const int c0 = ci * 32;
        const bool stamp = dbg != nullptr && ci == 1;
        if (stamp) dbg[1] = clock64();
        const int slot = st.ring == 2 ? (st.j & 1) : 0;
        const uint32_t par = static_cast<uint32_t>(st.ring == 2 ? (st.j >> 1) : st.j) & 1u;
        mbar_wait(&st.bars[slot], par, err_flag, 208);
        const uint32_t tile_s = smem_u32(st.tiles) + slot * kBytes;
        uint4 rq[4], mq[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) rq[u] = lds128(tile_s + wtile_off(lane, u));
        if (kMrfIn) {
#pragma unroll
            for (int u = 0; u < 4; ++u) mq[u] = lds128(tile_s + 2048 + wtile_off(lane, u));
        }
        // Refill the slot with chunk j + ring — a later chunk of this tile or an early one of the warp's next tile —
        // BEFORE this chunk's math, so the copy has the math, the store and (ring = 2) a whole further chunk to arrive
        // (issued after the math it was still 1200-1900 cycles late: round-2 session 17). The TMA write must not overtake
        // the ld.shared above: "issued" is not "performed" — under the MMAs' operand traffic a shared-memory load can sit
        // in the queue longer than an L2-hit TMA copy takes (session 19: sparse corruption without the fence). The proxy
        // fence (MEMBAR.CTA + async-proxy fence) retires this lane's loads and orders them before the async-proxy write;
        // nothing else is outstanding here, so it is cheap. The warp sync extends that to all 32 lanes.
        fence_proxy_async();
        __syncwarp();
        if (elect_one()) {
            const int nci = ci + st.ring;
            if (nci < n_chunks) tma_in_issue<kMrfIn>(st, slot, tm_res, tm_mrf, n_base + nci * 32, warp_row0, b);
            else if (nvalid) tma_in_issue<kMrfIn>(st, slot, tm_res, tm_mrf, n_base + (nci - n_chunks) * 32, nrow0, nb);
        }
        __syncwarp();      // the tcgen05.ld / wait below are .sync.aligned: the elected lane must be back with the others
        float bs_nxt[8];
        bias8s_early(bias_saddr + c0 * 4, bs_nxt);
        if (kLd2) {
            tmem_ld_wait(cur);
            if (more) tmem_ld32_issue(taddr + c0 + 32, nxt);
        } else {
            tmem_ld32(taddr + c0, cur);
        }
        if (!more) acc_drained();
        if (stamp) dbg[2] = clock64();
        Chunk32 o;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            float bs[8], r[8], a[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) bs[j] = bs_nxt[j];
            if (g < 3) bias8s_early(bias_saddr + (c0 + (g + 1) * 8) * 4, bs_nxt);
            unpack8(rq[g], r);
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] = cur[g * 8 + j] + bs[j] + fminf(r[j], r[j] * rinv);   // rinv >= 1: inverse lrelu
            if (any_masked) {
#pragma unroll
                for (int j = 0; j < 8; ++j) a[j] = in_len ? a[j] : 0.f;
            }
            if (kMrfIn) {
                float m[8];
                unpack8(mq[g], m);
#pragma unroll
                for (int j = 0; j < 8; ++j) a[j] = m[j] + a[j] * mscale;
            } else if (kStoreMrf) {
#pragma unroll
                for (int j = 0; j < 8; ++j) a[j] = a[j] * mscale;
            }
            if (!kStoreMrf) {
#pragma unroll
                for (int j = 0; j < 8; ++j) a[j] = fmaxf(a[j], a[j] * slope);      // slope in [0, 1): == leaky-relu
            }
            o.q[g] = pack8(a);
        }
        if (stamp) dbg[3] = clock64();
        io.store_tma(tm_out, n_base + c0, warp_row0, b, o, stamp ? dbg + 4 : nullptr);
        if (stamp) dbg[5] = clock64();
        ++st.j;
    };
    for (int ci = 0; ci < n_chunks; ci += 2) {
        step(ci, v[0], v[kLd2 ? 1 : 0], true);
        step(ci + 1, v[kLd2 ? 1 : 0], v[0], ci + 2 < n_chunks);
    }
    if (dbg) dbg[6] = clock64();
}
#endif  // __CUDACC__

}  // namespace ttsb
