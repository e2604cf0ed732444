// Single-head attention (MultiHeadAttn._forward, transformer.py:131-146: bmm(q, k^T) * scale, key-padding mask, softmax,
// bmm(p, v)) with BOTH contractions on the 5th-generation tensor cores:
//
//   S  = Q K^T      tcgen05.mma  M = 128 queries, N = 128 keys, K = 64 (d_head)   A = Q tile, B = K tile   (both K-major)
//   O += P V        tcgen05.mma  M = 128 queries, N = 64  (d_head), K = 128 keys  A = P tile, B = V^T tile (both K-major)
//
// One CTA = 128 queries of one utterance; keys / values stream in tiles of 128 through a double-buffered TMA ring; the
// score tile lives in TMEM, one thread owns one query row of it (no shuffles in the online softmax), P goes back to shared
// memory as fp16 in the UMMA K-major swizzled layout, and the running output stays in registers (the P V product of a
// tile lands in a fresh TMEM accumulator and is folded in with the softmax correction). Key tiles beyond the
// utterance's length are skipped. V^T ([B, 64, S_pad], keys contiguous) is produced by a small transpose kernel so that
// the second contraction's B operand is K-major like everything else in this library (no [S, S] tensor ever exists in
// HBM, unlike transformer.py:131-141).
#include <cstdlib>
#include "conv.cuh"
#include "kernels.cuh"

namespace ttsb {

constexpr int kAttQ = 128;     // queries per CTA
constexpr int kAttK = 128;     // keys per tile
constexpr int kAttD = 64;      // d_head

// qkv [B, S, 192] (q | k | v) -> vt [B, 64, S_pad]: vt[b, d, s] = v[b, s, d]; keys >= S are written as zeros
__global__ void __launch_bounds__(256) attention_vt_kernel(const __half* __restrict__ qkv, int S, int S_pad,
                                                           __half* __restrict__ vt) {
    __shared__ __half tile[64][66];
    const int b = blockIdx.y, s0 = blockIdx.x * 64;
    for (int i = threadIdx.x; i < 64 * 8; i += 256) {
        const int r = i >> 3, c8 = i & 7;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (s0 + r < S) v = *reinterpret_cast<const uint4*>(qkv + (static_cast<size_t>(b) * S + s0 + r) * 192 + 128 + c8 * 8);
        const __half* hv = reinterpret_cast<const __half*>(&v);
#pragma unroll
        for (int j = 0; j < 8; ++j) tile[r][c8 * 8 + j] = hv[j];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 64 * 32; i += 256) {
        const int d = i >> 5, s2 = (i & 31) * 2;
        __half2 h = __halves2half2(tile[s2][d], tile[s2 + 1][d]);
        *reinterpret_cast<__half2*>(vt + (static_cast<size_t>(b) * 64 + d) * S_pad + s0 + s2) = h;
    }
}

struct AttTcArgs {
    const int* lens;
    int S;
    float scale_log2;      // softmax scale * log2(e)
    __half* out;           // [B, S, 64]
    int* err_flag;
};

__global__ void __launch_bounds__(128, 2)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmap_qk, const __grid_constant__ CUtensorMap tmap_vt,
                    const __grid_constant__ AttTcArgs args) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_u32 = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw_u32 & 1023u)) & 1023u);
    uint8_t* sQ = smem;                          // 128 x 128 B
    uint8_t* sK = sQ + 16384;                    // [2] 128 x 128 B
    uint8_t* sV = sK + 2 * 16384;                // [2][2 key chunks] 64 (d) x 128 B (64 keys)
    // P tile = two 64-key chunks of 128 x 128 B: chunk 0 reuses the current K buffer (dead once the score MMAs have
    // completed; it is reloaded only after this tile's P V MMAs), chunk 1 has its own 16 KB — two CTAs fit on an SM
    uint8_t* sP1 = sV + 2 * 16384;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP1 + 16384);
    uint64_t* q_full = bars;
    uint64_t* kv_full = bars + 1;                // [2]
    uint64_t* s_full = bars + 3;
    uint64_t* o_full = bars + 4;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);

    const int b = blockIdx.y;
    const int q0 = blockIdx.x * kAttQ;
    const int S = args.S;
    const int len = args.lens != nullptr ? min(__ldg(args.lens + b), S) : S;
    const int n_tiles = (len + kAttK - 1) / kAttK;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m = threadIdx.x;                   // this thread's query row inside the tile = its TMEM lane

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap_qk);
        tma_prefetch_desc(&tmap_vt);
        mbar_init(q_full, 1);
        mbar_init(&kv_full[0], 1);
        mbar_init(&kv_full[1], 1);
        mbar_init(s_full, 1);
        mbar_init(o_full, 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc<256>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_s = tmem_base;           // 128 columns: score tile
    const uint32_t tmem_o = tmem_base + 128;     // 64 columns: P V of the current key tile

    auto load_kv = [&](int kt) {                 // one thread
        const int buf = kt & 1;
        mbar_expect_tx(&kv_full[buf], 16384 + 16384);
        tma_load_3d(sK + buf * 16384, &tmap_qk, &kv_full[buf], 64, kt * kAttK, b);            // k = channels 64..127
        tma_load_3d(sV + buf * 16384, &tmap_vt, &kv_full[buf], kt * kAttK, 0, b);             // keys 0..63 of the tile
        tma_load_3d(sV + buf * 16384 + 8192, &tmap_vt, &kv_full[buf], kt * kAttK + 64, 0, b); // keys 64..127
    };
    if (threadIdx.x == 0 && n_tiles > 0) {
        mbar_expect_tx(q_full, 16384);
        tma_load_3d(sQ, &tmap_qk, q_full, 0, q0, b);                                          // q = channels 0..63
        load_kv(0);
    }

    const uint32_t idesc_s = umma_idesc_f16(kAttQ, kAttK);
    const uint32_t idesc_o = umma_idesc_f16(kAttQ, kAttD);
    float o[kAttD];
#pragma unroll
    for (int j = 0; j < kAttD; ++j) o[j] = 0.f;
    float m_run = -INFINITY, l_run = 0.f;
    const uint32_t lane_addr = static_cast<uint32_t>(warp * 32) << 16;
    const uint32_t phase = m & 7;

    for (int kt = 0; kt < n_tiles; ++kt) {
        const int buf = kt & 1;
        if (threadIdx.x == 0) {
            if (kt + 1 < n_tiles) load_kv(kt + 1);            // buffer buf^1: tile kt-1's MMAs were waited for (o_full)
            if (kt == 0) mbar_wait(q_full, 0, args.err_flag, 601);
            mbar_wait(&kv_full[buf], (kt >> 1) & 1, args.err_flag, 602);
            tc_fence_after();
            const uint64_t qd = umma_desc_sw128_kmajor(smem_u32(sQ), 0);
            const uint64_t kd = umma_desc_sw128_kmajor(smem_u32(sK + buf * 16384), 0);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16(tmem_s, qd + 2 * k, kd + 2 * k, idesc_s, k > 0 ? 1u : 0u);
            umma_commit(s_full);
        }
        mbar_wait(s_full, kt & 1, args.err_flag, 603);
        tc_fence_after();
        // pass 1: row maximum of the scaled, masked scores
        const int key0 = kt * kAttK;
        float mx = -INFINITY;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            float v[32];
            tmem_ld32(tmem_s + lane_addr + c * 32, v);
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (key0 + c * 32 + j < len) mx = fmaxf(mx, v[j] * args.scale_log2);
        }
        const float m_new = fmaxf(m_run, mx);                 // finite: every processed tile holds >= 1 valid key
        const float corr = exp2f(m_run - m_new);
        m_run = m_new;
        // pass 2: p = exp2(s - m) as fp16 into the K-major swizzled P tile (two 64-key chunks), row sum in fp32
        float ps = 0.f;
        const uint32_t p_row[2] = {smem_u32(sK + buf * 16384) + m * 128, smem_u32(sP1) + m * 128};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            float v[32];
            tmem_ld32(tmem_s + lane_addr + c * 32, v);
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                float p[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int key = key0 + c * 32 + g * 8 + j;
                    p[j] = key < len ? exp2f(v[g * 8 + j] * args.scale_log2 - m_new) : 0.f;
                }
                const uint4 pk = pack8(p);
                // the row sum uses the ROUNDED probabilities, the values the second contraction really multiplies
                float pr[8];
                unpack8(pk, pr);
#pragma unroll
                for (int j = 0; j < 8; ++j) ps += pr[j];
                const int chunk = c >> 1, u = (c & 1) * 4 + g;
                sts128(p_row[chunk] + ((static_cast<uint32_t>(u) ^ phase) << 4), pk);
            }
        }
        l_run = l_run * corr + ps;
        tc_fence_before();
        fence_proxy_async();
        __syncthreads();
        if (threadIdx.x == 0) {
            tc_fence_after();
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {
                const uint64_t pd = umma_desc_sw128_kmajor(smem_u32(ch == 0 ? sK + buf * 16384 : sP1), 0);
                const uint64_t vd = umma_desc_sw128_kmajor(smem_u32(sV + buf * 16384 + ch * 8192), 0);
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_f16(tmem_o, pd + 2 * k, vd + 2 * k, idesc_o, (ch | k) ? 1u : 0u);
            }
            umma_commit(o_full);
        }
        mbar_wait(o_full, kt & 1, args.err_flag, 604);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            float v[32];
            tmem_ld32(tmem_o + lane_addr + c * 32, v);
#pragma unroll
            for (int j = 0; j < 32; ++j) o[c * 32 + j] = o[c * 32 + j] * corr + v[j];
        }
        tc_fence_before();
        __syncthreads();                                      // TMEM and the P tile are free for the next key tile
    }
    const int q = q0 + m;
    if (q < S) {
        const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
        __half* orow = args.out + (static_cast<size_t>(b) * S + q) * kAttD;
#pragma unroll
        for (int g = 0; g < 8; ++g) {
            float x[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) x[j] = o[g * 8 + j] * inv;
            *reinterpret_cast<uint4*>(orow + g * 8) = pack8(x);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<256>(tmem_base);
}

size_t attention_tc_scratch_bytes(int B, int S) {
    return static_cast<size_t>(B) * kAttD * round_up(S, 64) * sizeof(__half);
}

int launch_attention_tc(const __half* qkv, const int* lens, int B, int S, float scale, __half* out, __half* vt_scratch,
                        int* err_flag, cudaStream_t s) {
    TTSB_REQUIRE(vt_scratch != nullptr, "attention scratch missing");
    const int S_pad = round_up(S, 64);
    attention_vt_kernel<<<dim3(S_pad / 64, B), 256, 0, s>>>(qkv, S, S_pad, vt_scratch);
    count_launch();
    TTSB_CHECK_CUDA(cudaGetLastError());
    const CUtensorMap* t = nullptr;
    TTSB_PROPAGATE(get_act_tensor_map(qkv, 192, B, S, 192, 64, kAttQ, &t));
    const CUtensorMap tm_qk = *t;
    TTSB_PROPAGATE(get_act_tensor_map(vt_scratch, S_pad, B, kAttD, S_pad, 64, kAttD, &t));
    const CUtensorMap tm_vt = *t;
    AttTcArgs a;
    a.lens = lens; a.S = S; a.scale_log2 = scale * 1.4426950408889634f; a.out = out; a.err_flag = err_flag;
    const size_t smem = 1024 + 16384 + 2 * 32768 + 16384 + 64;
    static PerDeviceOnce configured;
    if (!configured.here()) {
        TTSB_CHECK_CUDA(cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        configured.here() = true;
    }
    attention_tc_kernel<<<dim3(ceil_div(S, kAttQ), B), 128, smem, s>>>(tm_qk, tm_vt, a);
    count_launch();
    TTSB_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace ttsb
