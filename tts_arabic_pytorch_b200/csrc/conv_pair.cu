// conv_pair: one ResBlock1 step  x' = x + conv2(lrelu(conv1(lrelu(x)) + b1)) + b2  in ONE launch
// (vocoder/hifigan/models.py:46-53, one (c1, c2) iteration; conv1 dilated, conv2 dilation 1).
//
// Why: with one launch per conv (conv_tc2) a pair moves six activation tensors through HBM
// (lrelu(x) in, t out, t in, x in, x' out, lrelu(x') out). For C <= 64 that traffic — not the
// tensor pipe — bounds the layer (profiles/r01_s18_launches_b64.csv: the C=32/64 stages take as long
// as the C=128 stage with 1/4..1/2 of its FLOPs). Here the pair reads x once and writes x' once:
//
//   TMA      x panel (128 + 2*h1 rows, raw)                          -> smem X slot
//   warps    lrelu in place (half2 max(x, 0.1x))                      -> the A operand of conv1
//   tcgen05  conv1: k row-shifted views of the panel x W1 (resident)  -> TMEM acc1
//   warps    acc1 + b1 -> lrelu -> row mask -> fp16, written in the UMMA swizzled layout
//                                                                     -> smem TT panel (never leaves the SM)
//   tcgen05  conv2: k row-shifted views of TT x W2                    -> TMEM acc2
//   warps    acc2 + b2 + x (re-read from L2) -> mask -> x' (or the MRF accumulate / stage output)
//
// A tile owns m_out = 128 - 2*h2 output rows: conv1 is evaluated on 128 rows (the tile plus conv2's
// halo), conv2 on 128 rows of which the last 2*h2 are discarded (they read beyond the TT rows).
// MMA issue order is conv1(i+1) before conv2(i), so the tensor pipe works on the next tile while the
// mid epilogue turns acc1(i) into TT(i); both accumulators are double-buffered in TMEM (4*C columns).
//
// Warp roles (512 threads): 0 TMA producer (x panels, resident weights), 1 conv1 MMA issuer, 2 conv2 MMA issuer,
// 3 W2 ring producer (streamed W2 only), 4-7 lrelu transform + mid epilogue, 8-11 and 12-15 two final-epilogue groups
// (the lean epilogue of epilogue.cuh) that take alternate items.
#include <cstdlib>
#include "conv.cuh"

namespace ttsb {

constexpr int kPairThreads = 512;   // 16 warps = 4 per scheduler: 128 registers per thread
constexpr int kPairXformWarps = 4;
constexpr int kPairMaxX = 8;
constexpr int kPairMaxB = 8;

struct ConvPairArgs {
    int B, T;
    int m_out;        // output rows owned by a tile
    int tiles_t;      // ceil(T / m_out)
    int n_work;       // B * tiles_t
    int C, chunk_k, n_chunks, n_taps;
    int h2;           // (k - 1) / 2
    int dil;          // conv1 dilation (tap step in rows); h1 = h2 * dil
    int rows_panel;   // x panel rows per chunk (multiple of 8, >= 128 + 2*h1)
    int tt_rows;      // TT panel rows per chunk (multiple of 8, >= 128 + 2*h2)
    int x_slots, tt_slots;
    int w2_resident, b_stages;
    int tt_pair;      // C = 32, resident W2: TT rows are 128 bytes [mid[t] | mid[t + 1]] and W2 comes as ceil(k / 2) tiles of
                      // 32 x 64 (taps 2g | 2g + 1): two taps per K = 64 group on 128-byte-swizzled operand rows (42 instead of
                      // 69 cycles per tcgen05.mma, profiles/r01_s21_mma_rate.txt); the mid epilogue writes every row twice
    int c2_split;     // resident W2, two TT slots: warp 3 (idle then) issues conv2 of the ODD items, warp 2 of the even ones
    int w2_x2;        // streamed W2 only: one pass of the W2 ring feeds conv2 of TWO consecutive items (both TT slots, both acc2 buffers)
    int smem_res;     // 1: kernel instantiated with kSmemRes (host-side record; see conv_pair_forward)
    int in_act;       // 1: x is stored activated (lrelu(x)): the TMA panel IS conv1's operand — no in-place transform, the conv1
                      // issuer waits for the panel itself; the final epilogue recovers x for the residual (EpiParams::res_inv)
    const __half* w1;
    const __half* w2;
    const float* bias1;
    float slope;
    int* err_flag;
    int debug;        // timing-decomposition switches (TTSB_PAIR_DEBUG; results are wrong when set): 1 no lrelu transform,
                      // 2 no TT stores, 4 no final epilogue work, 8 no MMAs
    long long* timeline;   // debug (tools/timeline_pair.py): 128 clock64() slots per CTA for the first 256 CTAs, or null
    EpiParams epi;    // final epilogue: bias = b2, residual = x, outputs
};

// slot = 8 + item*16 + k for the CTA's first 7 items; k: 0 x issued, 1 x landed (transform), 2 transform done,
// 3 conv1 operands ready, 4 conv1 issued, 5 conv2 operands ready, 6 conv2 issued, 7 acc1 seen (mid), 8 TT slot free,
// 9 mid done, 10 final epilogue starts waiting, 11 acc2 seen, 12 final done
__device__ __forceinline__ void tlp_mark(const ConvPairArgs& a, int item, int k) {
    if (a.timeline == nullptr) return;
    if (blockIdx.x < 256 && item < 7) a.timeline[blockIdx.x * 128 + 8 + item * 16 + k] = clock64();
}

__device__ __forceinline__ uint32_t lrelu_h2(uint32_t u, __half2 slope2) {
    __half2 x = *reinterpret_cast<__half2*>(&u);
    __half2 y = __hmax2(x, __hmul2(x, slope2));
    return *reinterpret_cast<uint32_t*>(&y);
}

// kTmemCols = 4 * C: 128 -> C = 32 (32-channel rows, 64 B swizzle, 2 K steps per tap), 256 / 512 -> C = 64 / 128
// lrelu of one landed x slot, in place, by the 128 threads of one warp group (tid = 0..127); arrives on xl_full (count 4)
__device__ __forceinline__ void pair_transform_slot(const ConvPairArgs& args, uint8_t* slot, int units, uint64_t* x_full,
                                                    uint64_t* xl_full, uint32_t parity, int tid, int item) {
    mbar_wait(x_full, parity, args.err_flag, 309);
    if (tid == 0) tlp_mark(args, item, 1);
    const __half2 slope2 = __float2half2_rn(args.slope);
    const uint32_t base = smem_u32(slot);
    // four independent 16-byte units per thread and pass: loads first, then the math, then the stores. The shared-memory
    // accessors are volatile asm, so a one-unit loop body is a serial load -> math -> store chain of ~110 cycles per
    // unit (profiles/r01_s48_pair_two_final_groups.txt: ~1000 cycles per 17 KB panel for this group).
    for (int i0 = tid; i0 < ((args.debug & 1) ? 0 : units); i0 += 128 * 4) {
        uint4 v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (i0 + j * 128 < units) v[j] = lds128(base + (i0 + j * 128) * 16);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            v[j].x = lrelu_h2(v[j].x, slope2); v[j].y = lrelu_h2(v[j].y, slope2);
            v[j].z = lrelu_h2(v[j].z, slope2); v[j].w = lrelu_h2(v[j].w, slope2);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (i0 + j * 128 < units) sts128(base + (i0 + j * 128) * 16, v[j]);
    }
    fence_proxy_async();
    __syncwarp();
    if (elect_one()) mbar_arrive(xl_full);
    if (tid == 0) tlp_mark(args, item, 2);
}

// kSmemRes: the final epilogue takes its residual rows from the x panel in shared memory (activated chain, deep x ring;
// see run_epilogue_lean) and releases the panel itself: x_empty then counts conv1's commit + the 4 final-epilogue warps.
template <int kTmemCols, int kEpi, bool kSmemRes, bool kTtPair>
__global__ void __launch_bounds__(kPairThreads, 1)
conv_pair_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ ConvPairArgs args) {
    constexpr int kKSteps = kTmemCols == 128 ? 2 : 4;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_u32 = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw_u32 & 1023u)) & 1023u);

    const int C = args.C;
    // a template parameter, not args.tt_pair: a run-time flag in the mid epilogue's store loop cost the C = 64
    // instantiations 5-15 % (round-2 session 28)
    constexpr bool tt_pair = kTtPair;
    const int row_bytes = args.chunk_k * 2;
    const int panel_bytes = args.rows_panel * row_bytes;
    const int xslot_bytes = args.n_chunks * panel_bytes;
    const int tt_row_bytes = tt_pair ? 128 : row_bytes;
    const int ttp_bytes = args.tt_rows * tt_row_bytes;
    const int ttslot_bytes = args.n_chunks * ttp_bytes;
    const int btile_bytes = C * row_bytes;
    const int n_btiles = args.n_chunks * args.n_taps;
    const int b2tile_bytes = tt_pair ? C * 128 : btile_bytes;                  // W2 tiles (resident form)
    const int n_b2tiles = tt_pair ? (args.n_taps + 1) / 2 : n_btiles;
    const int h1 = args.h2 * args.dil;

    uint8_t* smem_x = smem;
    uint8_t* smem_tt = smem_x + args.x_slots * xslot_bytes;
    uint8_t* smem_w1 = smem_tt + args.tt_slots * ttslot_bytes;
    uint8_t* smem_w2 = smem_w1 + n_btiles * btile_bytes;
    uint8_t* smem_end = smem_w2 + (args.w2_resident ? n_b2tiles * b2tile_bytes : args.b_stages * btile_bytes);
    uint64_t* x_full = reinterpret_cast<uint64_t*>(smem_end);
    uint64_t* xl_full = x_full + kPairMaxX;
    uint64_t* x_empty = xl_full + kPairMaxX;
    uint64_t* full_b = x_empty + kPairMaxX;
    uint64_t* empty_b = full_b + kPairMaxB;
    uint64_t* tt_full = empty_b + kPairMaxB;     // [2]
    uint64_t* tt_empty = tt_full + 2;            // [2]
    uint64_t* acc1_full = tt_empty + 2;          // [2]
    uint64_t* acc1_empty = acc1_full + 2;        // [2]
    uint64_t* acc2_full = acc1_empty + 2;        // [2]
    uint64_t* acc2_empty = acc2_full + 2;        // [2]
    uint64_t* w_full = acc2_empty + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 1);
    uint8_t* smem_stage = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tmem_slot + 4) + 127) & ~static_cast<uintptr_t>(127));
    float* sbias1 = reinterpret_cast<float*>(smem_stage + 8 * 2048);   // [C] conv1 bias, [C] conv2 bias: the epilogues read
    float* sbias2 = sbias1 + args.C;                                   // them as shared-memory broadcasts, not through L1/L2

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int grid = gridDim.x;

    if (warp == 0 && elect_one()) {
        if (args.timeline != nullptr && blockIdx.x < 256) args.timeline[blockIdx.x * 128] = clock64();
        tma_prefetch_desc(&tmap_x);
        for (int i = 0; i < args.x_slots; ++i) { mbar_init(&x_full[i], 1); mbar_init(&xl_full[i], kPairXformWarps); mbar_init(&x_empty[i], kSmemRes ? 5 : 1); }
        for (int i = 0; i < args.b_stages; ++i) { mbar_init(&full_b[i], 1); mbar_init(&empty_b[i], 1); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tt_full[i], 4); mbar_init(&tt_empty[i], 1);
            mbar_init(&acc1_full[i], 1); mbar_init(&acc1_empty[i], 4);
            mbar_init(&acc2_full[i], 1); mbar_init(&acc2_empty[i], 4);
        }
        mbar_init(w_full, 1);
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc<kTmemCols>(tmem_slot);
    for (int i = threadIdx.x; i < 2 * C; i += kPairThreads) sbias1[i] = i < C ? args.bias1[i] : args.epi.bias[i - C];
    // rows [128, tt_rows) of every TT panel are only read by the discarded output rows; keep them finite
    for (int i = threadIdx.x; i < args.tt_slots * ttslot_bytes / 16; i += kPairThreads)
        reinterpret_cast<uint4*>(smem_tt)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t acc1_col = 0, acc2_col = 2 * C;

    if (warp == 0) {
        // ---------------- TMA producer: resident weights once, then one x panel per item ----------------
        if (elect_one()) {
            const uint32_t wbytes = n_btiles * btile_bytes;
            mbar_expect_tx(w_full, wbytes + (args.w2_resident ? n_b2tiles * b2tile_bytes : 0));
            for (int i = 0; i < n_btiles; ++i)
                bulk_load_1d(smem_w1 + i * btile_bytes, reinterpret_cast<const uint8_t*>(args.w1) + static_cast<size_t>(i) * btile_bytes,
                             btile_bytes, w_full);
            if (args.w2_resident)
                for (int i = 0; i < n_b2tiles; ++i)
                    bulk_load_1d(smem_w2 + i * b2tile_bytes, reinterpret_cast<const uint8_t*>(args.w2) + static_cast<size_t>(i) * b2tile_bytes,
                                 b2tile_bytes, w_full);
            int sx = 0, it = 0;
            uint32_t px = 1;   // parity to wait on the EMPTY barrier (first lap passes)
            for (int idx = blockIdx.x; idx < args.n_work; idx += grid, ++it) {
                const int b = idx / args.tiles_t;
                const int t0 = (idx - b * args.tiles_t) * args.m_out;
                mbar_wait(&x_empty[sx], px, args.err_flag, 301);
                mbar_expect_tx(&x_full[sx], xslot_bytes);
                for (int c = 0; c < args.n_chunks; ++c)
                    tma_load_3d(smem_x + sx * xslot_bytes + c * panel_bytes, &tmap_x, &x_full[sx], c * args.chunk_k,
                                t0 - args.h2 - h1, b);
                tlp_mark(args, it, 0);
                if (++sx == args.x_slots) { sx = 0; px ^= 1; }
            }
        }
    } else if (warp == 3 && !args.c2_split) {
        // ---------------- W2 ring producer (only when W2 does not fit next to W1) ----------------
        // its own warp: an x load must never queue behind W2 tiles that wait for conv2 to drain the ring
        if (!args.w2_resident && elect_one()) {
            int sb = 0;
            uint32_t pb = 1;
            // w2_x2: one pass per PAIR of items (the conv2 issuer applies every tile to both)
            const int pass_stride = args.w2_x2 ? 2 * grid : grid;
            for (int idx = blockIdx.x; idx < args.n_work; idx += pass_stride) {
                const uint8_t* wp = reinterpret_cast<const uint8_t*>(args.w2);
                for (int i = 0; i < n_btiles; ++i) {
                    mbar_wait(&empty_b[sb], pb, args.err_flag, 302);
                    mbar_expect_tx(&full_b[sb], btile_bytes);
                    bulk_load_1d(smem_w2 + sb * btile_bytes, wp, btile_bytes, &full_b[sb]);
                    wp += btile_bytes;
                    if (++sb == args.b_stages) { sb = 0; pb ^= 1; }
                }
            }
        }
    } else if (warp == 1 || warp == 2 || warp == 3) {
        // ---------------- MMA issuers: warp 1 issues every conv1, warp 2 every conv2 ----------------
        // Two threads because the per-item serial chain of ONE issuer (4 barrier waits, 2 x k MMAs, 4 commits)
        // was the tile period (profiles/r01_s25_pair_decomposition.txt); the tensor pipe interleaves both streams.
        // One elected thread each (see conv_tc2.cu on elect.sync).
        // c2_split (resident W2, two TT slots): conv2 has THREE... two issuers of its own — warp 2 takes the even items
        // (TT slot 0, acc2 buffer 0), warp 3 the odd ones (slot 1, buffer 1). tcgen05.mma issue blocks on the pipe, so one
        // thread's (barrier waits + issue + commits) per item was the item period of the k = 3 pairs (round-2 session 31:
        // ~1 360 + 380 of 1 750 cycles at C = 32); the two accumulators and TT slots are independent, so are the threads.
        if (elect_one()) {
            const uint32_t idesc = umma_idesc_f16(kTileM, C);
            const uint32_t row_u = row_bytes >> 4;
            const uint64_t desc_hi = (static_cast<uint64_t>(((8u * row_bytes) >> 4) | (1u << 14) | ((row_bytes == 128 ? 2u : 4u) << 29)) << 32) |
                                     (1u << 16);
            const uint32_t btile_u = btile_bytes >> 4;
            auto issue = [&](uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t accumulate) {
                if (args.debug & 8) return;
#pragma unroll
                for (int k = 0; k < kKSteps; ++k)
                    umma_f16(d_tmem, desc_hi | ((a_lo + 2 * k) & 0x3FFFu), desc_hi | ((b_lo + 2 * k) & 0x3FFFu), idesc,
                             accumulate | static_cast<uint32_t>(k));
            };
            mbar_wait(w_full, 0, args.err_flag, 303);
            tc_fence_after();
            if (warp == 1) {
                const uint32_t x_lo0 = (smem_u32(smem_x) & 0x3FFFFu) >> 4;
                const uint32_t w1_lo0 = (smem_u32(smem_w1) & 0x3FFFFu) >> 4;
                const uint32_t panel_u = panel_bytes >> 4, xslot_u = xslot_bytes >> 4;
                const uint32_t tap1_u = args.dil * row_u;
                int sx = 0, b1 = 0, it = 0;
                uint32_t pxl = 0, pe1 = 3;   // xl_full parity; per-buffer parity bits of acc1_empty (first lap passes)
                for (int idx = blockIdx.x; idx < args.n_work; idx += grid, ++it) {
                    mbar_wait(args.in_act ? &x_full[sx] : &xl_full[sx], pxl, args.err_flag, 304);
                    mbar_wait(&acc1_empty[b1], (pe1 >> b1) & 1u, args.err_flag, 305);
                    pe1 ^= 1u << b1;
                    tc_fence_after();
                    tlp_mark(args, it, 3);
                    const uint32_t d = tmem_base + acc1_col + b1 * C;
                    uint32_t accumulate = 0;
                    uint32_t b_lo = w1_lo0;
                    for (int c = 0; c < args.n_chunks; ++c) {
                        uint32_t a_lo = x_lo0 + sx * xslot_u + c * panel_u;
                        for (int tap = 0; tap < args.n_taps; ++tap) {
                            issue(d, a_lo, b_lo, accumulate);
                            accumulate = 1;
                            a_lo += tap1_u;
                            b_lo += btile_u;
                        }
                    }
                    umma_commit(&x_empty[sx]);
                    umma_commit(&acc1_full[b1]);
                    tlp_mark(args, it, 4);
                    if (++sx == args.x_slots) { sx = 0; pxl ^= 1; }
                    b1 ^= 1;
                }
            } else {
                const uint32_t tt_lo0 = (smem_u32(smem_tt) & 0x3FFFFu) >> 4;
                const uint32_t w2_lo0 = (smem_u32(smem_w2) & 0x3FFFFu) >> 4;
                const uint32_t ttp_u = ttp_bytes >> 4, ttslot_u = ttslot_bytes >> 4;
                const int split = args.c2_split ? 1 : 0;
                const int me = split && warp == 3 ? 1 : 0;             // which half of the items this thread issues
                const int step = split ? 2 * grid : grid;
                int st = me, sb = 0, b2 = me, it = me;
                uint32_t ptt = 0, pb = 0, pe2 = 3;
                // Streamed W2, two items per pass. The ring (b_stages x one per-tap tile) delivers a tile every ~560 cycles
                // whatever the consumer does (bytes in flight / copy latency: C = 64, k = 11 timelines), against 4 x 48 cycles of
                // MMAs per tile and item — conv2 was waiting for weights two thirds of the time. Items i and i + 1 sit in the two
                // TT slots and accumulate into the two acc2 buffers, so every tile that arrives is applied to both: half the
                // weight bytes per output row. (Costs latency only: conv2(i) starts once mid(i + 1) is done.)
                if (args.w2_x2) {
                    const uint32_t d0 = tmem_base + acc2_col, d1 = d0 + C;
                    for (int idx = blockIdx.x; idx < args.n_work; idx += 2 * grid, it += 2) {
                        const bool two = idx + grid < args.n_work;
                        mbar_wait(&tt_full[0], ptt & 1u, args.err_flag, 306);
                        mbar_wait(&acc2_empty[0], pe2 & 1u, args.err_flag, 307);
                        if (two) {
                            mbar_wait(&tt_full[1], (ptt >> 1) & 1u, args.err_flag, 306);
                            mbar_wait(&acc2_empty[1], (pe2 >> 1) & 1u, args.err_flag, 307);
                        }
                        pe2 ^= two ? 3u : 1u;
                        tc_fence_after();
                        tlp_mark(args, it, 5);
                        uint32_t accumulate = 0;
                        for (int c = 0; c < args.n_chunks; ++c) {
                            uint32_t a0 = tt_lo0 + c * ttp_u, a1 = a0 + ttslot_u;
                            for (int tap = 0; tap < args.n_taps; ++tap) {
                                mbar_wait(&full_b[sb], pb, args.err_flag, 308);
                                tc_fence_after();
                                const uint32_t b_lo = w2_lo0 + sb * btile_u;
                                issue(d0, a0, b_lo, accumulate);
                                if (two) issue(d1, a1, b_lo, accumulate);
                                accumulate = 1;
                                a0 += row_u;
                                a1 += row_u;
                                umma_commit(&empty_b[sb]);
                                if (++sb == args.b_stages) { sb = 0; pb ^= 1; }
                            }
                        }
                        umma_commit(&tt_empty[0]);
                        umma_commit(&acc2_full[0]);
                        if (two) {
                            umma_commit(&tt_empty[1]);
                            umma_commit(&acc2_full[1]);
                        }
                        tlp_mark(args, it, 6);
                        if (two) tlp_mark(args, it + 1, 6);
                        ptt ^= two ? 3u : 1u;
                    }
                } else if (tt_pair) {
                    // two taps per K = 64 group: A rows are [mid[t] | mid[t + 1]] (128-byte swizzle), group g reads the
                    // panel shifted by 2g rows against W2's pair tile g; the last (odd) tap uses the first half only
                    const uint64_t desc_hi2 = (static_cast<uint64_t>(((8u * 128u) >> 4) | (1u << 14) | (2u << 29)) << 32) | (1u << 16);
                    const uint32_t b2tile_u = static_cast<uint32_t>(b2tile_bytes) >> 4;
                    for (int idx = blockIdx.x + me * grid; idx < args.n_work; idx += step, it += split + 1) {
                        mbar_wait(&tt_full[st], (ptt >> st) & 1u, args.err_flag, 306);
                        mbar_wait(&acc2_empty[b2], (pe2 >> b2) & 1u, args.err_flag, 307);
                        pe2 ^= 1u << b2;
                        tc_fence_after();
                        tlp_mark(args, it, 5);
                        const uint32_t d = tmem_base + acc2_col + b2 * C;
                        uint32_t a_lo = tt_lo0 + st * ttslot_u;
                        uint32_t b_lo = w2_lo0;
                        uint32_t accumulate = 0;
                        for (int g = 0; g < n_b2tiles; ++g) {
                            const int nk = (2 * g + 1 < args.n_taps) ? 4 : 2;
                            if (!(args.debug & 8)) {
#pragma unroll
                                for (int k = 0; k < 4; ++k)
                                    if (k < nk)
                                        umma_f16(d, desc_hi2 | ((a_lo + 2 * k) & 0x3FFFu), desc_hi2 | ((b_lo + 2 * k) & 0x3FFFu), idesc,
                                                 accumulate | static_cast<uint32_t>(k));
                            }
                            accumulate = 1;
                            a_lo += 16;            // two rows of 128 bytes
                            b_lo += b2tile_u;
                        }
                        umma_commit(&tt_empty[st]);
                        umma_commit(&acc2_full[b2]);
                        tlp_mark(args, it, 6);
                        ptt ^= 1u << st;
                        if (!split) {
                            if (args.tt_slots == 2) st ^= 1;
                            b2 ^= 1;
                        }
                    }
                } else
                for (int idx = blockIdx.x + me * grid; idx < args.n_work; idx += step, it += split + 1) {
                    mbar_wait(&tt_full[st], (ptt >> st) & 1u, args.err_flag, 306);
                    mbar_wait(&acc2_empty[b2], (pe2 >> b2) & 1u, args.err_flag, 307);
                    pe2 ^= 1u << b2;
                    tc_fence_after();
                    tlp_mark(args, it, 5);
                    const uint32_t d = tmem_base + acc2_col + b2 * C;
                    uint32_t accumulate = 0;
                    uint32_t b_res = w2_lo0;
                    for (int c = 0; c < args.n_chunks; ++c) {
                        uint32_t a_lo = tt_lo0 + st * ttslot_u + c * ttp_u;
                        for (int tap = 0; tap < args.n_taps; ++tap) {
                            uint32_t b_lo = b_res;
                            if (!args.w2_resident) {
                                mbar_wait(&full_b[sb], pb, args.err_flag, 308);
                                tc_fence_after();
                                b_lo = w2_lo0 + sb * btile_u;
                            }
                            issue(d, a_lo, b_lo, accumulate);
                            accumulate = 1;
                            a_lo += row_u;
                            b_res += btile_u;
                            if (!args.w2_resident) {
                                umma_commit(&empty_b[sb]);
                                if (++sb == args.b_stages) { sb = 0; pb ^= 1; }
                            }
                        }
                    }
                    umma_commit(&tt_empty[st]);
                    umma_commit(&acc2_full[b2]);
                    tlp_mark(args, it, 6);
                    ptt ^= 1u << st;
                    if (!split) {
                        if (args.tt_slots == 2) st ^= 1;
                        b2 ^= 1;
                    }
                }
            }
        }
    } else if (warp < 8) {
        // ---------------- mid epilogue: acc1 -> lrelu(acc + b1) * rowmask -> TT panel (UMMA layout) ----------------
        const int q = warp & 3;
        const int m = q * 32 + lane;
        int b1 = 0, st = 0;
        uint32_t pf = 0;
        uint32_t pte = 3;   // per-slot parity bits to wait on tt_empty (first lap passes)
        const float slope = args.slope;
        const uint32_t phase = row_bytes == 128 ? (m & 7) : ((m >> 1) & 3);
        int it = 0;
        const uint32_t sb1 = smem_u32(sbias1);
        // This group also runs the lrelu transform of the landed x panels (in place), `la` items ahead of its own
        // accumulator work: the four warps the transform used to own are the second final-epilogue group now — the
        // final epilogue (~4.5 k cycles per item) was the pipeline's slowest stage, this group has the slack.
        const int tid = threadIdx.x - 4 * 32;
        const int units = xslot_bytes >> 4;
        int sxt = 0;
        uint32_t pxt = 0;
        auto transform = [&](int j) {
            pair_transform_slot(args, smem_x + sxt * xslot_bytes, units, &x_full[sxt], &xl_full[sxt], pxt, tid, j);
            if (++sxt == args.x_slots) { sxt = 0; pxt ^= 1; }
        };
        int n_items = 0;
        for (int idx = blockIdx.x; idx < args.n_work; idx += grid) ++n_items;
        // x(it + la) is issued once conv1(it + la - x_slots) has completed; la = 0: the final groups transform
        const int la = args.in_act ? 0 : min(2, args.x_slots - 1);
        for (int j = 0; j < la && j < n_items; ++j) transform(j);
        // lens[b] of the NEXT item is fetched one item ahead: with the SM's L1 carved out as shared memory the load is an
        // L2 round trip (~700 cycles), and this group's item period is the pipeline's (round-2 timeline, gpurun r2_s1)
        constexpr int kC = kTmemCols / 4;
        auto len_of = [&](int idx) {
            if (idx >= args.n_work || args.epi.lens == nullptr) return args.T;
            return min(args.T, __ldg(args.epi.lens + idx / args.tiles_t) * args.epi.len_mul);
        };
        int len_nxt = len_of(blockIdx.x);
        for (int idx = blockIdx.x; idx < args.n_work; idx += grid, ++it) {
            const int b = idx / args.tiles_t;
            const int t0 = (idx - b * args.tiles_t) * args.m_out;
            const int t = t0 - args.h2 + m;
            const int len_rows = len_nxt;
            len_nxt = len_of(idx + grid);
            const bool valid = t >= 0 && t < len_rows;
            mbar_wait(&acc1_full[b1], (pf >> b1) & 1u, args.err_flag, 310);
            pf ^= 1u << b1;
            tc_fence_after();
            if (m == 0) tlp_mark(args, it, 7);
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc1_col + b1 * kC;
            // accumulator chunk c+1 is in flight while chunk c is converted (tcgen05.ld is asynchronous until its wait)
            float v[2][32];
            tmem_ld32_issue(taddr, v[0]);
            mbar_wait(&tt_empty[st], (pte >> st) & 1u, args.err_flag, 311);
            pte ^= 1u << st;
            if (m == 0) tlp_mark(args, it, 8);
            const uint32_t tt_base = smem_u32(smem_tt + st * ttslot_bytes) + m * row_bytes;
#pragma unroll
            for (int ci = 0; ci < kC / 32; ++ci) {
                const int c0 = ci * 32;
                // this chunk's bias values are requested before the accumulator load is waited for (bias8s_early)
                float bs_nxt[8];
                bias8s_early(sb1 + c0 * 4, bs_nxt);
                tmem_ld_wait(v[ci & 1]);
                if (ci == 0 && m == 0) tlp_mark(args, it, 13);
                if (ci + 1 < kC / 32) tmem_ld32_issue(taddr + c0 + 32, v[(ci + 1) & 1]);
                constexpr int kChunkK = kTmemCols == 128 ? 32 : 64;
                const int chunk = c0 / kChunkK;
                const int u0 = (c0 - chunk * kChunkK) >> 3;
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    float a[8], bs[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) bs[j] = bs_nxt[j];
                    if (g < 3) bias8s_early(sb1 + (c0 + (g + 1) * 8) * 4, bs_nxt);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float y = v[ci & 1][g * 8 + j] + bs[j];
                        a[j] = valid ? fmaxf(y, y * slope) : 0.f;      // slope in (0, 1): max(y, slope * y) == lrelu(y)
                    }
                    if (tt_pair) {
                        // row m, first half, and row m - 1, second half (128-byte rows, unit ^= row & 7)
                        const uint4 pk = pack8(a);
                        const uint32_t ra = tt_base - m * row_bytes + m * 128;      // this row; the slot is 1 KB aligned
                        sts128(ra + ((static_cast<uint32_t>(g) ^ ((ra >> 7) & 7u)) << 4), pk);
                        if (m > 0) sts128(ra - 128 + ((static_cast<uint32_t>(4 + g) ^ (((ra - 128) >> 7) & 7u)) << 4), pk);
                    } else if (!(args.debug & 2)) sts128(tt_base + chunk * ttp_bytes + ((u0 + g) ^ phase) * 16, pack8(a));
                }
            }
            if (m == 0) tlp_mark(args, it, 14);
            tc_fence_before();
            fence_proxy_async();
            __syncwarp();
            if (m == 0) tlp_mark(args, it, 15);
            if (elect_one()) {
                mbar_arrive(&acc1_empty[b1]);
                mbar_arrive(&tt_full[st]);
            }
            if (m == 0) tlp_mark(args, it, 9);
            b1 ^= 1;
            if (args.tt_slots == 2) st ^= 1;
            if (la > 0 && it + la < n_items) transform(it + la);
        }
    } else {
        // ---------------- final epilogue: acc2 + b2 + x -> mask -> outputs (lean epilogue) ----------------
        // two groups of four warps: group g takes the CTA's items g, g + 2, ... = accumulator buffer g
        const int g = (warp - 8) >> 2;
        const int q = warp & 3;
        constexpr bool kMrf = kEpi == 2;
        uint8_t* stage = smem_stage + (warp - 8) * 2048;
        const int b2 = g;
        uint32_t par = 0;
        LeanPrefetch<kMrf> pre_cur, pre_nxt;
        const int idx_first = blockIdx.x + g * grid;
        // kSmemRes: only lens[b] travels a tile ahead (the residual rows come from shared memory)
        auto prefetch = [&](const RowIO& io, long row0, bool on, LeanPrefetch<kMrf>& p, int b) {
            if (kSmemRes) p.len_rows = (on && args.epi.lens != nullptr) ? __ldg(args.epi.lens + b) * args.epi.len_mul : 0x7fffffff;
            else lean_prefetch(args.epi, io, row0, 0, on, p, b);
        };
        if (idx_first < args.n_work) {
            const int b0 = idx_first / args.tiles_t;
            const int t00 = (idx_first - b0 * args.tiles_t) * args.m_out;
            const int w0 = t00 + q * 32;
            RowIO io{stage, lane, min(32, max(0, min(args.T, t00 + args.m_out) - w0))};
            prefetch(io, static_cast<long>(b0) * args.T + w0, true, pre_cur, b0);
        }
        const bool tl_on = q == 0 && lane == 0;
        int n_items = 0;
        for (int idx = blockIdx.x; idx < args.n_work; idx += grid) ++n_items;
        // x ring position of this group's items (it = g, g + 2, ...): slot it % x_slots, parity (it / x_slots) & 1
        int sxr = g % args.x_slots;
        uint32_t pxr = static_cast<uint32_t>(g / args.x_slots) & 1u;
        const int prow = q * 32 + lane + args.h2 + h1;                 // this lane's row inside the x panel
        const uint32_t res_phase = row_bytes == 128 ? (prow & 7) : ((prow >> 1) & 3);
        int it = g;
        for (int idx = idx_first; idx < args.n_work; idx += 2 * grid, it += 2) {
            const int b = idx / args.tiles_t;
            const int t0 = (idx - b * args.tiles_t) * args.m_out;
            const int t = t0 + q * 32 + lane;
            if (tl_on) tlp_mark(args, it, 10);
            TmemAcc acc{tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc2_col + b2 * C};
            auto wait_acc = [&] {
                mbar_wait(&acc2_full[b2], par, args.err_flag, 312);
                // the panel was written by TMA and is read with ld.shared here: observe its own barrier (long complete)
                if (kSmemRes) mbar_wait(&x_full[sxr], pxr, args.err_flag, 313);
                tc_fence_after();
                if (tl_on) tlp_mark(args, it, 11);
            };
            auto drained = [&] {
                tc_fence_before();
                __syncwarp();
                if (elect_one()) mbar_arrive(&acc2_empty[b2]);
            };
            const int nidx = idx + 2 * grid;
            const bool nvalid = nidx < args.n_work;
            const int nb = nvalid ? nidx / args.tiles_t : 0;
            const int nt0 = nvalid ? (nidx - nb * args.tiles_t) * args.m_out : 0;
            const int nw0 = nt0 + q * 32;
            RowIO nio{stage, lane, min(32, max(0, min(args.T, nt0 + args.m_out) - nw0))};
            const uint32_t res_saddr = smem_u32(smem_x + sxr * xslot_bytes) + prow * row_bytes;
            // (detail stamps only where registers allow: the global-residual variants are at the 128-register limit)
            long long* dbg = (kSmemRes && tl_on && it == 4 && args.timeline != nullptr && blockIdx.x < 256) ? args.timeline + blockIdx.x * 128 + 120 : nullptr;
            if (args.debug & 4) {
                wait_acc();
                drained();
            } else {
                // The next tile's residual rows are requested right AFTER this tile's epilogue (they are in flight while the
                // next accumulator is waited for), not before it: 16 more live registers during the epilogue spill at the
                // 128-register budget, and a spilled load result is waited for on the spot (an L2 round trip per item).
                // Shared-memory residual: only lens[b] is fetched, a tile ahead.
                if (kSmemRes) prefetch(nio, static_cast<long>(nb) * args.T + nw0, nvalid, pre_nxt, nb);
                run_epilogue_lean<kMrf, true, !kMrf, kSmemRes, kSmemRes, kEpi == 3>(args.epi, acc, b, t, 0, C, wait_acc, drained, stage, pre_cur,
                                        t0 + args.m_out, smem_u32(sbias2), nullptr, nullptr, nullptr, nullptr, dbg, res_saddr, res_phase);
                if (!kSmemRes) prefetch(nio, static_cast<long>(nb) * args.T + nw0, nvalid, pre_nxt, nb);
            }
            if (kSmemRes) {
                // every residual read of this warp is done (ld.shared results were consumed above): release the panel
                __syncwarp();
                if (elect_one()) mbar_arrive(&x_empty[sxr]);
                sxr += 2;
                if (sxr >= args.x_slots) { sxr -= args.x_slots; pxr ^= 1u; }
            }
            pre_cur = pre_nxt;
            if (tl_on) tlp_mark(args, it, 12);
            par ^= 1;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc<kTmemCols>(tmem_base);
}

// ------------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------------
static const size_t kPairSmemMax = 232448;
static const size_t kPairFixed = 1024 /*alignment*/ + 1024 /*barriers + tmem slot*/ + 16384 /*staging: 8 warps*/ + 1024 /*biases*/ + 256;

ConvPairPlan conv_pair_plan(const ConvLayer& L1, const ConvLayer& L2) {
    ConvPairPlan p;
    static const int enabled = getenv("TTSB_PAIR") ? atoi(getenv("TTSB_PAIR")) : 1;
    if (!enabled) return p;
    const int C = L1.n_total;
    if (L1.cin != C || L2.cin != C || L2.n_total != C || L1.n_tile != C || L2.n_tile != C) return p;
    if (L1.n_taps != L2.n_taps || L1.n_taps % 2 != 1 || L1.chunk_k != L2.chunk_k) return p;
    if ((C != 32 && C != 64 && C != 128) || L1.chunk_k != (C == 32 ? 32 : 64)) return p;
    const int k = L1.n_taps, h2 = (k - 1) / 2;
    const int dil = k > 1 ? L1.tap_off[0][1] - L1.tap_off[0][0] : 1;
    for (int i = 0; i < k; ++i) {
        if (L1.tap_off[0][i] != (i - h2) * dil || L2.tap_off[0][i] != i - h2) return p;
        if (L1.tap_off[1][i] != L1.tap_off[0][i] || L2.tap_off[1][i] != L2.tap_off[0][i]) return p;
    }
    if (dil < 1 || L1.bias == nullptr || L2.bias == nullptr) return p;
    p.C = C; p.h2 = h2; p.dil = dil;
    p.m_out = kTileM - 2 * h2;
    if (p.m_out < 64) return p;
    // C = 32: TT rows of 128 bytes holding [t | t + 1], W2 as pair tiles (see ConvPairArgs::tt_pair). The 128-byte swizzle
    // wants every panel and tile 1 KB aligned: the x slots (64-byte rows) are padded to a multiple of 16 rows.
    static const int want_tt_pair = getenv("TTSB_PAIR_TT2") ? atoi(getenv("TTSB_PAIR_TT2")) : 1;
    // k = 3 pairs are bound by the epilogue groups, not by the MMAs: there the second store per row costs more than the
    // cheaper MMAs give back (129 -> 141 us per 16 utterances); k = 7 / 11: 197 -> 156 / 274 -> 170 us (profiles/r02_s27_...)
    const bool tt_pair = want_tt_pair >= 2 ? (C == 32 && L2.w_pair_packed != nullptr && k >= 3)
                                           : (want_tt_pair && C == 32 && L2.w_pair_packed != nullptr && k >= 5);
    p.rows_panel = round_up(kTileM + 2 * h2 * dil, tt_pair ? 16 : 8);
    p.tt_rows = round_up(kTileM + 2 * h2, 8);
    if (p.rows_panel > 256) return p;
    const size_t row_bytes = L1.chunk_k * 2;
    const size_t xslot = static_cast<size_t>(L1.n_chunks) * p.rows_panel * row_bytes;
    const size_t ttslot = static_cast<size_t>(L1.n_chunks) * p.tt_rows * (tt_pair ? 128 : row_bytes);
    const size_t btile = static_cast<size_t>(C) * row_bytes;
    const size_t w1 = static_cast<size_t>(L1.n_chunks) * k * btile;
    const size_t w2_res = tt_pair ? static_cast<size_t>((k + 1) / 2) * C * 128 : w1;
    const size_t avail = kPairSmemMax - kPairFixed;
    static const int force_stream = getenv("TTSB_PAIR_STREAM") ? atoi(getenv("TTSB_PAIR_STREAM")) : 0;
    for (int res = (force_stream && !tt_pair) ? 0 : 1; res >= (tt_pair ? 1 : 0) && !p.ok; --res) {
        const int min_b = std::min<int>(4, L1.n_chunks * k);
        const size_t wbytes = w1 + (res ? w2_res : min_b * btile);
        if (wbytes + 2 * xslot + ttslot > avail) continue;
        const size_t rem = avail - wbytes;
        p.tt_slots = rem >= 2 * ttslot + 2 * xslot ? 2 : 1;
        p.x_slots = static_cast<int>(std::min<size_t>(res ? 6 : 3, (rem - p.tt_slots * ttslot) / xslot));
        p.w2_resident = res;
        p.b_stages = 1;
        if (!res) {
            const size_t left = rem - p.tt_slots * ttslot - p.x_slots * xslot;
            p.b_stages = static_cast<int>(std::min<size_t>(kPairMaxB, min_b + left / btile));
            p.b_stages = std::min(p.b_stages, L1.n_chunks * k);
        }
        p.smem_bytes = kPairFixed + p.x_slots * xslot + p.tt_slots * ttslot + w1 + (res ? w2_res : p.b_stages * btile);
        p.tt_pair = (res && tt_pair) ? 1 : 0;
        p.ok = 1;
    }
    p.tmem_cols = 32;
    while (p.tmem_cols < 4 * C) p.tmem_cols *= 2;
    return p;
}

template <int kCols, int kEpi, bool kSmemRes, bool kTtPair>
static int launch_pair_impl2(const CUtensorMap& tm, const ConvPairArgs& a, int grid, size_t smem, cudaStream_t s) {
    static PerDeviceOnce configured;
    if (!configured.here()) {
        TTSB_CHECK_CUDA(cudaFuncSetAttribute(conv_pair_kernel<kCols, kEpi, kSmemRes, kTtPair>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             static_cast<int>(kPairSmemMax)));
        configured.here() = true;
    }
    conv_pair_kernel<kCols, kEpi, kSmemRes, kTtPair><<<grid, kPairThreads, smem, s>>>(tm, a);
    count_launch();
    TTSB_CHECK_CUDA(cudaGetLastError());
    return 0;
}

template <int kCols, int kEpi, bool kSmemRes>
static int launch_pair_impl(const CUtensorMap& tm, const ConvPairArgs& a, int grid, size_t smem, cudaStream_t s) {
    if constexpr (kCols == 128) {
        if (a.tt_pair) return launch_pair_impl2<kCols, kEpi, kSmemRes, true>(tm, a, grid, smem, s);
    }
    return launch_pair_impl2<kCols, kEpi, kSmemRes, false>(tm, a, grid, smem, s);
}

// kEpi: 1 lean, 2 lean + MRF accumulate, 3 lean with ONE activated output and nothing else (the activated chain's pairs 0 / 1)
template <int kCols>
static int launch_pair(const CUtensorMap& tm, const ConvPairArgs& a, int grid, size_t smem, cudaStream_t s, bool mrf) {
    static const int want_act_only = getenv("TTSB_PAIR_ACT_ONLY") ? atoi(getenv("TTSB_PAIR_ACT_ONLY")) : 1;
    const bool act_only = want_act_only && !mrf && a.epi.out_raw == nullptr && a.epi.out_act != nullptr;
    if (a.smem_res) {
        if (act_only) return launch_pair_impl<kCols, 3, true>(tm, a, grid, smem, s);
        return mrf ? launch_pair_impl<kCols, 2, true>(tm, a, grid, smem, s) : launch_pair_impl<kCols, 1, true>(tm, a, grid, smem, s);
    }
    if (act_only) return launch_pair_impl<kCols, 3, false>(tm, a, grid, smem, s);
    return mrf ? launch_pair_impl<kCols, 2, false>(tm, a, grid, smem, s) : launch_pair_impl<kCols, 1, false>(tm, a, grid, smem, s);
}

int conv_pair_forward(const ConvLayer& L1, const ConvLayer& L2, const ConvPairPlan& plan, const ConvRuntime& rt,
                      const __half* x, int B, int T, float slope, EpiParams epi, cudaStream_t stream, int in_act) {
    TTSB_REQUIRE(plan.ok, "conv pair plan is not feasible");
    TTSB_REQUIRE(B > 0 && T > 0, "empty batch");
    TTSB_REQUIRE(slope > 0.f && slope < 1.f, "leaky-relu slope must be in (0, 1)");
    TTSB_REQUIRE(epi.ln_g == nullptr && epi.head_w == nullptr && epi.out_f32 == nullptr && epi.out_f32_t == nullptr &&
                     epi.act_tanh == 0 && epi.pre_ln_relu == 0,
                 "conv pair supports the lean epilogue only");
    const CUtensorMap* tm = nullptr;
    TTSB_PROPAGATE(get_act_tensor_map(x, plan.C, B, T, plan.C, L1.chunk_k, plan.rows_panel, &tm));
    ConvPairArgs a;
    a.B = B; a.T = T;
    a.m_out = plan.m_out;
    a.tiles_t = ceil_div(T, plan.m_out);
    a.n_work = B * a.tiles_t;
    a.C = plan.C; a.chunk_k = L1.chunk_k; a.n_chunks = L1.n_chunks; a.n_taps = L1.n_taps;
    a.h2 = plan.h2; a.dil = plan.dil;
    a.rows_panel = plan.rows_panel; a.tt_rows = plan.tt_rows;
    a.x_slots = plan.x_slots; a.tt_slots = plan.tt_slots;
    a.w2_resident = plan.w2_resident; a.b_stages = plan.b_stages;
    static const int want_w2_x2 = getenv("TTSB_PAIR_W2X2") ? atoi(getenv("TTSB_PAIR_W2X2")) : 1;
    a.w2_x2 = (want_w2_x2 && !plan.w2_resident && plan.tt_slots == 2) ? 1 : 0;
    static const int want_c2_split = getenv("TTSB_PAIR_C2_SPLIT") ? atoi(getenv("TTSB_PAIR_C2_SPLIT")) : 1;
    // measured (profiles/r02_s39_pair_two_conv2_issuers.txt): C = 32 pairs 1-5 % faster, C = 64 k = 7 5 % slower, the rest
    // unchanged — the conv2 issue time is pipe latency under the epilogues' shared-memory traffic, not thread overhead
    a.c2_split = (want_c2_split && plan.w2_resident && plan.tt_slots == 2 && (plan.C == 32 || want_c2_split >= 2)) ? 1 : 0;
    a.in_act = in_act ? 1 : 0;
    a.tt_pair = plan.tt_pair;
    a.w1 = L1.w_packed; a.w2 = plan.tt_pair ? L2.w_pair_packed : L2.w_packed;
    a.bias1 = L1.bias;
    a.slope = slope;
    a.err_flag = rt.err_flag;
    a.timeline = rt.timeline;
    static const int debug = getenv("TTSB_PAIR_DEBUG") ? atoi(getenv("TTSB_PAIR_DEBUG")) : 0;
    a.debug = debug;
    epi.T = T;
    epi.n_total = plan.C;
    epi.bias = L2.bias;
    epi.residual = x;
    epi.ld_res = plan.C;
    epi.res_inv = in_act ? 1.f / slope : 1.f;
    a.epi = epi;
    // shared-memory residual: activated input (the panel holds what the residual add needs), one channel chunk, and an x
    // ring deep enough to keep a panel until its item's final epilogue (producer -> conv1 -> mid -> conv2 -> final)
    static const int want_smem_res = getenv("TTSB_PAIR_SMEM_RES") ? atoi(getenv("TTSB_PAIR_SMEM_RES")) : 1;
    static const int smem_res_min_slots = getenv("TTSB_PAIR_SMEM_RES_SLOTS") ? atoi(getenv("TTSB_PAIR_SMEM_RES_SLOTS")) : 5;
    a.smem_res = (want_smem_res && in_act && L1.n_chunks == 1 && plan.x_slots >= smem_res_min_slots) ? 1 : 0;
    const int grid = std::min(num_sms(), a.n_work);
    const bool mrf = epi.mrf_mode != MRF_NONE;
    switch (plan.tmem_cols) {
        case 128: return launch_pair<128>(*tm, a, grid, plan.smem_bytes, stream, mrf);
        case 256: return launch_pair<256>(*tm, a, grid, plan.smem_bytes, stream, mrf);
        case 512: return launch_pair<512>(*tm, a, grid, plan.smem_bytes, stream, mrf);
    }
    TTSB_REQUIRE(false, "bad tmem_cols for conv pair");
    return 1;
}

}  // namespace ttsb
