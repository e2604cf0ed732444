// conv_tc2: persistent, warp-specialised tcgen05 kernel for "row GEMM with taps" (conv.cuh).
//
// Same math, data layout and smem/TMEM operand scheme as conv_tc.cu (v1: one tile per CTA, kept as
// a cross-check); what changes is the schedule, driven by the v1 in-kernel timeline
// (profiles/r01_s3_timeline.txt): v1 spent ~860 cycles of scalar bookkeeping per (chunk, tap)
// iteration in the single MMA-issuing thread, ~2500 cycles of per-CTA setup, and an epilogue that
// serialised ~16 dependent global loads after the accumulator was ready.
//
//   * one persistent CTA per SM slot; a CTA owns ONE N tile and walks (utterance, row-tile) work
//     items with stride = #CTAs of that N tile; barriers / TMEM / tensor-map setup happen once
//   * lean issue loops: ring stage + phase kept incrementally (no div/mod), UMMA descriptors built
//     once and advanced by adding to their low word, tap offsets as off0 + tap*step
//   * TMEM accumulators double-buffered (when 2*n_tile <= 512 columns): the epilogue of tile i
//     overlaps the MMAs of tile i+1
//   * epilogue requests its residual / MRF rows BEFORE the accumulator is ready (epilogue.cuh)
//
// Warp roles (192 threads): warp 0 TMA producer, warp 1 MMA issuer, warps 2..5 epilogue.
#include <cstdio>
#include <cstdlib>
#include <utility>
#include <vector>
#include "conv.cuh"

namespace ttsb {

struct ConvTc2Args {
    int B, T;
    int tiles_t;       // row tiles per utterance
    int rpp;           // row tiles per weight pass (1, 2 or 4): one work item = rpp consecutive tiles
    int groups_t;      // work items per utterance = ceil(tiles_t / rpp)
    int n_work;        // B * groups_t
    int resident;      // 1: all weight tiles of this N tile stay in smem for the CTA's lifetime
    int cluster;       // 1, or 2: CTA pairs share every streamed weight tile (each loads half, multicast to both)
    int n_tiles_n;
    int n_chunks, n_taps;
    int chunk_k;       // 64 or 32
    int rows_panel;
    int n_tile, n_sub;
    int a_slots, b_stages;
    int acc_bufs;      // 1 or 2
    int a_lanes, b_lanes;   // producer lanes; each divides its ring size so a ring slot is always served by the same lane
    int class_split;
    int shift0[2];     // halo_lo + off(tap 0), per tap class
    int step[2];       // off(tap+1) - off(tap)
    int halo_lo;
    const __half* w;
    int tma_out;       // bit 0 / 1 / 2: out_raw / out_act / mrf_buf leave through TMA stores (lean epilogue)
    int stage2;        // lean epilogue: a second set of 4 x 2 KB staging tiles (transposes) next to the output tiles
    int in_ring;       // TMA-in epilogue: chunks of look-ahead (input tiles per warp and kind), 1 or 2
    int sbias_bytes;   // size of the parameter tile region that precedes them
    int issue_mode;    // MMA issuer: 0 generic loops, 1 straight-line K steps, 2 + the next weight stage's barrier is tested ahead
    int params_smem;   // general epilogue: [bias][ln_g][ln_b][head_w] of this N tile staged in shared memory (room permitting)
    int* err_flag;
    long long* timeline;   // debug (tools/timeline.py): 64 clock64() slots per CTA for the first 256 CTAs, or null
    EpiParams epi;
};

// slot = 8 + tile*8 + k for the CTA's first 7 tiles; k: 0 mma got TMEM, 1 mma got A, 2 mma issued all,
// 3 epilogue starts waiting, 4 accumulator seen, 5 TMEM handed back, 6 epilogue done, 7 panel load issued
__device__ __forceinline__ void tl2_mark(const ConvTc2Args& a, int slot) {
    if (a.timeline == nullptr) return;
    if (blockIdx.x < 256 && slot < 64) a.timeline[blockIdx.x * 128 + slot] = clock64();
}

// kEpi: 0 = general epilogue, 1 = lean (vocoder hot subset), 2 = lean + MRF accumulate, 3 = act-only lean without a residual
// (run_epilogue_act), 4..7 = TMA-in lean with a residual (run_epilogue_tma): 4 -> out_act, 5 MRF_FIRST, 6 MRF_ADD, 7 MRF_LAST
template <int kTmemCols, int kMinBlocks, int kEpi>
__global__ void __launch_bounds__(192, kMinBlocks)
conv_tc2_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ ConvTc2Args args,
                const __grid_constant__ CUtensorMap tmap_raw, const __grid_constant__ CUtensorMap tmap_act,
                const __grid_constant__ CUtensorMap tmap_mrf, const __grid_constant__ CUtensorMap tmap_res) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_u32 = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw_u32 & 1023u)) & 1023u);

    const int row_bytes = args.chunk_k * 2;
    const int panel_bytes = args.rows_panel * row_bytes;
    const int btile_bytes = args.n_tile * row_bytes;
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + args.a_slots * panel_bytes;
    const int n_btiles = args.n_chunks * args.n_taps;
    const int b_slots = args.resident ? n_btiles : args.b_stages;
    uint64_t* full_a = reinterpret_cast<uint64_t*>(smem_b + b_slots * btile_bytes);
    uint64_t* empty_a = full_a + args.a_slots;
    uint64_t* full_b = empty_a + args.a_slots;
    uint64_t* empty_b = full_b + args.b_stages;
    uint64_t* tmem_full = empty_b + args.b_stages;   // [2]
    uint64_t* tmem_empty = tmem_full + 2;            // [2]
    uint64_t* w_full = tmem_empty + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 1);
    // 4 x 2 KB staging tiles for the epilogue warps' coalesced row I/O (epilogue.cuh)
    // 1 KB aligned: the staging tiles double as TMA-store sources with the 64-byte swizzle (address-bit XOR)
    uint64_t* in_bar = reinterpret_cast<uint64_t*>(tmem_slot + 4);   // [4 warps][2]: TMA-in epilogue (kEpi >= 4)
    uint8_t* smem_stage = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(in_bar + 8) + 1023) & ~static_cast<uintptr_t>(1023));
    float* sbias = reinterpret_cast<float*>(smem_stage + 4 * 2048);   // [n_tile] this N tile's bias (lean epilogue: smem broadcast)

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    // Work items: cluster `cid` owns N tile cid % n_tiles_n and walks item groups p = first, first+stride, ...;
    // CTA `crank` of the cluster takes item p*csize + crank (a missing last item is a dummy: all rows out
    // of range). Both CTAs of a pair therefore make exactly the same number of weight passes.
    const int csize = args.cluster;
    const int crank = csize == 2 ? static_cast<int>(cluster_ctarank()) : 0;
    const int cid = blockIdx.x / csize;
    const int ntile = cid % args.n_tiles_n;
    const int first = cid / args.n_tiles_n;
    const int stride = (gridDim.x / csize) / args.n_tiles_n;
    const int n_groups = (args.n_work + csize - 1) / csize;
    const uint16_t cmask = static_cast<uint16_t>((1u << csize) - 1u);

    if (warp == 0 && elect_one()) {
        tl2_mark(args, 0);
        tma_prefetch_desc(&tmap_a);
        for (int i = 0; i < args.a_slots; ++i) { mbar_init(&full_a[i], 1); mbar_init(&empty_a[i], 1); }
        for (int i = 0; i < args.b_stages; ++i) { mbar_init(&full_b[i], 1); mbar_init(&empty_b[i], csize); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 4); }
        mbar_init(w_full, 1);
        if (kEpi >= 4 && kEpi <= 7)
            for (int i = 0; i < 8; ++i) mbar_init(&in_bar[i], 1);
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc<kTmemCols>(tmem_slot);
    if (kEpi != 0 && kEpi != 8) {
        for (int i = threadIdx.x; i < args.n_tile; i += 192)
            sbias[i] = args.epi.bias != nullptr ? args.epi.bias[ntile * args.n_tile + i] : 0.f;
    } else if (args.params_smem) {
        // (selects, not an indexed pointer array: that array was the kernel's only stack frame)
        for (int i = threadIdx.x; i < 4 * args.n_tile; i += 192) {
            const int which = i / args.n_tile;
            const float* p = which == 0 ? args.epi.bias : (which == 1 ? args.epi.ln_g : (which == 2 ? args.epi.ln_b : args.epi.head_w));
            sbias[i] = p != nullptr ? p[ntile * args.n_tile + i % args.n_tile] : 0.f;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (csize == 2) cluster_sync_all();   // peer barriers must be initialised before any multicast / remote arrive
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) tl2_mark(args, 1);

    if (warp == 0) {
        // ---------------- TMA producer (one elected lane; see the note on elect.sync below) ----------------
        if (elect_one()) {
            const uint8_t* wtiles = reinterpret_cast<const uint8_t*>(args.w) +
                                    static_cast<size_t>(ntile) * n_btiles * btile_bytes;
            int sa = 0, sb = 0;
            uint32_t pa = 1, pb = 1;   // parity to wait on the EMPTY barriers (first lap passes)
            if (args.resident) {
                mbar_expect_tx(w_full, n_btiles * btile_bytes);
                for (int i = 0; i < n_btiles; ++i)
                    bulk_load_1d(smem_b + i * btile_bytes, wtiles + static_cast<size_t>(i) * btile_bytes, btile_bytes, w_full);
            }
            int item = 0;
            const int half_bytes = btile_bytes / 2;
            for (int p = first; p < n_groups; p += stride, ++item) {
                const int idx = p * csize + crank;
                const bool valid = idx < args.n_work;
                const int b = valid ? idx / args.groups_t : 0;
                const int tile0 = valid ? (idx - b * args.groups_t) * args.rpp : args.groups_t * args.rpp;
                const uint8_t* wp = wtiles;
                for (int c = 0; c < args.n_chunks; ++c) {
                    for (int r = 0; r < args.rpp; ++r) {
                        mbar_wait(&empty_a[sa], pa, args.err_flag, 201);
                        mbar_expect_tx(&full_a[sa], panel_bytes);
                        tma_load_3d(smem_a + sa * panel_bytes, &tmap_a, &full_a[sa], c * args.chunk_k,
                                    (tile0 + r) * kTileM - args.halo_lo, b);
                        if (c == 0 && r == 0 && item < 7) tl2_mark(args, 8 + item * 8 + 7);
                        if (++sa == args.a_slots) { sa = 0; pa ^= 1; }
                    }
                    if (!args.resident) {
                        for (int tap = 0; tap < args.n_taps; ++tap) {
                            mbar_wait(&empty_b[sb], pb, args.err_flag, 202);
                            mbar_expect_tx(&full_b[sb], btile_bytes);
                            if (csize == 2)
                                bulk_load_1d_multicast(smem_b + sb * btile_bytes + crank * half_bytes, wp + crank * half_bytes,
                                                       half_bytes, &full_b[sb], cmask);
                            else
                                bulk_load_1d(smem_b + sb * btile_bytes, wp, btile_bytes, &full_b[sb]);
                            wp += btile_bytes;
                            if (++sb == args.b_stages) { sb = 0; pb ^= 1; }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // elect.sync (not `lane == 0`): ptxas then knows exactly one lane is active and feeds the
        // UTCHMMA uniform-register operands with plain R2UR moves; with a threadIdx-derived predicate it
        // wraps EVERY tcgen05.mma in an ELECT/R2UR.BROADCAST/VOTEU waterfall loop (~260 cycles each,
        // profiles/r01_s10_timeline_v2.txt)
        if (elect_one()) {
            // ---------------- MMA issuer ----------------
            const int nsub_cols = args.n_tile / args.n_sub;
            const uint32_t idesc = umma_idesc_f16(kTileM, nsub_cols);
            const int cls = ntile >= args.class_split ? 1 : 0;
            const int ksteps = args.chunk_k >> 4;
            const uint32_t row_u = row_bytes >> 4;                       // descriptor address units (16 B)
            // descriptor high word: SBO (8-row atom pitch), version 1, swizzle mode
            const uint32_t desc_hi = ((8u * row_bytes) >> 4) | (1u << 14) | ((row_bytes == 128 ? 2u : 4u) << 29);
            const uint32_t lo_flag = 1u << 16;                           // LBO field (unused for swizzled K-major)
            const uint32_t a_lo0 = ((smem_u32(smem_a) & 0x3FFFFu) >> 4) + args.shift0[cls] * row_u;
            const uint32_t b_lo0 = (smem_u32(smem_b) & 0x3FFFFu) >> 4;
            const uint32_t panel_u = panel_bytes >> 4, btile_u = btile_bytes >> 4;
            const uint32_t tap_u = args.step[cls] * row_u;               // may be "negative" (wraps, added mod 2^32)
            const uint32_t sub_u = (nsub_cols * row_bytes) >> 4;
            int sa = 0, sb = 0, buf = 0;
            uint32_t pa = 0, pb = 0;      // parity to wait on the FULL barriers
            uint32_t pe0 = 1, pe1 = 1;    // parity to wait on tmem_empty[0/1]
            const int buf_cols = args.rpp * args.n_tile;
            const bool fast_issue = args.issue_mode != 0 && args.rpp == 1 && args.n_sub == 1 && ksteps == 4;
            bool b_ready = false;
            if (args.resident) {
                mbar_wait(w_full, 0, args.err_flag, 207);
                tc_fence_after();
            }
            int tl_i = 0;
            for (int p = first; p < n_groups; p += stride, ++tl_i) {
                mbar_wait(&tmem_empty[buf], buf ? pe1 : pe0, args.err_flag, 203);
                tc_fence_after();
                if (tl_i < 7) tl2_mark(args, 8 + tl_i * 8 + 0);
                if (buf) pe1 ^= 1; else pe0 ^= 1;
                const uint32_t d_tmem = tmem_base + buf * buf_cols;
                uint32_t accumulate = 0;
                if (fast_issue) {
                    // One row tile, one N sub-tile, K = 64 per (chunk, tap) — every C >= 128 vocoder layer. The generic
                    // loops below cost ~115 SASS instructions per tap and ~250 cycles per tcgen05.mma for the issuing
                    // thread, twice the MMA's own time (profiles/r01_s57_issue_overhead.txt). Here the four K steps go out
                    // back to back (the 14-bit address field cannot carry: a view start + 96 B is still a shared-memory
                    // address) and, in mode 2, the barrier of the NEXT weight stage is tested before this stage's MMAs are
                    // issued, which takes the barrier read off the per-stage dependent chain.
                    const uint64_t hi = static_cast<uint64_t>(desc_hi) << 32;
                    long long wait_b = 0, wait_n = 0;   // debug timeline: cycles / stages this tile waited for weight tiles
                    for (int c = 0; c < args.n_chunks; ++c) {
                        mbar_wait(&full_a[sa], pa, args.err_flag, 204);
                        uint32_t al = ((a_lo0 + sa * panel_u) & 0x3FFFu) | lo_flag;
                        const int a_slot0 = sa;
                        if (++sa == args.a_slots) { sa = 0; pa ^= 1; }
                        tc_fence_after();
                        if (c == 0 && tl_i < 7) tl2_mark(args, 8 + tl_i * 8 + 1);
                        for (int tap = 0; tap < args.n_taps; ++tap) {
                            uint32_t bl;
                            int cur = 0;
                            if (args.resident) {
                                bl = ((b_lo0 + (c * args.n_taps + tap) * btile_u) & 0x3FFFu) | lo_flag;
                            } else {
                                if (!b_ready) {
                                    const long long w0 = args.timeline != nullptr ? clock64() : 0;
                                    mbar_wait(&full_b[sb], pb, args.err_flag, 205);
                                    if (args.timeline != nullptr) { wait_b += clock64() - w0; ++wait_n; }
                                }
                                tc_fence_after();
                                bl = ((b_lo0 + sb * btile_u) & 0x3FFFu) | lo_flag;
                                cur = sb;
                                if (++sb == args.b_stages) { sb = 0; pb ^= 1; }
                            }
                            umma_f16(d_tmem, hi | al, hi | bl, idesc, accumulate);
                            umma_f16(d_tmem, hi | (al + 2u), hi | (bl + 2u), idesc, 1u);
                            umma_f16(d_tmem, hi | (al + 4u), hi | (bl + 4u), idesc, 1u);
                            umma_f16(d_tmem, hi | (al + 6u), hi | (bl + 6u), idesc, 1u);
                            al = ((al + tap_u) & 0x3FFFu) | lo_flag;
                            accumulate = 1;
                            if (!args.resident) {
                                // look at the next stage while this stage's MMAs execute
                                b_ready = args.issue_mode >= 2 && mbar_test_wait(&full_b[sb], pb);
                                if (csize == 2) umma_commit_multicast(&empty_b[cur], cmask);
                                else umma_commit(&empty_b[cur]);
                            }
                        }
                        umma_commit(&empty_a[a_slot0]);
                    }
                    umma_commit(&tmem_full[buf]);
                    if (tl_i < 7) tl2_mark(args, 8 + tl_i * 8 + 2);
                    if (args.timeline != nullptr && blockIdx.x < 256 && tl_i < 7)
                        args.timeline[blockIdx.x * 128 + 64 + tl_i * 8 + 7] = wait_b * 1000 + wait_n;
                    if (args.acc_bufs == 2) buf ^= 1;
                    continue;
                }
                for (int c = 0; c < args.n_chunks; ++c) {
                    uint32_t a_lo[4];
                    int a_slot[4];
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        if (r < args.rpp) {
                            mbar_wait(&full_a[sa], pa, args.err_flag, 204);
                            a_lo[r] = a_lo0 + sa * panel_u;
                            a_slot[r] = sa;
                            if (++sa == args.a_slots) { sa = 0; pa ^= 1; }
                        }
                    }
                    tc_fence_after();
                    if (c == 0 && tl_i < 7) tl2_mark(args, 8 + tl_i * 8 + 1);
                    for (int tap = 0; tap < args.n_taps; ++tap) {
                        uint32_t b_lo;
                        if (args.resident) {
                            b_lo = b_lo0 + (c * args.n_taps + tap) * btile_u;
                        } else {
                            mbar_wait(&full_b[sb], pb, args.err_flag, 205);
                            tc_fence_after();
                            b_lo = b_lo0 + sb * btile_u;
                        }
#pragma unroll
                        for (int r = 0; r < 4; ++r) {
                            if (r < args.rpp) {
                                for (int s = 0; s < args.n_sub; ++s) {
#pragma unroll
                                    for (int k = 0; k < 4; ++k) {
                                        if (k < ksteps) {
                                            const uint64_t ad = (static_cast<uint64_t>(desc_hi) << 32) | ((a_lo[r] + 2 * k) & 0x3FFFu) | lo_flag;
                                            const uint64_t bd = (static_cast<uint64_t>(desc_hi) << 32) | ((b_lo + s * sub_u + 2 * k) & 0x3FFFu) | lo_flag;
                                            umma_f16(d_tmem + r * args.n_tile + s * nsub_cols, ad, bd, idesc,
                                                     accumulate | static_cast<uint32_t>(k));
                                        }
                                    }
                                }
                                a_lo[r] += tap_u;
                            }
                        }
                        accumulate = 1;
                        if (!args.resident) {
                            if (csize == 2) umma_commit_multicast(&empty_b[sb], cmask);
                            else umma_commit(&empty_b[sb]);
                            if (++sb == args.b_stages) { sb = 0; pb ^= 1; }
                        }
                    }
#pragma unroll
                    for (int r = 0; r < 4; ++r)
                        if (r < args.rpp) umma_commit(&empty_a[a_slot[r]]);
                }
                umma_commit(&tmem_full[buf]);
                if (tl_i < 7) tl2_mark(args, 8 + tl_i * 8 + 2);
                if (args.acc_bufs == 2) buf ^= 1;
            }
        }
    } else {
        // ---------------- epilogue: 4 warps x 32 lanes = 128 rows ----------------
        const int q = warp & 3;
        int buf = 0;
        uint32_t pf0 = 0, pf1 = 0;
        const int buf_cols = args.rpp * args.n_tile;
        const int n_base = ntile * args.n_tile;
        int tl_i = 0;
        const bool tl_on = threadIdx.x == 64;
        uint8_t* stage = smem_stage + q * 2048;
        uint8_t* stage_in = args.stage2 ? smem_stage + 4 * 2048 + args.sbias_bytes + q * 2048 : nullptr;
        // lean path: the first residual / MRF chunk of the NEXT work item is requested before this
        // item's accumulator is waited for, so its latency hides behind a whole tile
        constexpr bool kLean = kEpi != 0 && kEpi != 8;
        constexpr bool kMrf = kEpi == 2;
        constexpr bool kAct = kEpi == 3;
        LeanPrefetch<kMrf> pre_cur, pre_nxt;
        // kEpi >= 3: nothing but lens[b] travels a tile ahead in registers
        auto prefetch = [&](const RowIO& io, long row0, bool on, LeanPrefetch<kMrf>& p, int pb) {
            if (kEpi >= 3) p.len_rows = (on && args.epi.lens != nullptr) ? __ldg(args.epi.lens + pb) * args.epi.len_mul : 0x7fffffff;
            else lean_prefetch(args.epi, io, row0, n_base, on, p, pb);
        };
        constexpr bool kTmaIn = kEpi >= 4 && kEpi <= 7;
        constexpr bool kMrfIn = kEpi == 6 || kEpi == 7;
        constexpr bool kStoreMrf = kEpi == 5 || kEpi == 6;
        TmaInState in_st{smem_stage + 4 * 2048 + args.sbias_bytes + q * (args.in_ring * (kMrfIn ? 4096 : 2048)), in_bar + q * 2,
                         args.in_ring, 0};
        if (kLean && first < n_groups) {
            const int idx0 = first * csize + crank;
            const bool v0 = idx0 < args.n_work;
            const int b0 = v0 ? idx0 / args.groups_t : 0;
            const int w0 = (v0 ? (idx0 - b0 * args.groups_t) * args.rpp : args.groups_t * args.rpp) * kTileM + q * 32;
            RowIO io{stage, lane, min(32, max(0, args.T - w0))};
            prefetch(io, static_cast<long>(b0) * args.T + w0, v0, pre_cur, b0);
            // TMA-in: the first `ring` chunks of the first tile (a dummy item's rows are out of range: zero-filled)
            if (kTmaIn && elect_one())
                for (int c = 0; c < args.in_ring; ++c)
                    tma_in_issue<kMrfIn>(in_st, c, &tmap_res, &tmap_mrf, n_base + c * 32, w0, b0);
        }
        for (int p = first; p < n_groups; p += stride, ++tl_i) {
            const int idx = p * csize + crank;
            const bool valid = idx < args.n_work;
            const int b = valid ? idx / args.groups_t : 0;
            const int tile0 = valid ? (idx - b * args.groups_t) * args.rpp : args.groups_t * args.rpp;
            const uint32_t par = buf ? pf1 : pf0;
            if (tl_on && tl_i < 7) tl2_mark(args, 8 + tl_i * 8 + 3);
            for (int r = 0; r < args.rpp; ++r) {
                const int t = (tile0 + r) * kTileM + q * 32 + lane;
                TmemAcc acc{tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * buf_cols + r * args.n_tile};
                auto wait_acc = [&] {
                    if (r == 0) {
                        mbar_wait(&tmem_full[buf], par, args.err_flag, 206);
                        tc_fence_after();
                        if (tl_on && tl_i < 7) tl2_mark(args, 8 + tl_i * 8 + 4);
                    }
                };
                auto drained = [&] {
                    if (r == args.rpp - 1) {
                        // every TMEM read of this warp has completed (tcgen05.wait::ld in acc.load)
                        tc_fence_before();
                        __syncwarp();
                        if (elect_one()) mbar_arrive(&tmem_empty[buf]);
                        if (tl_on && tl_i < 7) tl2_mark(args, 8 + tl_i * 8 + 5);
                    }
                };
                if (kLean) {
                    // next (item, r): either the next row tile of this item or the first of the next item
                    const bool last_r = r == args.rpp - 1;
                    const int nidx = last_r ? (p + stride) * csize + crank : idx;
                    const bool nvalid = nidx < args.n_work && (!last_r || p + stride < n_groups);
                    const int nb = nvalid ? nidx / args.groups_t : 0;
                    const int nw0 = ((nvalid ? (nidx - nb * args.groups_t) * args.rpp : args.groups_t * args.rpp) +
                                     (last_r ? 0 : r + 1)) * kTileM + q * 32;
                    RowIO nio{stage, lane, min(32, max(0, args.T - nw0))};
                    prefetch(nio, static_cast<long>(nb) * args.T + nw0, nvalid, pre_nxt, nb);
                    long long* dbg = (tl_on && args.timeline != nullptr && blockIdx.x < 256 && tl_i < 7)
                                         ? args.timeline + blockIdx.x * 128 + 64 + tl_i * 8 : nullptr;
                    if constexpr (kTmaIn)
                        run_epilogue_tma<kMrfIn, kStoreMrf, !kMrfIn>(args.epi, acc.taddr, b, t, n_base, args.n_tile, wait_acc, drained, stage,
                                                                     in_st, pre_cur.len_rows, smem_u32(sbias), kStoreMrf ? &tmap_mrf : &tmap_act,
                                                                     &tmap_res, &tmap_mrf, p + stride < n_groups, nb, nw0, args.err_flag, dbg);
                    else if constexpr (kAct)
                        run_epilogue_act<false, true>(args.epi, acc.taddr, b, t, n_base, args.n_tile, wait_acc, drained, stage, stage_in,
                                                      pre_cur, smem_u32(sbias), &tmap_act, dbg);
                    else
                    run_epilogue_lean<kMrf, true, !(kMrf && kMinBlocks >= 2)>(args.epi, acc, b, t, n_base, args.n_tile, wait_acc, drained, stage, pre_cur,
                                            0x7fffffff, smem_u32(sbias), (args.tma_out & 1) ? &tmap_raw : nullptr,
                                            (args.tma_out & 2) ? &tmap_act : nullptr, (args.tma_out & 4) ? &tmap_mrf : nullptr, stage_in, dbg);
                    pre_cur = pre_nxt;
                } else {
                    long long* dbg = (tl_on && args.timeline != nullptr && blockIdx.x < 256 && tl_i < 7)
                                         ? args.timeline + blockIdx.x * 128 + 64 + tl_i * 8 : nullptr;
                    run_epilogue<kEpi == 8 ? 1 : 0>(args.epi, acc, b, t, t < args.T, n_base, args.n_tile, wait_acc, drained, stage, dbg,
                                                    args.params_smem ? smem_u32(sbias) : 0u);
                }
            }
            if (tl_on && tl_i < 7) tl2_mark(args, 8 + tl_i * 8 + 6);
            if (buf) pf1 ^= 1; else pf0 ^= 1;
            if (args.acc_bufs == 2) buf ^= 1;
        }
    }
    if (warp >= 2 && args.tma_out != 0 && elect_one()) tma_store_wait_all0();   // smem must outlive the bulk stores
    tc_fence_before();
    __syncthreads();
    if (csize == 2) cluster_sync_all();   // no CTA may exit while its peer can still multicast into it
    if (warp == 2) tmem_dealloc<kTmemCols>(tmem_base);
}

// ------------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled2)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                     const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                     const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled2 get_encode_fn2() {
    static PFN_encodeTiled2 fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled2>(p);
    }
    return fn;
}

// 3-D tensor map over channel-last activations [B][T][cin] (row pitch ld_in) with a box of
// chunk_k channels x rows_panel rows; out-of-range rows are zero-filled (= conv zero padding).
// Tensor maps are pure functions of (pointer, geometry): memoise them, the same workspace buffers
// come back every step (host time matters once a step is ~10^3 launches of ~50 us).
int get_act_tensor_map(const __half* in, int ld_in, int B, int T, int cin, int chunk_k, int rows_panel,
                       const CUtensorMap** out, int t_stride) {
    // t_stride (0 = T): rows between two utterances in memory when the tensor holds more rows per utterance than the T
    // this launch looks at (the generator runs each chunk of a padded batch at the chunk's own longest length)
    if (t_stride <= 0) t_stride = T;
    PFN_encodeTiled2 enc = get_encode_fn2();
    TTSB_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled driver entry point not available");
    TTSB_REQUIRE((reinterpret_cast<uintptr_t>(in) & 15) == 0 && (ld_in % 8) == 0, "input alignment");
    struct TmKey {
        const void* p; int ld, B, T, cin, ck, rows, ts;
        bool operator==(const TmKey& o) const {
            return p == o.p && ld == o.ld && B == o.B && T == o.T && cin == o.cin && ck == o.ck && rows == o.rows && ts == o.ts;
        }
    };
    static thread_local std::vector<std::pair<TmKey, CUtensorMap>> cache;
    const TmKey key{in, ld_in, B, T, cin, chunk_k, rows_panel, t_stride};
    for (auto& kv : cache)
        if (kv.first == key) { *out = &kv.second; return 0; }
    CUtensorMap tm_new;
    cuuint64_t dims[3] = {static_cast<cuuint64_t>(cin), static_cast<cuuint64_t>(T), static_cast<cuuint64_t>(B)};
    cuuint64_t strides[2] = {static_cast<cuuint64_t>(ld_in) * 2, static_cast<cuuint64_t>(t_stride) * ld_in * 2};
    cuuint32_t box[3] = {static_cast<cuuint32_t>(chunk_k), static_cast<cuuint32_t>(rows_panel), 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&tm_new, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<__half*>(in), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE,
                     chunk_k == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    TTSB_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with code " + std::to_string(static_cast<int>(r)));
    if (cache.size() >= 256) cache.clear();   // pointers handed out earlier are not kept across calls
    cache.emplace_back(key, tm_new);
    *out = &cache.back().second;
    return 0;
}

int num_sms() {
    static int n[kMaxDevices] = {};
    const int dev = current_device();
    if (!n[dev]) {
        cudaDeviceGetAttribute(&n[dev], cudaDevAttrMultiProcessorCount, dev);
        if (n[dev] <= 0) n[dev] = 148;
    }
    return n[dev];
}

struct OutMaps {
    CUtensorMap raw, act, mrf, res;
    int epi_kind;   // kEpi of the launch, chosen by conv_forward_tc2
};

template <int kCols, int kMinBlocks, int kEpi>
static int launch_two_impl(const CUtensorMap& tm, const ConvTc2Args& a, int grid, size_t smem, cudaStream_t s, const OutMaps& om) {
    static PerDeviceOnce configured;
    if (!configured.here()) {
        TTSB_CHECK_CUDA(cudaFuncSetAttribute(conv_tc2_kernel<kCols, kMinBlocks, kEpi>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
        configured.here() = true;
    }
    if (a.cluster == 2) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(192);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        TTSB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, conv_tc2_kernel<kCols, kMinBlocks, kEpi>, tm, a, om.raw, om.act, om.mrf, om.res));
    } else {
        conv_tc2_kernel<kCols, kMinBlocks, kEpi><<<grid, 192, smem, s>>>(tm, a, om.raw, om.act, om.mrf, om.res);
    }
    count_launch();
    TTSB_CHECK_CUDA(cudaGetLastError());
    return 0;
}

static bool host_epi_is_lean(const EpiParams& e) {
    return e.ln_g == nullptr && e.head_w == nullptr && e.out_f32_t == nullptr && e.out_f32 == nullptr &&
           e.act_tanh == 0 && e.pre_ln_relu == 0;
}
template <int kCols, int kMinBlocks>
static int launch_two(const CUtensorMap& tm, const ConvTc2Args& a, int grid, size_t smem, cudaStream_t s, const OutMaps& om) {
    // the act-only / TMA-in epilogues are instantiated where the vocoder's C >= 128 layers and up-samplers run
    // (>= 128 TMEM columns, one or two CTAs per SM); conv_forward_tc2 only picks them there
    constexpr bool kBig = kCols >= 128 && kMinBlocks <= 2;
    switch (om.epi_kind) {
        case 0: return launch_two_impl<kCols, kMinBlocks, 0>(tm, a, grid, smem, s, om);
        case 8: return launch_two_impl<kCols, kMinBlocks, 8>(tm, a, grid, smem, s, om);
        case 2: return launch_two_impl<kCols, kMinBlocks, 2>(tm, a, grid, smem, s, om);
        case 3: if constexpr (kBig) return launch_two_impl<kCols, kMinBlocks, 3>(tm, a, grid, smem, s, om); break;
        case 4: if constexpr (kBig) return launch_two_impl<kCols, kMinBlocks, 4>(tm, a, grid, smem, s, om); break;
        case 5: if constexpr (kBig) return launch_two_impl<kCols, kMinBlocks, 5>(tm, a, grid, smem, s, om); break;
        case 6: if constexpr (kBig) return launch_two_impl<kCols, kMinBlocks, 6>(tm, a, grid, smem, s, om); break;
        case 7: if constexpr (kBig) return launch_two_impl<kCols, kMinBlocks, 7>(tm, a, grid, smem, s, om); break;
        default: break;
    }
    TTSB_REQUIRE(om.epi_kind == 1, "epilogue kind not instantiated for this tile shape");
    return launch_two_impl<kCols, kMinBlocks, 1>(tm, a, grid, smem, s, om);
}

int conv_forward_tc2(const ConvLayer& L, const ConvRuntime& rt, const __half* in, int ld_in, int B,
                     int T, const EpiParams& epi, cudaStream_t stream) {
    TTSB_REQUIRE(ld_in >= L.cin, "input row pitch smaller than layer Cin");
    // tap offsets must be an arithmetic sequence per class (true for conv, transposed conv, linear)
    int step[2] = {0, 0};
    for (int c = 0; c < 2; ++c) {
        step[c] = L.n_taps > 1 ? L.tap_off[c][1] - L.tap_off[c][0] : 0;
        for (int i = 1; i < L.n_taps; ++i)
            TTSB_REQUIRE(L.tap_off[c][i] - L.tap_off[c][i - 1] == step[c], "tap offsets must be equally spaced");
    }

    const CUtensorMap* tmp = nullptr;
    TTSB_PROPAGATE(get_act_tensor_map(in, ld_in, B, T, L.cin, L.chunk_k, L.rows_panel, &tmp, rt.in_t_stride));
    const CUtensorMap tm = *tmp;      // by value: the cache may reallocate on the next lookup

    ConvTc2Args a;
    a.B = B; a.T = T;
    a.tiles_t = ceil_div(T, kTileM);
    a.rpp = L.rpp;
    a.groups_t = ceil_div(a.tiles_t, a.rpp);
    a.n_work = B * a.groups_t;
    a.resident = L.resident;
    a.n_tiles_n = L.n_tiles();
    a.n_chunks = L.n_chunks; a.n_taps = L.n_taps; a.chunk_k = L.chunk_k;
    a.rows_panel = L.rows_panel;
    a.n_tile = L.n_tile; a.n_sub = L.n_sub;
    a.a_slots = L.a_slots2; a.b_stages = L.b_stages2;
    a.acc_bufs = L.acc_bufs;
    a.a_lanes = 1; a.b_lanes = 1;
    for (int d = 2; d <= 4; ++d) if (a.a_slots % d == 0) a.a_lanes = d;
    for (int d = 2; d <= 8; ++d) if (a.b_stages % d == 0) a.b_lanes = d;
    a.class_split = L.class_split;
    a.halo_lo = L.halo_lo;
    for (int c = 0; c < 2; ++c) {
        a.shift0[c] = L.halo_lo + L.tap_off[c][0];
        a.step[c] = step[c];
    }
    a.w = L.w_packed; a.err_flag = rt.err_flag; a.epi = epi;
    a.timeline = rt.timeline;
    static const int epi_debug = getenv("TTSB_EPI_DEBUG") ? atoi(getenv("TTSB_EPI_DEBUG")) : 0;
    a.epi.debug = epi_debug;
    static const int issue_mode = getenv("TTSB_ISSUE") ? atoi(getenv("TTSB_ISSUE")) : 2;
    a.issue_mode = issue_mode;

    // lean epilogue: outputs leave through TMA stores of 32-row x 32-column blocks (64-byte swizzle = the staging
    // tiles' XOR pattern); the maps are the activation map function with a 32 x 32 box over [B][T][n_total]
    OutMaps om;
    om.raw = om.act = om.mrf = tm;
    a.tma_out = 0;
    static const int want_tma_out = getenv("TTSB_TMA_OUT") ? atoi(getenv("TTSB_TMA_OUT")) : 1;
    if (want_tma_out && host_epi_is_lean(epi)) {
        struct { __half* p; int ld; CUtensorMap* m; int bit; } outs[3] = {
            {epi.out_raw, epi.ld_raw, &om.raw, 1}, {epi.out_act, epi.ld_act, &om.act, 2}, {epi.mrf_buf, L.n_total, &om.mrf, 4}};
        for (auto& o : outs) {
            if (o.p == nullptr || (reinterpret_cast<uintptr_t>(o.p) & 15) != 0 || o.ld % 8 != 0 || o.ld < L.n_total) continue;
            const CUtensorMap* t = nullptr;
            TTSB_PROPAGATE(get_act_tensor_map(o.p, o.ld, B, T, L.n_total, 32, 32, &t));
            *o.m = *t;
            a.tma_out |= o.bit;
        }
    }
    // general epilogue with LayerNorm: stage the per-column parameters in shared memory when the plan leaves room
    // (2 KB are always reserved for the lean epilogue's bias tile)
    a.params_smem = 0;
    size_t smem_bytes = L.smem_bytes2;
    if (!host_epi_is_lean(epi) && epi.ln_g != nullptr) {
        const size_t need = 4 * static_cast<size_t>(L.n_tile) * sizeof(float);
        const size_t extra = need > 2048 ? need - 2048 : 0;
        const size_t limit = 233472 / L.occ2 - 1024;
        if (L.smem_bytes2 + extra <= limit && L.smem_bytes2 + extra <= 232448) {
            a.params_smem = 1;
            smem_bytes += extra;
        }
    }
    // epilogue kind (conv_tc2_kernel's kEpi)
    a.stage2 = 0;
    a.in_ring = 0;
    a.sbias_bytes = static_cast<int>(2048 + (smem_bytes - L.smem_bytes2));
    om.res = tm;
    om.epi_kind = !host_epi_is_lean(epi) ? (epi.ln_g != nullptr ? 8 : 0) : (epi.mrf_mode == MRF_NONE ? 1 : 2);
    static const int want_act = getenv("TTSB_EPI_ACT") ? atoi(getenv("TTSB_EPI_ACT")) : 1;
    static const int want_tma_in = getenv("TTSB_EPI_TMA_IN") ? atoi(getenv("TTSB_EPI_TMA_IN")) : 1;
    const bool big3 = L.tmem_cols2 >= 128 && L.occ2 <= 2 && L.n_tile % 32 == 0;
    const bool big = big3 && L.rpp == 1 && L.n_tile % 64 == 0;
    if (om.epi_kind == 1 && want_act && big3 && epi.residual == nullptr) {
        if (epi.out_raw == nullptr && epi.out_act != nullptr && (a.tma_out & 2)) {
            om.epi_kind = 3;
        } else if (epi.out_act == nullptr && epi.out_raw != nullptr && (a.tma_out & 1)) {
            // a raw-only output is the activated one with slope 1: max(v, 1 * v) == v exactly
            a.epi.out_act = epi.out_raw; a.epi.ld_act = epi.ld_raw; a.epi.act_slope = 1.f; a.epi.out_raw = nullptr;
            om.act = om.raw;
            a.tma_out = 2;
            om.epi_kind = 3;
        }
    }
    if (om.epi_kind != 0 && om.epi_kind != 8 && om.epi_kind != 3 && want_tma_in && big && epi.residual != nullptr && epi.out_raw == nullptr &&
        (reinterpret_cast<uintptr_t>(epi.residual) & 15) == 0 && epi.ld_res % 8 == 0 && epi.ld_res >= L.n_total) {
        int kind = 0;
        switch (epi.mrf_mode) {
            case MRF_NONE: kind = (epi.out_act != nullptr && (a.tma_out & 2)) ? 4 : 0; break;
            case MRF_FIRST: kind = (a.tma_out & 4) ? 5 : 0; break;
            case MRF_ADD: kind = (a.tma_out & 4) ? 6 : 0; break;
            case MRF_LAST: kind = (epi.out_act != nullptr && (a.tma_out & 2) && (a.tma_out & 4)) ? 7 : 0; break;
        }
        // input tiles: 4 warps x ring x (residual [, MRF]) x 2 KB, two chunks of look-ahead if they fit, else one
        const size_t limit = std::min<size_t>(232448, 233472 / L.occ2 - 1024);
        const size_t per_ring = 4 * 2048 * (kind >= 6 ? 2 : 1);
        static const int max_ring = getenv("TTSB_EPI_RING") ? atoi(getenv("TTSB_EPI_RING")) : 2;
        int ring = 0;
        for (int r = std::min(2, max_ring); r >= 1 && ring == 0; --r)
            if (smem_bytes + r * per_ring <= limit) ring = r;
        // no room: give up the fourth weight stage (the layers whose plan has three run the same main loop)
        static const int may_trade = getenv("TTSB_EPI_TRADE_B") ? atoi(getenv("TTSB_EPI_TRADE_B")) : 1;
        if (kind != 0 && ring == 0 && may_trade && !L.resident && a.b_stages >= 4) {
            const size_t btile = static_cast<size_t>(L.n_tile) * L.chunk_k * 2;
            for (int r = std::min(2, max_ring); r >= 1 && ring == 0; --r)
                if (smem_bytes - btile + r * per_ring <= limit) ring = r;
            if (ring != 0) {
                a.b_stages -= 1;
                smem_bytes -= btile;
            }
        }
        if (kind != 0 && ring != 0) {
            const CUtensorMap* t = nullptr;
            TTSB_PROPAGATE(get_act_tensor_map(epi.residual, epi.ld_res, B, T, L.n_total, 32, 32, &t));
            om.res = *t;
            om.epi_kind = kind;
            a.in_ring = ring;
            smem_bytes += ring * per_ring;
        }
    }
    if (om.epi_kind == 1 || om.epi_kind == 2) {
        static const int want_stage2 = getenv("TTSB_STAGE2") ? atoi(getenv("TTSB_STAGE2")) : 1;
        const size_t limit = std::min<size_t>(232448, 233472 / L.occ2 - 1024);
        if (want_stage2 && host_epi_is_lean(epi) && (epi.residual != nullptr || epi.mrf_mode != MRF_NONE) && a.tma_out != 0 &&
            smem_bytes + 8192 <= limit) {
            a.stage2 = 1;
            smem_bytes += 8192;
        }
    }
    static const int verbose = getenv("TTSB_EPI_VERBOSE") ? atoi(getenv("TTSB_EPI_VERBOSE")) : 0;
    if (verbose)
        fprintf(stderr, "conv_tc2: cin %d n %d taps %d rows_panel %d occ %d a_slots %d b_stages %d resident %d -> epilogue kind %d ring %d stage2 %d smem %zu\n",
                L.cin, L.n_total, L.n_taps, L.rows_panel, L.occ2, L.a_slots2, L.b_stages2, L.resident, om.epi_kind, a.in_ring, a.stage2,
                smem_bytes);
    // CTA pairs share streamed weight tiles through TMA multicast (halves the L2->SM weight traffic that
    // bounds the C >= 128 layers); needs at least two work items per N tile
    static const int want_cluster = getenv("TTSB_CLUSTER") ? atoi(getenv("TTSB_CLUSTER")) : 2;
    a.cluster = (!L.resident && want_cluster == 2 && a.n_work >= 2 && (L.n_tile / 2) % 8 == 0) ? 2 : 1;
    // persistent grid: `occ` CTAs per SM, a multiple of (cluster size x number of N tiles)
    int ctas = num_sms() * L.occ2;
    const long total = static_cast<long>(a.n_work) * a.n_tiles_n;
    if (ctas > total) ctas = static_cast<int>(total);
    const int unit = a.n_tiles_n * a.cluster;
    ctas = (ctas / unit) * unit;
    if (ctas < unit) ctas = unit;
    // the register cap follows the planned CTAs per SM (1: 255, 2: 168, 3: 112 registers per thread)
    if (L.occ2 >= 3) {
        switch (L.tmem_cols2) {
            case 32: return launch_two<32, 3>(tm, a, ctas, smem_bytes, stream, om);
            case 64: return launch_two<64, 3>(tm, a, ctas, smem_bytes, stream, om);
            case 128: return launch_two<128, 3>(tm, a, ctas, smem_bytes, stream, om);
        }
        TTSB_REQUIRE(false, "occupancy 3 needs <= 128 TMEM columns");
    }
    if (L.occ2 == 2) {
        switch (L.tmem_cols2) {
            case 32: return launch_two<32, 2>(tm, a, ctas, smem_bytes, stream, om);
            case 64: return launch_two<64, 2>(tm, a, ctas, smem_bytes, stream, om);
            case 128: return launch_two<128, 2>(tm, a, ctas, smem_bytes, stream, om);
            case 256: return launch_two<256, 2>(tm, a, ctas, smem_bytes, stream, om);
        }
        TTSB_REQUIRE(false, "occupancy 2 needs <= 256 TMEM columns");
    }
    switch (L.tmem_cols2) {
        case 32: return launch_two<32, 1>(tm, a, ctas, smem_bytes, stream, om);
        case 64: return launch_two<64, 1>(tm, a, ctas, smem_bytes, stream, om);
        case 128: return launch_two<128, 1>(tm, a, ctas, smem_bytes, stream, om);
        case 256: return launch_two<256, 1>(tm, a, ctas, smem_bytes, stream, om);
        case 512: return launch_two<512, 1>(tm, a, ctas, smem_bytes, stream, om);
    }
    TTSB_REQUIRE(false, "bad tmem_cols");
    return 1;
}

}  // namespace ttsb
