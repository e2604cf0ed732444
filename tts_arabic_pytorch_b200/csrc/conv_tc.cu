// tcgen05 implicit-GEMM kernel for "row GEMM with taps" (see conv.cuh).
//
// One CTA = 128 output rows (time positions of one utterance) x n_tile output columns.
//   M = 128 rows            -> TMEM lanes
//   N = n_tile columns      -> TMEM columns (fp32 accumulators)
//   K = n_taps * Cin        -> walked as (channel chunk of 64|32) x tap
//
// A operand (activations): for every channel chunk ONE halo panel of `rows_panel` consecutive
// rows x chunk_k channels is TMA-loaded (3-D tensor map over [B][T][C]; rows outside [0,T) are
// zero-filled by the TMA unit == nn.Conv1d zero padding). Each tap's A tile is a row-shifted
// VIEW of that panel: the UMMA shared-memory descriptor start address is advanced by
// shift*row_bytes, so a k=11 dilated conv reads its activations from L2 once, not 11 times.
// B operand (weights): pre-swizzled on the host into n_tile x chunk_k tiles, streamed with 1-D
// bulk copies through a ring of `b_stages` buffers.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = MMA issuer (one thread),
// warps 2..5 = epilogue (TMEM -> registers -> fused epilogue -> HBM), warp 2 owns TMEM alloc.
//
// Replaces (reference): cuDNN conv / cuBLAS GEMM calls behind nn.Conv1d, nn.ConvTranspose1d and
// nn.Linear on the hot path (vocoder/hifigan/models.py:46-53,111-127, transformer.py:83-88,
// 122,148, model.py:54-57,406).
#include <cstdlib>
#include "conv.cuh"

namespace ttsb {

struct ConvTcArgs {
    int B, T;
    int n_chunks, n_taps;
    int chunk_k;       // 64 or 32
    int rows_panel;    // multiple of 8
    int halo_lo;
    int n_tile, n_sub;
    int a_slots, b_stages;
    int class_split;
    int desc_mode;
    int debug_flags;   // timing decomposition only: 1 = skip epilogue memory traffic, 2 = shrink weight loads
    int tap_off[2][kMaxTaps];
    const __half* w;
    int* err_flag;
    long long* timeline;   // debug: 64 clock64() slots per CTA for the first 256 CTAs, or null
    EpiParams epi;
};

// K-major swizzled descriptor for a tile whose rows are `row_bytes` (128 or 64) apart.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t row_bytes, uint32_t base_off) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((saddr & 0x3FFFF) >> 4);
    d |= static_cast<uint64_t>(1) << 16;
    d |= static_cast<uint64_t>((8 * row_bytes) >> 4) << 32;  // SBO: 8-row swizzle atom pitch
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(base_off & 7) << 49;
    d |= static_cast<uint64_t>(row_bytes == 128 ? 2 : 4) << 61;  // SWIZZLE_128B / SWIZZLE_64B
    return d;
}

__device__ __forceinline__ void tl_mark(const ConvTcArgs& a, int slot) {
    if (a.timeline == nullptr) return;
    const int cta = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
    if (cta < 256 && slot < 64) a.timeline[cta * 64 + slot] = clock64();
}

template <int kTmemCols>
__global__ void __launch_bounds__(192, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ ConvTcArgs args) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_u32 = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw_u32 & 1023u)) & 1023u);

    const int row_bytes = args.chunk_k * 2;
    const int panel_bytes = args.rows_panel * row_bytes;
    const int btile_bytes = args.n_tile * row_bytes;
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + args.a_slots * panel_bytes;
    uint64_t* full_a = reinterpret_cast<uint64_t*>(smem_b + args.b_stages * btile_bytes);
    uint64_t* empty_a = full_a + args.a_slots;
    uint64_t* full_b = empty_a + args.a_slots;
    uint64_t* empty_b = full_b + args.b_stages;
    uint64_t* tmem_full = empty_b + args.b_stages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

    const bool per_tap = args.desc_mode == 3;  // one TMA tile per (chunk, tap), no shifted views
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int t0 = blockIdx.x * kTileM;
    const int ntile = blockIdx.y;
    const int b = blockIdx.z;

    if (warp == 0 && lane == 0) {
        tl_mark(args, 0);
        tma_prefetch_desc(&tmap_a);
        for (int i = 0; i < args.a_slots; ++i) { mbar_init(&full_a[i], 1); mbar_init(&empty_a[i], 1); }
        for (int i = 0; i < args.b_stages; ++i) { mbar_init(&full_b[i], 1); mbar_init(&empty_b[i], 1); }
        mbar_init(tmem_full, 1);
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc<kTmemCols>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) tl_mark(args, 1);

    if (warp == 0) {
        if (lane == 0) {
            // ---------------- TMA producer ----------------
            const __half* wbase = args.w + static_cast<size_t>(ntile) * args.n_chunks * args.n_taps *
                                               (static_cast<size_t>(args.n_tile) * args.chunk_k);
            const int cls = ntile >= args.class_split ? 1 : 0;
            for (int c = 0; c < args.n_chunks; ++c) {
                if (!per_tap) {
                    const int sa = c % args.a_slots, ua = c / args.a_slots;
                    mbar_wait(&empty_a[sa], (ua & 1) ^ 1, args.err_flag, 101);
                    mbar_expect_tx(&full_a[sa], panel_bytes);
                    tma_load_3d(smem_a + sa * panel_bytes, &tmap_a, &full_a[sa], c * args.chunk_k,
                                t0 - args.halo_lo, b);
                }
                for (int tap = 0; tap < args.n_taps; ++tap) {
                    const int i = c * args.n_taps + tap;
                    if (per_tap) {
                        const int sa = i % args.a_slots, ua = i / args.a_slots;
                        mbar_wait(&empty_a[sa], (ua & 1) ^ 1, args.err_flag, 101);
                        mbar_expect_tx(&full_a[sa], panel_bytes);
                        tma_load_3d(smem_a + sa * panel_bytes, &tmap_a, &full_a[sa], c * args.chunk_k,
                                    t0 + args.tap_off[cls][tap], b);
                    }
                    const int sb = i % args.b_stages, ub = i / args.b_stages;
                    mbar_wait(&empty_b[sb], (ub & 1) ^ 1, args.err_flag, 102);
                    const int bbytes = ((args.debug_flags & 2) && i >= args.b_stages) ? 16 : btile_bytes;
                    mbar_expect_tx(&full_b[sb], bbytes);
                    bulk_load_1d(smem_b + sb * btile_bytes,
                                 wbase + static_cast<size_t>(i) * args.n_tile * args.chunk_k,
                                 bbytes, &full_b[sb]);
                    if (i == 0) tl_mark(args, 2);
                }
            }
            tl_mark(args, 3);
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ---------------- MMA issuer ----------------
            const int nsub_cols = args.n_tile / args.n_sub;
            const uint32_t idesc = umma_idesc_f16(kTileM, nsub_cols);
            const int cls = ntile >= args.class_split ? 1 : 0;
            const int ksteps = args.chunk_k / 16;
            for (int c = 0; c < args.n_chunks; ++c) {
                int sa = c % args.a_slots, ua = c / args.a_slots;
                if (!per_tap) mbar_wait(&full_a[sa], ua & 1, args.err_flag, 103);
                for (int tap = 0; tap < args.n_taps; ++tap) {
                    const int i = c * args.n_taps + tap;
                    if (per_tap) {
                        sa = i % args.a_slots; ua = i / args.a_slots;
                        mbar_wait(&full_a[sa], ua & 1, args.err_flag, 103);
                    }
                    const int sb = i % args.b_stages, ub = i / args.b_stages;
                    mbar_wait(&full_b[sb], ub & 1, args.err_flag, 104);
                    tc_fence_after();
                    tl_mark(args, 8 + (i < 40 ? i : 40));
                    const int shift = per_tap ? 0 : args.halo_lo + args.tap_off[cls][tap];
                    const uint32_t a_addr = smem_u32(smem_a + sa * panel_bytes) + shift * row_bytes;
                    const uint32_t b_addr = smem_u32(smem_b + sb * btile_bytes);
                    for (int s = 0; s < args.n_sub; ++s) {
                        for (int k = 0; k < ksteps; ++k) {
                            const uint32_t aa = a_addr + k * 32;
                            const uint32_t bb = b_addr + s * nsub_cols * row_bytes + k * 32;
                            // Row-shifted views start off the swizzle-atom boundary. Measured on
                            // B200 (profiles/r01_s1_probe_conv.json): the UMMA unit applies the
                            // swizzle XOR to ABSOLUTE shared-memory address bits, so the view is
                            // correct with base_offset = 0 (desc_mode 0, the default) and WRONG
                            // with base_offset = (addr>>7)&7 (desc_mode 1/2, kept as probes).
                            // desc_mode 3 never shifts (per-tap TMA tiles, shift == 0).
                            const uint32_t boff = args.desc_mode == 1 ? ((aa >> 7) & 7u)
                                                  : args.desc_mode == 2 ? ((aa >> 7) & (row_bytes == 128 ? 7u : 3u))
                                                  : 0u;
                            umma_f16(tmem_base + s * nsub_cols, make_desc(aa, row_bytes, boff),
                                     make_desc(bb, row_bytes, 0), idesc, (i > 0 || k > 0) ? 1u : 0u);
                        }
                    }
                    umma_commit(&empty_b[sb]);
                    if (per_tap) umma_commit(&empty_a[sa]);
                }
                if (!per_tap) umma_commit(&empty_a[sa]);
            }
            umma_commit(tmem_full);
            tl_mark(args, 4);
        }
    } else {
        // ---------------- epilogue: 4 warps x 32 lanes = 128 rows ----------------
        const int q = warp & 3;
        const int t = t0 + q * 32 + lane;
        TmemAcc acc{tmem_base + (static_cast<uint32_t>(q * 32) << 16)};
        auto wait_acc = [&] {
            mbar_wait(tmem_full, 0, args.err_flag, 105);
            tc_fence_after();
            if (threadIdx.x == 64) tl_mark(args, 5);
        };
        if (args.debug_flags & 1) {
            wait_acc();
            float v[32];
            acc.load(0, v);
            if (v[0] == 1234.5678f && args.epi.out_raw) args.epi.out_raw[0] = __float2half(v[1]);
        } else {
            run_epilogue(args.epi, acc, b, t, t < args.T, ntile * args.n_tile, args.n_tile, wait_acc, [] {});
        }
    }
    if (threadIdx.x == 64) tl_mark(args, 6);
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc<kTmemCols>(tmem_base);
    if (threadIdx.x == 64) tl_mark(args, 7);
}

// ------------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(p);
    }
    return fn;
}

template <int kCols>
static int launch_one(const CUtensorMap& tm, const ConvTcArgs& a, dim3 grid, size_t smem, cudaStream_t s) {
    static PerDeviceOnce configured;
    if (!configured.here()) {
        TTSB_CHECK_CUDA(cudaFuncSetAttribute(conv_tc_kernel<kCols>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
        configured.here() = true;
    }
    conv_tc_kernel<kCols><<<grid, 192, smem, s>>>(tm, a);
    count_launch();
    TTSB_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int conv_forward_tc(const ConvLayer& L, const ConvRuntime& rt, const __half* in, int ld_in, int B,
                    int T, const EpiParams& epi, cudaStream_t stream) {
    PFN_encodeTiled enc = get_encode_fn();
    TTSB_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled driver entry point not available");
    TTSB_REQUIRE((reinterpret_cast<uintptr_t>(in) & 15) == 0 && (ld_in % 8) == 0, "input alignment");
    TTSB_REQUIRE(ld_in >= L.cin, "input row pitch smaller than layer Cin");

    // desc_mode 3 (per-tap TMA tiles): 128-row boxes, A ring as deep as the B ring.
    int rows_panel = L.rows_panel, a_slots = L.a_slots, b_stages = L.b_stages;
    size_t smem_bytes = L.smem_bytes;
    if (rt.desc_mode == 3) {
        rows_panel = kTileM;
        const size_t a_bytes = static_cast<size_t>(kTileM) * L.chunk_k * 2;
        const size_t b_bytes = static_cast<size_t>(L.n_tile) * L.chunk_k * 2;
        int s = static_cast<int>((232448 - 2048) / (a_bytes + b_bytes));
        s = s > 4 ? 4 : s;
        TTSB_REQUIRE(s >= 2, "per-tap mode does not fit in shared memory");
        a_slots = b_stages = s;
        smem_bytes = 1024 + s * (a_bytes + b_bytes) + (4 * s + 1) * 8 + 16;
    }

    CUtensorMap tm;
    cuuint64_t dims[3] = {static_cast<cuuint64_t>(L.cin), static_cast<cuuint64_t>(T), static_cast<cuuint64_t>(B)};
    cuuint64_t strides[2] = {static_cast<cuuint64_t>(ld_in) * 2, static_cast<cuuint64_t>(T) * ld_in * 2};
    cuuint32_t box[3] = {static_cast<cuuint32_t>(L.chunk_k), static_cast<cuuint32_t>(rows_panel), 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<__half*>(in), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE,
                     L.chunk_k == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    TTSB_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with code " + std::to_string(static_cast<int>(r)));

    ConvTcArgs a;
    a.B = B; a.T = T;
    a.n_chunks = L.n_chunks; a.n_taps = L.n_taps; a.chunk_k = L.chunk_k;
    a.rows_panel = rows_panel; a.halo_lo = L.halo_lo;
    a.n_tile = L.n_tile; a.n_sub = L.n_sub;
    a.a_slots = a_slots; a.b_stages = b_stages;
    a.class_split = L.class_split; a.desc_mode = rt.desc_mode;
    {
        static int dbg = getenv("TTSB_DEBUG_FLAGS") ? atoi(getenv("TTSB_DEBUG_FLAGS")) : 0;
        a.debug_flags = dbg;
    }
    for (int c = 0; c < 2; ++c)
        for (int i = 0; i < kMaxTaps; ++i) a.tap_off[c][i] = L.tap_off[c][i];
    a.w = L.w_packed; a.err_flag = rt.err_flag; a.epi = epi;
    a.timeline = rt.timeline;

    dim3 grid(ceil_div(T, kTileM), L.n_tiles(), B);
    switch (L.tmem_cols) {
        case 32: return launch_one<32>(tm, a, grid, smem_bytes, stream);
        case 64: return launch_one<64>(tm, a, grid, smem_bytes, stream);
        case 128: return launch_one<128>(tm, a, grid, smem_bytes, stream);
        case 256: return launch_one<256>(tm, a, grid, smem_bytes, stream);
        case 512: return launch_one<512>(tm, a, grid, smem_bytes, stream);
    }
    TTSB_REQUIRE(false, "bad tmem_cols");
    return 1;
}

}  // namespace ttsb
