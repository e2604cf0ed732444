// Layer construction (weight re-layout for UMMA), tile-shape selection, the SIMT check
// implementation, and dispatch for the "row GEMM with taps" primitive (conv.cuh).
#include <algorithm>
#include <cstring>
#include <cstdlib>
#include "conv.cuh"

namespace ttsb {

static thread_local std::string g_last_error;
void set_last_error(const std::string& msg) { g_last_error = msg; }
const char* get_last_error() { return g_last_error.c_str(); }

static long long g_launches = 0;
void count_launch(int n) { __atomic_fetch_add(&g_launches, static_cast<long long>(n), __ATOMIC_RELAXED); }
long long launch_count() { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

// ---- stage profiler (see common.cuh) ----
namespace {
struct Prof {
    bool on = false;
    std::vector<cudaEvent_t> ev;
    std::vector<int> tag;
    size_t used = 0;
};
Prof g_prof;
}  // namespace
void prof_mark(int tag, cudaStream_t stream) {
    if (!g_prof.on) return;
    if (g_prof.used == g_prof.ev.size()) {
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) return;
        g_prof.ev.push_back(e);
        g_prof.tag.push_back(0);
    }
    g_prof.tag[g_prof.used] = tag;
    cudaEventRecord(g_prof.ev[g_prof.used], stream);
    ++g_prof.used;
}
int prof_enable(int on) {
    g_prof.on = on != 0;
    g_prof.used = 0;
    return 0;
}
int prof_collect(double* ms_by_tag, int n_tags) {
    for (int i = 0; i < n_tags; ++i) ms_by_tag[i] = 0.0;
    TTSB_CHECK_CUDA(cudaDeviceSynchronize());
    for (size_t i = 0; i + 1 < g_prof.used; ++i) {
        float ms = 0.f;
        TTSB_CHECK_CUDA(cudaEventElapsedTime(&ms, g_prof.ev[i], g_prof.ev[i + 1]));
        const int t = g_prof.tag[i];
        if (t >= 0 && t < n_tags) ms_by_tag[t] += ms;
    }
    g_prof.used = 0;
    return 0;
}

int conv_forward_tc(const ConvLayer& L, const ConvRuntime& rt, const __half* in, int ld_in, int B,
                    int T, const EpiParams& epi, cudaStream_t stream);
int conv_forward_tc2(const ConvLayer& L, const ConvRuntime& rt, const __half* in, int ld_in, int B,
                     int T, const EpiParams& epi, cudaStream_t stream);

// Byte offset of element (n, kk) inside one pre-swizzled n_tile x chunk_k weight tile.
// Rows are chunk_k*2 bytes; 16-byte chunks are XOR-swizzled with the row phase exactly like the
// TMA/UMMA SWIZZLE_128B (bits[4:6] ^= bits[7:9]) and SWIZZLE_64B (bits[4:5] ^= bits[7:8]) modes.
__host__ __device__ static inline uint32_t wtile_offset(int n, int kk, int chunk_k) {
    const uint32_t row_bytes = chunk_k * 2;
    const uint32_t lin = n * row_bytes + (kk >> 3) * 16;
    const uint32_t phase = row_bytes == 128 ? ((lin >> 7) & 7u) : ((lin >> 7) & 3u);
    return (lin ^ (phase << 4)) + (kk & 7) * 2;
}

static const size_t kSmemMax = 232448;  // 227 KB opt-in limit per CTA on sm_100

int conv_layer_create(ConvLayer& L, int cin_logical, int cin_stored, int n_total, int n_taps,
                      const int* tap_off0, const int* tap_off1, int class_split,
                      const float* w_logical, const float* bias, int n_tile_hint) {
    TTSB_REQUIRE(n_taps >= 1 && n_taps <= kMaxTaps, "tap count");
    L = ConvLayer();
    L.chunk_k = (cin_stored % 64 == 0) ? 64 : 32;
    TTSB_REQUIRE(cin_stored % L.chunk_k == 0 && cin_stored >= cin_logical, "Cin must be a multiple of 32");
    L.cin = cin_stored;
    L.n_total = n_total;
    L.n_taps = n_taps;
    L.n_chunks = cin_stored / L.chunk_k;
    L.class_split = class_split;
    int lo = 0, hi = 0;
    for (int i = 0; i < n_taps; ++i) {
        L.tap_off[0][i] = tap_off0[i];
        L.tap_off[1][i] = tap_off1 ? tap_off1[i] : tap_off0[i];
        for (int c = 0; c < 2; ++c) {
            lo = std::min(lo, L.tap_off[c][i]);
            hi = std::max(hi, L.tap_off[c][i]);
        }
    }
    L.halo_lo = -lo;
    L.halo_hi = hi;
    L.rows_panel = round_up(kTileM + L.halo_lo + L.halo_hi, 8);
    TTSB_REQUIRE(L.rows_panel <= 256, "receptive field too wide for one TMA box");

    // N tiling: one CTA covers n_tile columns; a single MMA covers at most 256.
    int n_tile = n_tile_hint > 0 ? n_tile_hint : (n_total <= 512 ? n_total : 256);
    TTSB_REQUIRE(n_total % n_tile == 0 && n_tile % 16 == 0 && n_tile <= 512, "N tiling");
    L.n_tile = n_tile;
    L.n_sub = n_tile <= 256 ? 1 : 2;
    TTSB_REQUIRE((n_tile / L.n_sub) % 16 == 0, "N sub-tile must be a multiple of 16");
    L.tmem_cols = 32;
    while (L.tmem_cols < n_tile) L.tmem_cols *= 2;

    // shared memory plan: keep every channel panel resident if it fits next to >= 2 weight stages
    const size_t row_bytes = L.chunk_k * 2;
    const size_t panel = L.rows_panel * row_bytes;
    const size_t btile = n_tile * row_bytes;
    size_t budget = kSmemMax - 2048;
    int max_b = 6;
    if (const char* e = getenv("TTSB_SMEM_BUDGET")) budget = std::min(budget, static_cast<size_t>(atol(e)));
    if (const char* e = getenv("TTSB_MAX_B_STAGES")) max_b = std::max(1, atoi(e));
    int a_slots = L.n_chunks;
    if (a_slots * panel + 2 * btile > budget) a_slots = std::min(L.n_chunks, 2);
    int b_stages = static_cast<int>((budget - a_slots * panel) / btile);
    b_stages = std::min(b_stages, std::min(max_b, L.n_chunks * n_taps));
    TTSB_REQUIRE(b_stages >= 1 && a_slots >= 1, "tile does not fit in shared memory");
    L.a_slots = a_slots;
    L.b_stages = b_stages;
    L.smem_bytes = 1024 + a_slots * panel + b_stages * btile + (2 * a_slots + 2 * b_stages + 1) * 8 + 16;

    // persistent-kernel plan (conv_tc2). Measured on B200 (profiles/r01_s4..s6):
    //  * TWO CTAs per SM beat one whenever a (reduced) configuration fits in half the smem and
    //    half the TMEM: two independent producer/MMA/epilogue pipelines hide each other's bubbles
    //  * weights stay resident in smem when they fit next to the panel ring
    //  * rpp > 1 (several row tiles per weight pass) only by request (TTSB_RPP), see DESIGN.md
    {
        const size_t full = kSmemMax - 2048;
        const size_t half = kSmemMax / 2 - 2048;
        const int total_b = L.n_chunks * n_taps;
        const size_t bar_bytes = 1024 + 1024 + 1024 + 8192 + 2048 + 256;   // alignment slack, barriers, epilogue staging tiles, bias tile
        const size_t w_all = static_cast<size_t>(total_b) * btile;
        const int min_a = std::min(L.n_chunks, 2);
        const size_t third = kSmemMax / 3 - 2048;
        // A layer whose ONE N tile is 256 columns wide (the C = 256 ResBlock convs) gets a single accumulator buffer in half
        // of the TMEM, so a two-CTA plan serialises its MMAs and its epilogue inside each CTA, with two panels and two
        // weight stages: one CTA per SM (double-buffered accumulator, 4 + 4 ring slots, room for the TMA-in epilogue tiles)
        // measured 65 vs 86 us on the k = 3 conv2 (profiles/r02_s22_occupancy.txt). Layers with several N tiles (the
        // up-samplers, conv-FF) keep two CTAs: there the second CTA works on another N tile of the same rows.
        const bool single_wide_tile = n_tile == 256 && n_total == n_tile;
        const bool can2 = min_a * panel + 2 * btile + bar_bytes <= half && n_tile <= 256 && !single_wide_tile;
        const bool can3 = min_a * panel + std::min(3, total_b) * btile + bar_bytes <= third && 2 * n_tile <= 128;
        L.occ2 = can2 ? 2 : 1;
        if (const char* e = getenv("TTSB_OCC2")) L.occ2 = std::max(1, std::min(atoi(e), can3 ? 3 : (can2 ? 2 : 1)));
        const size_t bud = std::min(budget, L.occ2 == 3 ? third : (L.occ2 == 2 ? half : full)) - bar_bytes;
        int want_res = w_all + std::min(L.n_chunks + 1, 2 * L.n_chunks) * panel <= bud ? 1 : 0;
        if (const char* e = getenv("TTSB_RESIDENT")) want_res = want_res && atoi(e) != 0;
        L.resident = want_res;
        const int tmem_cap = L.occ2 == 3 ? 128 : (L.occ2 == 2 ? 256 : 512);
        // row tiles per work item: light (resident) layers batch several 128-row tiles per item so the
        // per-item barrier round trips and MMA issue latencies overlap across independent accumulators
        int rpp = (L.occ2 == 2 && 4 * n_tile <= tmem_cap) ? 2 : 1;   // measured: +10-25 % on C<=64, k>=7
        if (const char* e = getenv("TTSB_RPP")) {
            const int v = atoi(e);
            if ((v == 1 || v == 2 || v == 4) && 2 * v * n_tile <= tmem_cap) rpp = v;
        }
        // a resident plan must keep rpp * min(n_chunks, 2) panels next to the weights: first give up the second
        // row tile per pass, then residency
        if (L.resident && (bud - w_all) / panel < static_cast<size_t>(rpp * std::min(L.n_chunks, 2))) rpp = 1;
        if (L.resident && (bud - w_all) / panel < static_cast<size_t>(std::min(L.n_chunks, 2))) L.resident = 0;
        L.rpp = rpp;
        L.acc_bufs = 2 * rpp * n_tile <= tmem_cap ? 2 : 1;
        L.tmem_cols2 = 32;
        while (L.tmem_cols2 < L.acc_bufs * rpp * n_tile) L.tmem_cols2 *= 2;
        if (L.resident) {
            int as2 = static_cast<int>((bud - w_all) / panel);
            as2 = std::min(as2, 4 * L.n_chunks * rpp);
            TTSB_REQUIRE(as2 >= rpp * std::min(L.n_chunks, 2), "resident plan: panel ring too small for rpp");
            L.a_slots2 = as2;
            L.b_stages2 = 1;
            L.smem_bytes2 = 1024 + as2 * panel + w_all + (2 * as2 + 2 + 5) * 8 + 16 + 1024 + 8192 + 2048 + 256;
        } else {
            const int item = rpp * L.n_chunks;                 // panels of one work item
            const int min_b = std::min(L.occ2 >= 2 ? 3 : 4, total_b);
            int as2 = 2 * item;                                // two work items in flight
            if (as2 * panel + min_b * btile > bud) as2 = item + rpp;      // + the next item's first chunk
            if (as2 * panel + min_b * btile > bud) as2 = item;
            if (as2 * panel + 2 * btile > bud) as2 = std::min(item, 3 * rpp);
            if (as2 * panel + 2 * btile > bud) as2 = std::min(item, 2 * rpp);
            if (as2 * panel + 2 * btile > bud) as2 = rpp;
            int bs2 = static_cast<int>((bud - as2 * panel) / btile);
            bs2 = std::min(bs2, std::min(std::max(max_b, 8), total_b));
            TTSB_REQUIRE(bs2 >= 2 && as2 >= rpp, "persistent tile does not fit in shared memory");
            L.a_slots2 = as2;
            L.b_stages2 = bs2;
            L.smem_bytes2 = 1024 + as2 * panel + bs2 * btile + (2 * as2 + 2 * bs2 + 5) * 8 + 16 + 1024 + 8192 + 2048 + 256;
        }
    }

    // pack weights: [n_tiles][chunk][tap] tiles, fp16, swizzled
    const size_t tile_elems = static_cast<size_t>(n_tile) * L.chunk_k;
    const size_t total = static_cast<size_t>(L.n_tiles()) * L.n_chunks * n_taps * tile_elems;
    std::vector<__half> packed(total, __float2half(0.f));
    for (int nt = 0; nt < L.n_tiles(); ++nt)
        for (int c = 0; c < L.n_chunks; ++c)
            for (int tap = 0; tap < n_taps; ++tap) {
                uint8_t* tile = reinterpret_cast<uint8_t*>(
                    packed.data() + ((static_cast<size_t>(nt) * L.n_chunks + c) * n_taps + tap) * tile_elems);
                for (int n = 0; n < n_tile; ++n) {
                    const float* wrow = w_logical + (static_cast<size_t>(nt * n_tile + n) * n_taps + tap) * cin_logical;
                    for (int kk = 0; kk < L.chunk_k; ++kk) {
                        const int ci = c * L.chunk_k + kk;
                        const float v = ci < cin_logical ? wrow[ci] : 0.f;
                        *reinterpret_cast<__half*>(tile + wtile_offset(n, kk, L.chunk_k)) = __float2half(v);
                    }
                }
            }
    TTSB_CHECK_CUDA(cudaMalloc(&L.w_packed, total * sizeof(__half)));
    TTSB_CHECK_CUDA(cudaMemcpy(L.w_packed, packed.data(), total * sizeof(__half), cudaMemcpyHostToDevice));
    if (cin_stored == 32 && n_total == 32 && n_tile == 32 && cin_logical == 32) {
        const int n_pairs = (n_taps + 1) / 2;
        std::vector<__half> pp(static_cast<size_t>(n_pairs) * 32 * 64, __float2half(0.f));
        for (int g = 0; g < n_pairs; ++g) {
            uint8_t* tile = reinterpret_cast<uint8_t*>(pp.data() + static_cast<size_t>(g) * 32 * 64);
            for (int n = 0; n < 32; ++n)
                for (int kk = 0; kk < 64; ++kk) {
                    const int tap = 2 * g + (kk >> 5);
                    const float v = tap < n_taps ? w_logical[(static_cast<size_t>(n) * n_taps + tap) * cin_logical + (kk & 31)] : 0.f;
                    *reinterpret_cast<__half*>(tile + wtile_offset(n, kk, 64)) = __float2half(v);
                }
        }
        TTSB_CHECK_CUDA(cudaMalloc(&L.w_pair_packed, pp.size() * sizeof(__half)));
        TTSB_CHECK_CUDA(cudaMemcpy(L.w_pair_packed, pp.data(), pp.size() * sizeof(__half), cudaMemcpyHostToDevice));
    }
    if (bias) {
        TTSB_CHECK_CUDA(cudaMalloc(&L.bias, n_total * sizeof(float)));
        TTSB_CHECK_CUDA(cudaMemcpy(L.bias, bias, n_total * sizeof(float), cudaMemcpyHostToDevice));
    }
    return 0;
}

void conv_layer_destroy(ConvLayer& L) {
    if (L.w_packed) cudaFree(L.w_packed);
    if (L.w_pair_packed) cudaFree(L.w_pair_packed);
    if (L.bias) cudaFree(L.bias);
    L.w_pair_packed = nullptr;
    L.w_packed = nullptr;
    L.bias = nullptr;
}

// ------------------------------------------------------------------------------------------------
// SIMT check implementation: plain smem-tiled fp32-FMA GEMM over the SAME packed weights and the
// same fp16 activations, accumulators to a global fp32 scratch, then the SAME epilogue code.
// It exists to (a) cross-check the tcgen05 kernel on the GPU and (b) validate weight packing and
// the epilogue independently of TMA/UMMA descriptors. Selected with TTSB_CONV_IMPL=simt.
// ------------------------------------------------------------------------------------------------
struct SimtArgs {
    int B, T, ld_in;
    int n_chunks, n_taps, chunk_k, n_tile, n_total, class_split;
    int tap_off[2][kMaxTaps];
    const __half* in;
    const __half* w;
    float* scratch;  // [B*T, n_total]
};

// block: 256 threads -> 32 rows x 32 cols (never straddles an N tile: n_tile % 32 == 0);
// thread = 2 rows x 2 cols
__global__ void __launch_bounds__(256) conv_simt_gemm_kernel(const SimtArgs a) {
    __shared__ float sa[32][65];
    __shared__ float sw[32][65];
    const int b = blockIdx.z;
    const int t0 = blockIdx.x * 32;
    const int n0 = blockIdx.y * 32;
    const int ntile = n0 / a.n_tile;
    const int cls = ntile >= a.class_split ? 1 : 0;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    float acc[2][2] = {};
    const size_t tile_elems = static_cast<size_t>(a.n_tile) * a.chunk_k;
    for (int c = 0; c < a.n_chunks; ++c) {
        for (int tap = 0; tap < a.n_taps; ++tap) {
            const int off = a.tap_off[cls][tap];
            for (int i = threadIdx.x; i < 32 * a.chunk_k; i += 256) {
                const int r = i / a.chunk_k, kk = i % a.chunk_k;
                const int t = t0 + r + off;
                float v = 0.f;
                if (t >= 0 && t < a.T)
                    v = __half2float(a.in[(static_cast<size_t>(b) * a.T + t) * a.ld_in + c * a.chunk_k + kk]);
                sa[r][kk] = v;
            }
            const uint8_t* tile = reinterpret_cast<const uint8_t*>(
                a.w + ((static_cast<size_t>(ntile) * a.n_chunks + c) * a.n_taps + tap) * tile_elems);
            for (int i = threadIdx.x; i < 32 * a.chunk_k; i += 256) {
                const int n = i / a.chunk_k, kk = i % a.chunk_k;
                const int nn = n0 + n - ntile * a.n_tile;
                sw[n][kk] = __half2float(*reinterpret_cast<const __half*>(tile + wtile_offset(nn, kk, a.chunk_k)));
            }
            __syncthreads();
            for (int kk = 0; kk < a.chunk_k; ++kk) {
                const float a0 = sa[ty * 2][kk], a1 = sa[ty * 2 + 1][kk];
                const float w0 = sw[tx * 2][kk], w1 = sw[tx * 2 + 1][kk];
                acc[0][0] += a0 * w0; acc[0][1] += a0 * w1;
                acc[1][0] += a1 * w0; acc[1][1] += a1 * w1;
            }
            __syncthreads();
        }
    }
    for (int i = 0; i < 2; ++i) {
        const int t = t0 + ty * 2 + i;
        if (t >= a.T) continue;
        for (int j = 0; j < 2; ++j)
            a.scratch[(static_cast<size_t>(b) * a.T + t) * a.n_total + n0 + tx * 2 + j] = acc[i][j];
    }
}

__global__ void __launch_bounds__(128) conv_simt_epilogue_kernel(const EpiParams e, float* scratch,
                                                                 int n_tile, int n_tiles) {
    const int b = blockIdx.z;
    const int ntile = blockIdx.y;
    const int t = blockIdx.x * 128 + threadIdx.x;
    const bool ok = t < e.T;
    GmemAcc acc{ok ? scratch + (static_cast<size_t>(b) * e.T + t) * e.n_total + ntile * n_tile : nullptr};
    run_epilogue(e, acc, b, t, ok, ntile * n_tile, n_tile, [] {}, [] {});
}

static int conv_forward_simt(const ConvLayer& L, const ConvRuntime& rt, const __half* in, int ld_in,
                             int B, int T, const EpiParams& epi, cudaStream_t stream) {
    const size_t need = static_cast<size_t>(B) * T * L.n_total;
    TTSB_REQUIRE(rt.simt_scratch != nullptr && rt.simt_scratch_elems >= need, "SIMT scratch too small");
    SimtArgs a;
    a.B = B; a.T = T; a.ld_in = ld_in;
    a.n_chunks = L.n_chunks; a.n_taps = L.n_taps; a.chunk_k = L.chunk_k;
    a.n_tile = L.n_tile; a.n_total = L.n_total; a.class_split = L.class_split;
    for (int c = 0; c < 2; ++c)
        for (int i = 0; i < kMaxTaps; ++i) a.tap_off[c][i] = L.tap_off[c][i];
    a.in = in; a.w = L.w_packed; a.scratch = rt.simt_scratch;
    TTSB_REQUIRE(L.n_tile % 32 == 0, "SIMT path needs n_tile % 32 == 0");
    dim3 grid(ceil_div(T, 32), L.n_total / 32, B);
    conv_simt_gemm_kernel<<<grid, 256, 0, stream>>>(a);
    count_launch(2);
    TTSB_CHECK_CUDA(cudaGetLastError());
    dim3 egrid(ceil_div(T, 128), L.n_tiles(), B);
    conv_simt_epilogue_kernel<<<egrid, 128, 0, stream>>>(epi, rt.simt_scratch, L.n_tile, L.n_tiles());
    TTSB_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int conv_forward(const ConvLayer& L, const ConvRuntime& rt, const __half* in, int ld_in, int B,
                 int T, EpiParams epi, cudaStream_t stream) {
    TTSB_REQUIRE(B > 0 && T > 0, "empty batch");
    epi.T = T;
    epi.n_total = L.n_total;
    if (!epi.bias) epi.bias = L.bias;
    if (epi.ln_g) TTSB_REQUIRE(L.n_tiles() == 1, "LayerNorm epilogue needs the whole row in one CTA");
    if (rt.impl == IMPL_SIMT) return conv_forward_simt(L, rt, in, ld_in, B, T, epi, stream);
    if (rt.tc_version == 2) return conv_forward_tc2(L, rt, in, ld_in, B, T, epi, stream);
    return conv_forward_tc(L, rt, in, ld_in, B, T, epi, stream);
}

}  // namespace ttsb
