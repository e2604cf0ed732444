// Shared device/host helpers for the ttsb200 kernels (sm_100a only).
//
// Everything here is plumbing: error propagation to the C-ABI, fp16 pack/unpack,
// and thin inline-PTX wrappers for mbarrier / TMA / tcgen05. No reference code
// corresponds to this file (the reference has no native code: SURVEY.md §2a).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda.h>
#include <stdint.h>
#include <stdio.h>
#include <exception>
#include <memory>
#include <string>

namespace ttsb {

// ----------------------------------------------------------------------------------------------
// host-side error handling: every C-ABI entry returns an int status, message kept thread-local
// ----------------------------------------------------------------------------------------------
void set_last_error(const std::string& msg);
const char* get_last_error();

#define TTSB_CHECK_CUDA(expr)                                                              \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess) {                                                           \
            ::ttsb::set_last_error(std::string(#expr) + " failed: " + cudaGetErrorString(_e) + \
                                   " (" + __FILE__ + ":" + std::to_string(__LINE__) + ")"); \
            return 2;                                                                      \
        }                                                                                  \
    } while (0)

#define TTSB_REQUIRE(cond, msg)                                                            \
    do {                                                                                   \
        if (!(cond)) {                                                                     \
            ::ttsb::set_last_error(std::string("requirement failed: ") + #cond + ": " + (msg) + \
                                   " (" + __FILE__ + ":" + std::to_string(__LINE__) + ")"); \
            return 1;                                                                      \
        }                                                                                  \
    } while (0)

#define TTSB_PROPAGATE(expr)                                                               \
    do {                                                                                   \
        int _s = (expr);                                                                   \
        if (_s != 0) return _s;                                                            \
    } while (0)

// No C++ exception (std::bad_alloc, std::string growth in the macros above) may cross the extern "C" boundary into ctypes:
// every int-returning entry point runs its body through this.
template <class F>
static inline int guarded_call(F&& f) {
    try {
        return f();
    } catch (const std::exception& e) {
        ::ttsb::set_last_error(std::string("C++ exception: ") + e.what());
        return 3;
    } catch (...) {
        ::ttsb::set_last_error("unknown C++ exception");
        return 3;
    }
}

// launch accounting for bench.py's `gpu_launches` (ttsb_launch_count in the C ABI)
void count_launch(int n = 1);
long long launch_count();

// Per-device process state. Every ttsb_*_create() takes a device ordinal, so nothing that belongs to a device (function
// attributes, the SM count, the error flag, scratch buffers) may be cached once per process.
constexpr int kMaxDevices = 64;
static inline int current_device() {
    int dev = 0;
    cudaGetDevice(&dev);
    return dev < 0 ? 0 : (dev >= kMaxDevices ? kMaxDevices - 1 : dev);
}
struct PerDeviceOnce {
    bool done[kMaxDevices] = {};
    bool& here() { return done[current_device()]; }
};
// Makes `device` current for the lifetime of the guard (forward entry points: the caller's current device may differ
// from the handle's) and restores the previous one.
struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    cudaError_t status = cudaSuccess;
    explicit DeviceGuard(int device) {
        status = cudaGetDevice(&prev);
        if (status == cudaSuccess && prev != device) {
            status = cudaSetDevice(device);
            switched = status == cudaSuccess;
        }
    }
    ~DeviceGuard() { if (switched) cudaSetDevice(prev); }
};
#define TTSB_DEVICE_GUARD(device)                    \
    ::ttsb::DeviceGuard _ttsb_guard(device);         \
    TTSB_CHECK_CUDA(_ttsb_guard.status)

// Stage profiler for bench.py's per-kernel rooflines (ttsb_prof_* in the C ABI): when enabled, every stage of the model
// entry points records a CUDA event ON THE LAUNCHING STREAM before its first launch; the device time between two
// consecutive marks is charged to the earlier mark's tag. Off (the default) it is one predictable branch per stage.
enum ProfTag : int {
    PROF_NONE = 0,        // gap between entry points / untagged work
    PROF_VOC_PRE, PROF_VOC_UPS0, PROF_VOC_S0, PROF_VOC_UPS1, PROF_VOC_S1, PROF_VOC_UPS2, PROF_VOC_S2, PROF_VOC_UPS3,
    PROF_VOC_S3, PROF_VOC_POST, PROF_VOC_PACK,
    PROF_FP_EMBED, PROF_FP_QKV_O, PROF_FP_ATTENTION, PROF_FP_FFN, PROF_FP_PREDICTORS, PROF_FP_GLUE, PROF_FP_REGULATE,
    PROF_FP_PROJ,
    PROF_N_TAGS
};
void prof_mark(int tag, cudaStream_t stream);

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline int round_up(int a, int b) { return ceil_div(a, b) * b; }

#ifdef __CUDACC__
// ----------------------------------------------------------------------------------------------
// small device helpers
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack_half2(uint32_t u) {
    __half2 h = *reinterpret_cast<__half2*>(&u);
    return __half22float2(h);
}
__device__ __forceinline__ float lrelu(float v, float slope) { return v > 0.f ? v : v * slope; }

__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(pred));
    return pred != 0;
}

// ----------------------------------------------------------------------------------------------
// mbarrier
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

// non-blocking phase test: used to look at the NEXT ring stage's barrier before the current stage's MMAs are issued, so
// the ~100-cycle barrier read overlaps the issue instead of heading every stage's dependent chain
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

// try_wait with a suspend-time hint (ns): the warp sleeps in hardware until the phase completes or the hint
// expires, instead of coming back every few hundred cycles to re-issue the poll.
__device__ __forceinline__ bool mbar_try_wait_hint(uint64_t* bar, uint32_t parity, uint32_t hint_ns) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(hint_ns)
        : "memory");
    return ok != 0;
}

// Bounded wait: a protocol bug must surface as an error code in `*err_flag`, never as a hung GPU box
// (a few seconds of SM clocks; normal waits take microseconds).
// The poll loop matters: with the default (short) try_wait time limit every waiting warp re-issues its
// poll + loop bookkeeping every few hundred cycles, and with ~10 warps of a CTA waiting at any time those
// instructions took ~1/3 of the SM's issue slots away from the MMA-issuing thread and the epilogue warps
// (profiles/r01_s24_ncu_full_b64.csv: issue-active 48 %, ALU pipe 39 % in a kernel that is mostly waiting).
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int* err_flag, int code) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
#ifdef TTSB_MBAR_SPIN   // experiment: default try_wait time limit (tools/build_flags.py)
    while (!mbar_try_wait(bar, parity)) {
#else
    while (!mbar_try_wait_hint(bar, parity, 100000u)) {
#endif
        if (clock64() - t0 > 6000000000ll) {
            if (err_flag) atomicExch(err_flag, code);
            __trap();
        }
    }
}

// ----------------------------------------------------------------------------------------------
// TMA (cp.async.bulk[.tensor]) — completion is signalled on an mbarrier as transaction bytes
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar,
                                            int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)),
          "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes,
                                             uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes),
          "r"(smem_u32(bar))
        : "memory");
}

// TMA store of one box from shared memory (bulk async-group completion: commit, then wait for the READ of the source
// tile before it is overwritten; the global writes complete in the background)
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
        ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// 1-D bulk copy delivered to the same CTA-relative offset (data and mbarrier) in every CTA of `cta_mask`
__device__ __forceinline__ void bulk_load_1d_multicast(void* smem_dst, const void* gsrc, uint32_t bytes,
                                                       uint64_t* bar, uint16_t cta_mask) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar)), "h"(cta_mask)
        : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ----------------------------------------------------------------------------------------------
// tcgen05 / TMEM
// ----------------------------------------------------------------------------------------------
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32(smem_result)), "n"(kCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols)
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]; one thread issues for the whole CTA.
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                         uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// mbarrier arrives once every MMA issued so far by this thread has completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                 ::"r"(smem_u32(bar))
                 : "memory");
}
// same, arriving on the barrier at this CTA-relative offset in every CTA of `cta_mask`
__device__ __forceinline__ void umma_commit_multicast(uint64_t* bar, uint16_t cta_mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(cta_mask)
                 : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread i of the warp gets TMEM lane (base_lane + i).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
          "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
          "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]),
          "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// Split form for software pipelining: issue the load of chunk c+1, then consume chunk c. The wait names the
// destination registers as in/out operands so that no use of them can be scheduled above it.
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, float (&v)[32]) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
          "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
          "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]),
          "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait(float (&v)[32]) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.wait::ld.sync.aligned;"
        : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]),
          "+r"(r[7]), "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]),
          "+r"(r[14]), "+r"(r[15]), "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]),
          "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]), "+r"(r[25]),
          "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
        :
        : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&v)[32]) {
    const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%32], "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31};"
        ::"r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
          "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]),
          "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]),
          "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]),
          "r"(r[29]), "r"(r[30]), "r"(r[31]), "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// Shared-memory matrix descriptor, K-major operand, 128-byte swizzle (cute::UMMA::SmemDescriptor
// bit layout: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48),
// base_offset [49,52), layout_type [61,64) with SWIZZLE_128B = 2).
// Rows are 128 B apart, 8-row swizzle atoms 1024 B apart (SBO = 1024).
__device__ __forceinline__ uint64_t umma_desc_sw128_kmajor(uint32_t saddr, uint32_t base_offset) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((saddr & 0x3FFFF) >> 4);
    d |= static_cast<uint64_t>(1) << 16;            // LBO (ignored for swizzled K-major; cute sets 1)
    d |= static_cast<uint64_t>(1024 >> 4) << 32;    // SBO
    d |= static_cast<uint64_t>(1) << 46;            // descriptor version (Blackwell)
    d |= static_cast<uint64_t>(base_offset & 7) << 49;
    d |= static_cast<uint64_t>(2) << 61;            // SWIZZLE_128B
    return d;
}
// Instruction descriptor for kind::f16: fp16 A/B (K-major both), fp32 accumulate, M x N tile.
__host__ __device__ constexpr uint32_t umma_idesc_f16(int M, int N) {
    return (1u << 4)                         // c_format = F32
           | (0u << 7) | (0u << 10)          // a/b format = F16
           | (0u << 15) | (0u << 16)         // a/b major = K
           | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}
#endif  // __CUDACC__

}  // namespace ttsb
