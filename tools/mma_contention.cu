// Micro-benchmark: does the tcgen05.mma stream (M=128, K=16, 128-byte operand rows) slow down when, on the same SM,
//   (1) a TMA producer keeps streaming 16 KB tiles from L2 into a shared-memory ring (what the weight ring of the
//       C >= 128 conv layers does), and / or
//   (2) four epilogue-like warps keep draining a second TMEM accumulator with tcgen05.ld ?
// The conv kernels sit at ~2.0 x the isolated MMA issue floor (profiles/r01_layer_floors_b64.txt); this separates
// "the tensor pipe is slowed by its neighbours" from "the pipe is idle between tiles".
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I tts_arabic_pytorch_b200/csrc tools/mma_contention.cu -o tools/mma_contention
#include <algorithm>
#include <cstdio>
#include <vector>
#include "common.cuh"

using namespace ttsb;

constexpr int kRingMax = 16;

struct Args {
    int n;              // MMA N
    int n_mma;          // instructions per repetition
    int mode;           // bit 0: TMA fill stream, bit 1: TMEM readers, bit 2: every CTA streams the same addresses
    int fill_bytes;     // bytes per bulk copy
    int ring;           // copies in flight
    const uint8_t* gsrc;
    unsigned gsrc_bytes;
    long long* out;     // [grid][6]: mma cycles (rep 2), fill bytes, fill cycles, tmem loads (warp 2), reader cycles, unused
};

__global__ void __launch_bounds__(192, 1) mma_contention_kernel(const Args a) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_u32 = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw_u32 & 1023u)) & 1023u);
    uint8_t* ring = smem + 48 * 1024;
    __shared__ uint64_t bar;
    __shared__ uint64_t full[kRingMax];
    __shared__ uint32_t tmem_slot;
    __shared__ volatile int done;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;  // 1.0h
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        for (int s = 0; s < kRingMax; ++s) mbar_init(&full[s], 1);
        done = 0;
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc<512>(&tmem_slot);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;

    if (warp == 0) {
        if (elect_one()) {
            const uint32_t idesc = umma_idesc_f16(128, a.n);
            const uint32_t desc_hi = ((8u * 128u) >> 4) | (1u << 14) | (2u << 29);
            const uint32_t lo_flag = 1u << 16;
            const uint32_t a_lo0 = (smem_u32(smem) & 0x3FFFFu) >> 4;
            const uint32_t b_lo0 = a_lo0 + 128 * 8;
            for (int rep = 0; rep < 3; ++rep) {
                const long long t0 = clock64();
                for (int i = 0; i < a.n_mma; i += 4) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t ad = (static_cast<uint64_t>(desc_hi) << 32) | ((a_lo0 + 2 * k) & 0x3FFFu) | lo_flag;
                        const uint64_t bd = (static_cast<uint64_t>(desc_hi) << 32) | ((b_lo0 + 2 * k) & 0x3FFFu) | lo_flag;
                        umma_f16(tmem_base, ad, bd, idesc, (i > 0 || k > 0) ? 1u : 0u);
                    }
                }
                umma_commit(&bar);
                mbar_wait(&bar, rep & 1, nullptr, 0);
                const long long t2 = clock64();
                if (rep == 2) a.out[blockIdx.x * 6 + 0] = t2 - t0;
            }
            done = 1;
        }
    } else if (warp == 1) {
        if ((a.mode & 1) && elect_one()) {
            const long long t0 = clock64();
            unsigned off = (a.mode & 4) ? 0u : (blockIdx.x * 7u * static_cast<unsigned>(a.fill_bytes)) % a.gsrc_bytes;
            long long bytes = 0;
            int it = 0;
            const int kRing = a.ring;
            while (!done) {
                const int s = it % kRing;
                if (it >= kRing) mbar_wait(&full[s], ((it / kRing) - 1) & 1, nullptr, 0);
                if (done) break;
                mbar_expect_tx(&full[s], a.fill_bytes);
                bulk_load_1d(ring + s * a.fill_bytes, a.gsrc + off, a.fill_bytes, &full[s]);
                off += a.fill_bytes;
                if (off + a.fill_bytes > a.gsrc_bytes) off = 0;
                bytes += a.fill_bytes;
                ++it;
            }
            const long long t1 = clock64();
            // every copy in flight must land before the CTA may exit
            const int issued = it;
            for (int j = (issued > kRing ? issued - kRing : 0); j < issued; ++j)
                mbar_wait(&full[j % kRing], (j / kRing) & 1, nullptr, 0);
            a.out[blockIdx.x * 6 + 1] = bytes;
            a.out[blockIdx.x * 6 + 2] = t1 - t0;
        }
    } else {
        if (a.mode & 2) {
            const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
            float sink = 0.f;
            long long loads = 0;
            const long long t0 = clock64();
            while (!done) {
#pragma unroll 1
                for (int c = 0; c < 4; ++c) {
                    float v[32];
                    tmem_ld32(tmem_base + lane_base + 256 + c * 32, v);
#pragma unroll
                    for (int q = 0; q < 32; ++q) sink += v[q];
                    ++loads;
                }
            }
            const long long t1 = clock64();
            if (sink == 123.456f) a.out[blockIdx.x * 6 + 5] = 1;     // keep the loads alive
            if (warp == 2 && (threadIdx.x & 31) == 0) {
                a.out[blockIdx.x * 6 + 3] = loads;
                a.out[blockIdx.x * 6 + 4] = t1 - t0;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem_base);
}

int main() {
    int dev = 0, sms = 0;
    cudaSetDevice(dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long* d_out;
    cudaMalloc(&d_out, sms * 6 * sizeof(long long));
    const unsigned gsrc_bytes = 4u << 20;
    uint8_t* d_src;
    cudaMalloc(&d_src, gsrc_bytes);
    cudaMemset(d_src, 0, gsrc_bytes);
    const int smem_bytes = 48 * 1024 + 160 * 1024 + 1024;
    cudaFuncSetAttribute(mma_contention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    std::vector<long long> h(sms * 6);
    printf("%5s %5s %5s %4s %6s %6s | %10s %12s %14s   (median over the CTAs of the grid)\n", "N", "grid", "fill", "ring", "tmemrd", "same",
           "cyc/MMA", "fill B/clk", "tmem ld32/kclk");
    const int n_mma = 4096;
    struct Case { int n, grid, mode, fill_bytes, ring; };
    std::vector<Case> cases;
    for (int grid : {sms, 1}) {
        for (int ring : {1, 2, 4, 8, 10}) cases.push_back({128, grid, 1, 16384, ring});
        for (int ring : {4, 8, 16}) cases.push_back({128, grid, 1, 8192, ring});
        for (int ring : {4, 16}) cases.push_back({128, grid, 1, 4096, ring});
        cases.push_back({128, grid, 1, 32768, 4});
    }
    cases.push_back({256, sms, 1, 16384, 10});
    cases.push_back({64, sms, 1, 16384, 10});
    cases.push_back({128, sms, 5, 16384, 10});
    cases.push_back({128, sms, 3, 16384, 10});
    for (const Case& c : cases) {
        cudaMemset(d_out, 0, sms * 6 * sizeof(long long));
        Args a{c.n, n_mma, c.mode, c.fill_bytes, c.ring, d_src, gsrc_bytes, d_out};
        mma_contention_kernel<<<c.grid, 192, smem_bytes>>>(a);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h.data(), d_out, sms * 6 * sizeof(long long), cudaMemcpyDeviceToHost);
        std::vector<double> mma, fill, rd;
        for (int i = 0; i < c.grid; ++i) {
            mma.push_back(static_cast<double>(h[6 * i]) / n_mma);
            fill.push_back(h[6 * i + 2] > 0 ? static_cast<double>(h[6 * i + 1]) / h[6 * i + 2] : 0.0);
            rd.push_back(h[6 * i + 4] > 0 ? 1000.0 * h[6 * i + 3] / h[6 * i + 4] : 0.0);
        }
        std::sort(mma.begin(), mma.end());
        std::sort(fill.begin(), fill.end());
        std::sort(rd.begin(), rd.end());
        const int m = c.grid / 2;
        printf("%5d %5d %5d %4d %6d %6d | %10.1f %12.1f %14.2f\n", c.n, c.grid, (c.mode & 1) ? c.fill_bytes : 0, c.ring, (c.mode >> 1) & 1,
               (c.mode >> 2) & 1, mma[m], fill[m], rd[m]);
    }
    return 0;
}
