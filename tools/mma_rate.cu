// Micro-benchmark: issue/execute rate of tcgen05.mma.cta_group::1.kind::f16 (M=128, K=16 per instruction)
// as a function of N and of the number of independent TMEM accumulators the instruction stream
// round-robins over. One CTA per SM, one issuing thread, operands = whatever is in shared memory.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I tts_arabic_pytorch_b200/csrc tools/mma_rate.cu -o tools/mma_rate
#include <algorithm>
#include <cstdio>
#include <vector>
#include "common.cuh"

using namespace ttsb;

struct Args {
    int n;          // MMA N
    int nacc;       // independent accumulators (round robin)
    int n_mma;      // instructions per measurement
    int row_bytes;  // 128 (SW128) or 64 (SW64)
    int a_stride_rows;   // A view advances by this many rows per MMA (0 = same view), wraps in 64 rows
    int commit_every;    // tcgen05.commit after this many MMAs (0 = only at the end)
    long long* out;      // [grid][2]: issue cycles, total cycles
};

__global__ void __launch_bounds__(128, 1) mma_rate_kernel(const Args a) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_u32 = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw_u32 & 1023u)) & 1023u);
    __shared__ uint64_t bar;
    __shared__ uint64_t bar2;
    __shared__ uint32_t tmem_slot;
    // A region: 256 rows, B region: 256 rows
    for (int i = threadIdx.x; i < 512 * a.row_bytes / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;  // 1.0h
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&bar2, 1); fence_mbar_init(); }
    if (threadIdx.x < 32) tmem_alloc<512>(&tmem_slot);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    if (threadIdx.x < 32 && elect_one()) {
        const uint32_t idesc = umma_idesc_f16(128, a.n);
        const uint32_t row_u = a.row_bytes >> 4;
        const uint32_t desc_hi = ((8u * a.row_bytes) >> 4) | (1u << 14) | ((a.row_bytes == 128 ? 2u : 4u) << 29);
        const uint32_t lo_flag = 1u << 16;
        const uint32_t a_lo0 = (smem_u32(smem) & 0x3FFFFu) >> 4;
        const uint32_t b_lo0 = a_lo0 + 256 * row_u;
        const int ksteps = a.row_bytes / 32;
        for (int rep = 0; rep < 3; ++rep) {
            const long long t0 = clock64();
            int acc = 0, shift = 0, since_commit = 0;
            // one "tap" = ksteps MMAs on one (A view, B tile) pair into one accumulator, like the conv kernels
            for (int i = 0; i < a.n_mma; i += ksteps) {
                const uint32_t a_lo = a_lo0 + shift * row_u;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (k < ksteps) {
                        const uint64_t ad = (static_cast<uint64_t>(desc_hi) << 32) | ((a_lo + 2 * k) & 0x3FFFu) | lo_flag;
                        const uint64_t bd = (static_cast<uint64_t>(desc_hi) << 32) | ((b_lo0 + 2 * k) & 0x3FFFu) | lo_flag;
                        umma_f16(tmem_base + acc * a.n, ad, bd, idesc, (i >= a.nacc * ksteps || k > 0) ? 1u : 0u);
                    }
                }
                if (++acc == a.nacc) acc = 0;
                shift += a.a_stride_rows;
                if (shift >= 64) shift -= 64;
                since_commit += ksteps;
                if (a.commit_every > 0 && since_commit >= a.commit_every) { umma_commit(&bar2); since_commit = 0; }
            }
            const long long t1 = clock64();
            umma_commit(&bar);
            mbar_wait(&bar, rep & 1, nullptr, 0);
            const long long t2 = clock64();
            if (rep == 2) {
                a.out[blockIdx.x * 2 + 0] = t1 - t0;
                a.out[blockIdx.x * 2 + 1] = t2 - t0;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc<512>(tmem_base);
}

int main() {
    int dev = 0, sms = 0;
    cudaSetDevice(dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long* d_out;
    cudaMalloc(&d_out, sms * 2 * sizeof(long long));
    cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
    std::vector<long long> h(sms * 2);
    printf("%5s %5s %5s %6s %6s | %10s %10s   (cycles per MMA, median over %d SMs; grid = all SMs)\n", "N", "nacc", "rowB", "astride",
           "commit", "issue", "total", sms);
    const int n_mma = 1024;
    struct Case { int n, nacc, row_bytes, astride, commit; };
    std::vector<Case> cases;
    for (int rb : {128, 64})
        for (int n : {32, 64, 128, 256})
            for (int nacc : {1, 2, 4}) {
                if (n * nacc > 512) continue;
                cases.push_back({n, nacc, rb, 0, 0});
            }
    for (int n : {32, 64, 128}) {
        cases.push_back({n, 1, 128, 1, 0});
        cases.push_back({n, 1, 128, 5, 0});
        cases.push_back({n, 2, 128, 5, 0});
        cases.push_back({n, 1, 128, 0, 4});
        cases.push_back({n, 2, 128, 0, 44});
    }
    for (const Case& c : cases) {
        Args a{c.n, c.nacc, n_mma, c.row_bytes, c.astride, c.commit, d_out};
        mma_rate_kernel<<<sms, 128, 72 * 1024>>>(a);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h.data(), d_out, sms * 2 * sizeof(long long), cudaMemcpyDeviceToHost);
        std::vector<long long> iss, tot;
        for (int i = 0; i < sms; ++i) { iss.push_back(h[2 * i]); tot.push_back(h[2 * i + 1]); }
        std::sort(iss.begin(), iss.end());
        std::sort(tot.begin(), tot.end());
        printf("%5d %5d %5d %6d %6d | %10.1f %10.1f\n", c.n, c.nacc, c.row_bytes, c.astride, c.commit,
               static_cast<double>(iss[sms / 2]) / n_mma, static_cast<double>(tot[sms / 2]) / n_mma);
    }
    return 0;
}
