// Micro-benchmark: how long does one failed mbarrier.try_wait block, with and without a suspend-time hint?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I tts_arabic_pytorch_b200/csrc tools/poll_rate.cu -o tools/poll_rate
#include <cstdio>
#include "common.cuh"
using namespace ttsb;

__global__ void poll_kernel(int mode, uint32_t hint, int n, long long* out) {
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    __syncthreads();
    const long long t0 = clock64();
    int fails = 0;
    for (int i = 0; i < n; ++i) {
        bool ok = mode == 0 ? mbar_try_wait(&bar, 0) : mbar_try_wait_hint(&bar, 0, hint);
        fails += ok ? 0 : 1;
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = fails; }
}

int main() {
    long long* d; cudaMalloc(&d, 16);
    long long h[2];
    const int n = 200;
    struct { int mode; uint32_t hint; } cases[] = {{0, 0}, {1, 1000}, {1, 10000}, {1, 100000}, {1, 1000000}, {1, 0x989680}};
    for (auto c : cases) {
        poll_kernel<<<1, 32>>>(c.mode, c.hint, n, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("mode %d hint %8u ns: %10.1f cycles per failed try_wait (%lld fails of %d)\n", c.mode, c.hint, double(h[0]) / n, h[1], n);
    }
    return 0;
}
