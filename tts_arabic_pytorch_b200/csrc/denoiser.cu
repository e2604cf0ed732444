// HiFi-GAN bias denoiser as two CUDA kernels for a whole padded batch
// (vocoder/hifigan/denoiser.py:66-72: STFT -> max(|X| - strength*bias, 0) * e^{j arg X} -> ISTFT; the transforms are
// torchaudio Spectrogram / InverseSpectrogram with n_fft = win = 1024, hop 256, periodic hann, center=True,
// pad_mode='reflect', onesided, not normalised — denoiser.py:43-48).
//
// The reference runs it once per utterance (models/fastpitch/networks.py:343-344), each call a train of cuFFT and
// elementwise launches plus a host sync; here every (utterance, frame) is one CTA that keeps the frame in shared
// memory from the reflect-padded window load to the windowed inverse transform, and a second kernel does the
// overlap-add with the window-envelope normalisation of torch.istft. Per-utterance semantics are kept: the reflect
// padding and the frame count use each utterance's own length, samples beyond it come out as zeros.
//
// HBM-bound: algorithmic bytes per audio sample = 4 (read) + 4 x 4 (each sample lives in 4 overlapping 1024-sample
// frames: fp32 frame write) + 4 x 4 (frame read) + 4 (write) = 40 B.
#include "model_common.cuh"

namespace ttsb {

constexpr int kFft = 1024;
constexpr int kHop = 256;
constexpr int kBins = kFft / 2 + 1;
constexpr int kFftThreads = 256;

__device__ __forceinline__ int bitrev10(int i) { return static_cast<int>(__brev(static_cast<unsigned>(i)) >> 22); }

// in-place radix-2 decimation-in-time FFT of 1024 complex points already stored in bit-reversed order;
// tw[q] = exp(-2 pi i q / 1024), q < 512; kInverse uses the conjugate twiddles (no 1/N scaling)
template <bool kInverse>
__device__ __forceinline__ void fft1024(float2* buf, const float2* tw) {
#pragma unroll 1
    for (int s = 1; s <= 10; ++s) {
        const int half = 1 << (s - 1);
        const int tw_step = kFft >> s;
        for (int j = threadIdx.x; j < kFft / 2; j += kFftThreads) {
            const int pos = j & (half - 1);
            const int i0 = ((j - pos) << 1) + pos;
            const int i1 = i0 + half;
            float2 w = tw[pos * tw_step];
            if (kInverse) w.y = -w.y;
            const float2 a = buf[i0], b = buf[i1];
            const float2 t = make_float2(w.x * b.x - w.y * b.y, w.x * b.y + w.y * b.x);
            buf[i0] = make_float2(a.x + t.x, a.y + t.y);
            buf[i1] = make_float2(a.x - t.x, a.y - t.y);
        }
        __syncthreads();
    }
}

// grid (F_max, B): frame f of utterance b. frames_out: [B, F_max, 1024] fp32 (windowed inverse transforms).
__global__ void __launch_bounds__(kFftThreads) denoise_frames_kernel(const float* wav, const int* n_samples, int n_max,
                                                                     const float* bias_spec, float strength,
                                                                     float* frames_out, int f_max) {
    __shared__ float2 buf[kFft];
    __shared__ float2 tmp[kFft];
    __shared__ float2 tw[kFft / 2];
    const int b = blockIdx.y, f = blockIdx.x;
    const int n = min(n_samples[b], n_max);
    if (n <= 0 || f > n / kHop) return;   // frames of this utterance: 1 + n / hop  (center=True)
    const float* x = wav + static_cast<size_t>(b) * n_max;
    for (int q = threadIdx.x; q < kFft / 2; q += kFftThreads) {
        float sn, cs;
        sincospif(static_cast<float>(q) * (2.0f / kFft), &sn, &cs);
        tw[q] = make_float2(cs, -sn);
    }
    // window (periodic hann = sin^2(pi i / N)) x reflect-padded signal, stored bit-reversed for the DIT transform
    for (int i = threadIdx.x; i < kFft; i += kFftThreads) {
        int j = f * kHop + i - kFft / 2;
        if (j < 0) j = -j;
        if (j >= n) j = 2 * (n - 1) - j;
        j = max(0, min(j, n - 1));        // only reachable for n <= 512, where the reference refuses to pad
        const float s = sinpif(static_cast<float>(i) * (1.0f / kFft));
        buf[bitrev10(i)] = make_float2(x[j] * s * s, 0.f);
    }
    __syncthreads();
    fft1024<false>(buf, tw);
    // spectral subtraction on the one-sided spectrum, hermitian completion, bit-reversed for the inverse
    for (int k = threadIdx.x; k < kBins; k += kFftThreads) {
        float2 v = buf[k];
        const float mag = sqrtf(v.x * v.x + v.y * v.y);
        const float m2 = fmaxf(mag - strength * bias_spec[k], 0.f);
        const float sc = mag > 0.f ? m2 / mag : 0.f;
        v.x *= sc; v.y *= sc;
        if (k == 0 || k == kFft / 2) v.y = 0.f;   // c2r ignores the imaginary part of DC and Nyquist
        tmp[bitrev10(k)] = v;
        if (k > 0 && k < kFft / 2) tmp[bitrev10(kFft - k)] = make_float2(v.x, -v.y);
    }
    __syncthreads();
    fft1024<true>(tmp, tw);
    float* out = frames_out + (static_cast<size_t>(b) * f_max + f) * kFft;
    for (int i = threadIdx.x; i < kFft; i += kFftThreads) {
        const float s = sinpif(static_cast<float>(i) * (1.0f / kFft));
        out[i] = tmp[i].x * (1.0f / kFft) * s * s;
    }
}

// out[b, t] = sum_f frame[f][t + 512 - 256 f] / sum_f w^2[t + 512 - 256 f] over the utterance's own frames
__global__ void denoise_overlap_add_kernel(const float* frames, const int* n_samples, int n_max, int f_max, float* out) {
    const int b = blockIdx.y;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_max) return;
    const int n = min(n_samples[b], n_max);
    float y = 0.f;
    if (t < n) {
        const int n_frames = 1 + n / kHop;
        const int p = t + kFft / 2;
        const int f_hi = min(n_frames - 1, p / kHop);
        const int f_lo = max(0, (p - kFft + kHop) / kHop);
        float acc = 0.f, env = 0.f;
        for (int f = f_lo; f <= f_hi; ++f) {
            const int i = p - f * kHop;
            const float s = sinpif(static_cast<float>(i) * (1.0f / kFft));
            const float w = s * s;
            acc += frames[(static_cast<size_t>(b) * f_max + f) * kFft + i];
            env += w * w;
        }
        y = env > 1e-11f ? acc / env : acc;
    }
    out[static_cast<size_t>(b) * n_max + t] = y;
}

}  // namespace ttsb

using namespace ttsb;

extern "C" {

size_t ttsb_denoiser_workspace_bytes(int B, int n_max) {
    if (B <= 0 || n_max <= 0) return 0;
    const size_t f_max = 1 + static_cast<size_t>(n_max) / kHop;
    return static_cast<size_t>(B) * f_max * kFft * sizeof(float) + 256;
}

int ttsb_denoiser_forward(const float* d_wav, const int32_t* d_n_samples, int B, int n_max, const float* d_bias_spec,
                          float strength, float* d_out, void* d_workspace, size_t workspace_bytes, void* stream_) {
    return guarded_call([&]() -> int {
    TTSB_REQUIRE(d_wav && d_n_samples && d_bias_spec && d_out && d_workspace, "null argument");
    TTSB_REQUIRE(B > 0 && n_max > 0, "empty batch");
    TTSB_REQUIRE(workspace_bytes >= ttsb_denoiser_workspace_bytes(B, n_max), "workspace too small");
    TTSB_REQUIRE(B <= 65535, "batch too large for one launch");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const int f_max = 1 + n_max / kHop;
    float* frames = static_cast<float*>(d_workspace);
    denoise_frames_kernel<<<dim3(f_max, B), kFftThreads, 0, stream>>>(d_wav, d_n_samples, n_max, d_bias_spec, strength,
                                                                      frames, f_max);
    count_launch();
    TTSB_CHECK_CUDA(cudaGetLastError());
    denoise_overlap_add_kernel<<<dim3(ceil_div(n_max, 256), B), 256, 0, stream>>>(frames, d_n_samples, n_max, f_max, d_out);
    count_launch();
    TTSB_CHECK_CUDA(cudaGetLastError());
    return 0;
    });
}

}  // extern "C"
