// Miscellaneous C-ABI entry points: error string, runtime switches, and the single-conv-site
// handle used by the parity tests (include/ttsb200.h).
#include "model_common.cuh"

using namespace ttsb;

struct ttsb_conv1d {
    ConvLayer layer;
    int kind = 0, cin = 0, cout = 0, stride = 1;
};

extern "C" {

const char* ttsb_last_error(void) { return get_last_error(); }
int ttsb_version(void) { return 1; }

int ttsb_set_conv_impl(int impl) {
    TTSB_REQUIRE(impl == IMPL_TC || impl == IMPL_SIMT, "impl must be 0 (tcgen05) or 1 (simt)");
    global_runtime().impl = impl;
    return 0;
}
int ttsb_set_desc_mode(int mode) {
    TTSB_REQUIRE(mode >= 0 && mode <= 3, "desc mode must be 0..3");
    global_runtime().desc_mode = mode;
    return 0;
}
int ttsb_set_tc_version(int v) {
    TTSB_REQUIRE(v == 1 || v == 2, "tc version must be 1 (one tile per CTA) or 2 (persistent)");
    global_runtime().tc_version = v;
    return 0;
}
int ttsb_get_conv_impl(void) { return global_runtime().impl; }
int ttsb_get_desc_mode(void) { return global_runtime().desc_mode; }
int64_t ttsb_launch_count(void) { return launch_count(); }

int ttsb_device_error_flag(int* h_flag) {
    TTSB_REQUIRE(h_flag != nullptr, "null argument");
    *h_flag = 0;
    TTSB_CHECK_CUDA(cudaDeviceSynchronize());
    GlobalRuntime& g = global_runtime();
    if (g.err_flag) TTSB_CHECK_CUDA(cudaMemcpy(h_flag, g.err_flag, sizeof(int), cudaMemcpyDeviceToHost));
    return 0;
}

int ttsb_debug_set_timeline(void* d_buf) {
    global_runtime().timeline = static_cast<long long*>(d_buf);
    return 0;
}

int ttsb_conv1d_create(int kind, int cin, int cout, int ksize, int dilation, int stride,
                       const float* h_weight, const float* h_bias, int device, ttsb_conv1d_t** out) {
    TTSB_REQUIRE(h_weight && out, "null argument");
    TTSB_CHECK_CUDA(cudaSetDevice(device));
    ttsb_conv1d* h = new ttsb_conv1d();
    h->kind = kind; h->cin = cin; h->cout = cout;
    int st;
    if (kind == 0) {
        h->stride = 1;
        const int cin_pad = cin % 64 == 0 ? cin : (cin <= 32 ? 32 : round_up(cin, 64));
        st = make_conv1d_layer(h->layer, h_weight, h_bias, cout, cin, ksize, dilation, cin_pad, 0);
    } else {
        h->stride = stride;
        st = make_convT1d_layer(h->layer, h_weight, h_bias, cin, cout, ksize, stride);
    }
    if (st != 0) { delete h; return st; }
    *out = h;
    return 0;
}
void ttsb_conv1d_destroy(ttsb_conv1d_t* h) {
    if (!h) return;
    conv_layer_destroy(h->layer);
    delete h;
}
int ttsb_conv1d_cin_pad(const ttsb_conv1d_t* h) { return h ? h->layer.cin : 0; }

int ttsb_conv1d_forward(ttsb_conv1d_t* h, const void* d_in, int B, int T, const void* d_residual,
                        float act_slope, const int32_t* d_lens, void* d_out, void* stream) {
    TTSB_REQUIRE(h && d_in && d_out, "null argument");
    ConvRuntime rt;
    TTSB_PROPAGATE(get_conv_runtime(static_cast<size_t>(B) * T * h->layer.n_total, rt));
    EpiParams e;
    e.lens = d_lens; e.len_mul = 1;
    e.residual = static_cast<const __half*>(d_residual); e.ld_res = h->layer.n_total;
    if (act_slope >= 0.f) {
        e.out_act = static_cast<__half*>(d_out); e.ld_act = h->layer.n_total; e.act_slope = act_slope;
    } else {
        e.out_raw = static_cast<__half*>(d_out); e.ld_raw = h->layer.n_total;
    }
    return conv_forward(h->layer, rt, static_cast<const __half*>(d_in), h->layer.cin, B, T, e,
                        static_cast<cudaStream_t>(stream));
}

}  // extern "C"
