// Miscellaneous C-ABI entry points: error string, runtime switches, and the single-conv-site
// handle used by the parity tests (include/ttsb200.h).
#include "model_common.cuh"

using namespace ttsb;

struct ttsb_conv1d {
    ConvLayer layer;
    int kind = 0, cin = 0, cout = 0, stride = 1, device = 0;
};

struct ttsb_convpair {
    ConvLayer c1, c2;
    ConvPairPlan plan;
    int device = 0;
};

extern "C" {

const char* ttsb_last_error(void) { return get_last_error(); }
int ttsb_version(void) { return 1; }

int ttsb_set_conv_impl(int impl) {
    return guarded_call([&]() -> int {
    TTSB_REQUIRE(impl == IMPL_TC || impl == IMPL_SIMT, "impl must be 0 (tcgen05) or 1 (simt)");
    global_runtime().impl = impl;
    return 0;
    });
}
int ttsb_set_desc_mode(int mode) {
    return guarded_call([&]() -> int {
    TTSB_REQUIRE(mode >= 0 && mode <= 3, "desc mode must be 0..3");
    global_runtime().desc_mode = mode;
    return 0;
    });
}
int ttsb_set_tc_version(int v) {
    return guarded_call([&]() -> int {
    TTSB_REQUIRE(v == 1 || v == 2, "tc version must be 1 (one tile per CTA) or 2 (persistent)");
    global_runtime().tc_version = v;
    return 0;
    });
}
int ttsb_get_conv_impl(void) { return global_runtime().impl; }
int ttsb_get_desc_mode(void) { return global_runtime().desc_mode; }
int64_t ttsb_launch_count(void) { return launch_count(); }

int ttsb_device_error_flag(int* h_flag) {
    return guarded_call([&]() -> int {
    TTSB_REQUIRE(h_flag != nullptr, "null argument");
    *h_flag = 0;
    GlobalRuntime& g = global_runtime();
    // every device this process has launched on; the first non-zero flag wins
    int prev = 0;
    TTSB_CHECK_CUDA(cudaGetDevice(&prev));
    for (int dev = 0; dev < kMaxDevices; ++dev) {
        if (!g.err_flag[dev]) continue;
        int v = 0;
        TTSB_CHECK_CUDA(cudaSetDevice(dev));
        TTSB_CHECK_CUDA(cudaDeviceSynchronize());
        TTSB_CHECK_CUDA(cudaMemcpy(&v, g.err_flag[dev], sizeof(int), cudaMemcpyDeviceToHost));
        if (v != 0 && *h_flag == 0) *h_flag = v;
    }
    TTSB_CHECK_CUDA(cudaSetDevice(prev));
    return 0;
    });
}

int ttsb_prof_enable(int on) { return prof_enable(on); }
int ttsb_prof_n_tags(void) { return PROF_N_TAGS; }
int ttsb_prof_collect(double* h_ms_by_tag, int n_tags) {
    return guarded_call([&]() -> int {
    TTSB_REQUIRE(h_ms_by_tag != nullptr && n_tags > 0, "null argument");
    return prof_collect(h_ms_by_tag, n_tags);
    });
}

int ttsb_debug_set_timeline(void* d_buf) {
    return guarded_call([&]() -> int {
    global_runtime().timeline = static_cast<long long*>(d_buf);
    return 0;
    });
}

int ttsb_conv1d_create(int kind, int cin, int cout, int ksize, int dilation, int stride,
                       const float* h_weight, const float* h_bias, int device, ttsb_conv1d_t** out) {
    return guarded_call([&]() -> int {
    TTSB_REQUIRE(h_weight && out, "null argument");
    TTSB_DEVICE_GUARD(device);
    ttsb_conv1d* h = new ttsb_conv1d();
    h->device = device;
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
    });
}
void ttsb_conv1d_destroy(ttsb_conv1d_t* h) {
    if (!h) return;
    conv_layer_destroy(h->layer);
    delete h;
}
int ttsb_conv1d_cin_pad(const ttsb_conv1d_t* h) { return h ? h->layer.cin : 0; }

int ttsb_conv1d_forward(ttsb_conv1d_t* h, const void* d_in, int B, int T, const void* d_residual,
                        float act_slope, const int32_t* d_lens, void* d_out, void* stream) {
    return guarded_call([&]() -> int {
    TTSB_REQUIRE(h && d_in && d_out, "null argument");
    TTSB_REQUIRE(act_slope <= 1.f, "leaky-relu slope must be <= 1 (the epilogue evaluates it as max(x, slope * x))");
    TTSB_DEVICE_GUARD(h->device);
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
    });
}

int ttsb_convpair_create(int channels, int ksize, int dilation, const float* h_w1, const float* h_b1,
                         const float* h_w2, const float* h_b2, int device, ttsb_convpair_t** out) {
    return guarded_call([&]() -> int {
    TTSB_REQUIRE(h_w1 && h_b1 && h_w2 && h_b2 && out, "null argument");
    TTSB_DEVICE_GUARD(device);
    ttsb_convpair* h = new ttsb_convpair();
    h->device = device;
    int st = make_conv1d_layer(h->c1, h_w1, h_b1, channels, channels, ksize, dilation, channels, 0);
    if (st == 0) st = make_conv1d_layer(h->c2, h_w2, h_b2, channels, channels, ksize, 1, channels, 0);
    if (st != 0) { ttsb_convpair_destroy(h); return st; }
    h->plan = conv_pair_plan(h->c1, h->c2);
    *out = h;
    return 0;
    });
}
void ttsb_convpair_destroy(ttsb_convpair_t* h) {
    if (!h) return;
    conv_layer_destroy(h->c1);
    conv_layer_destroy(h->c2);
    delete h;
}
int ttsb_convpair_plan(const ttsb_convpair_t* h, int* out8) {
    return guarded_call([&]() -> int {
    TTSB_REQUIRE(h && out8, "null argument");
    const ConvPairPlan& p = h->plan;
    const int v[8] = {p.ok, p.m_out, p.x_slots, p.tt_slots, p.w2_resident, p.b_stages, p.tmem_cols,
                      static_cast<int>(p.smem_bytes)};
    for (int i = 0; i < 8; ++i) out8[i] = v[i];
    return 0;
    });
}
int ttsb_convpair_forward(ttsb_convpair_t* h, const void* d_x, int B, int T, const int32_t* d_lens, float slope,
                          void* d_out, void* stream) {
    return guarded_call([&]() -> int {
    TTSB_REQUIRE(h && d_x && d_out, "null argument");
    TTSB_REQUIRE(h->plan.ok, "this (channels, ksize, dilation) has no fused plan");
    TTSB_DEVICE_GUARD(h->device);
    ConvRuntime rt;
    TTSB_PROPAGATE(get_conv_runtime(0, rt));
    EpiParams e;
    e.lens = d_lens; e.len_mul = 1;
    e.out_raw = static_cast<__half*>(d_out); e.ld_raw = h->plan.C;
    e.act_slope = slope;
    return conv_pair_forward(h->c1, h->c2, h->plan, rt, static_cast<const __half*>(d_x), B, T, slope, e,
                             static_cast<cudaStream_t>(stream));
    });
}

int ttsb_convpair_forward_act(ttsb_convpair_t* h, const void* d_x, int B, int T, const int32_t* d_lens, float slope,
                          void* d_out, void* stream) {
    return guarded_call([&]() -> int {
    TTSB_REQUIRE(h && d_x && d_out, "null argument");
    TTSB_REQUIRE(h->plan.ok, "this (channels, ksize, dilation) has no fused plan");
    TTSB_DEVICE_GUARD(h->device);
    ConvRuntime rt;
    TTSB_PROPAGATE(get_conv_runtime(0, rt));
    EpiParams e;
    e.lens = d_lens; e.len_mul = 1;
    e.out_act = static_cast<__half*>(d_out); e.ld_act = h->plan.C;
    e.act_slope = slope;
    return conv_pair_forward(h->c1, h->c2, h->plan, rt, static_cast<const __half*>(d_x), B, T, slope, e,
                             static_cast<cudaStream_t>(stream), 1);
    });
}

}  // extern "C"
