// Helpers shared by the model-level translation units: process-wide runtime switches,
// named-tensor lookup, workspace carving.
#pragma once
#include <map>
#include <string>
#include <vector>
#include "../../include/ttsb200.h"
#include "conv.cuh"
#include "kernels.cuh"

namespace ttsb {

struct GlobalRuntime {
    int impl = IMPL_TC;
    int tc_version = 2;
    int desc_mode = 0;  // verified on B200 (profiles/r01_s1_probe_conv.json): base_offset stays 0
    // per device (a handle on cuda:1 must not be handed cuda:0's flag / scratch)
    int* err_flag[kMaxDevices] = {};
    float* simt_scratch[kMaxDevices] = {};
    size_t simt_scratch_elems[kMaxDevices] = {};
    long long* timeline = nullptr;
};
GlobalRuntime& global_runtime();
// Returns a ConvRuntime view; for the SIMT path makes sure the fp32 scratch holds `elems` floats.
int get_conv_runtime(size_t simt_elems, ConvRuntime& rt);

class TensorTable {
public:
    TensorTable(const ttsb_tensor_t* w, int n) {
        for (int i = 0; i < n; ++i) map_[w[i].name] = &w[i];
    }
    const ttsb_tensor_t* find(const std::string& name) const {
        auto it = map_.find(name);
        return it == map_.end() ? nullptr : it->second;
    }
    static size_t numel(const ttsb_tensor_t* t) {
        size_t n = 1;
        for (int i = 0; i < t->ndim; ++i) n *= static_cast<size_t>(t->shape[i]);
        return n;
    }
private:
    std::map<std::string, const ttsb_tensor_t*> map_;
};

#define TTSB_GET_TENSOR(var, table, name_expr, nd)                                       \
    const ttsb_tensor_t* var = (table).find(name_expr);                                  \
    TTSB_REQUIRE(var != nullptr && var->ndim == (nd), std::string("missing or mis-shaped tensor ") + (name_expr))

// nn.Conv1d weight [Cout,Cin,K] -> dense logical [Cout][K][Cin] + centred dilated tap offsets
int make_conv1d_layer(ConvLayer& L, const float* w, const float* bias, int cout, int cin, int k,
                      int dilation, int cin_stored, int n_tile_hint);
// nn.ConvTranspose1d weight [Cin,Cout,K=2s], stride s, padding s/2 -> two-tap polyphase layer with
// N = s*Cout columns (column p*Cout+co = output phase p, channel co)
int make_convT1d_layer(ConvLayer& L, const float* w, const float* bias, int cin, int cout, int k,
                       int stride);
int upload_f32(const float* h, size_t n, float** d);
int prof_enable(int on);
int prof_collect(double* ms_by_tag, int n_tags);

struct Carver {
    uint8_t* base;
    size_t off = 0;
    explicit Carver(void* p) : base(static_cast<uint8_t*>(p)) {}
    template <class T>
    T* take(size_t n) {
        off = (off + 255) & ~static_cast<size_t>(255);
        T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
        off += n * sizeof(T);
        return p;
    }
};

}  // namespace ttsb
