// Context object behind the opaque rst_ctx handle: layer plan, weight registry, workspaces.
#pragma once

#include <memory>

#include "rst_internal.cuh"

namespace rst {

struct LayerDesc {          // one conv of the transfer net
    std::string name;       // "contract_start", "residual_block_0/conv0", "expand_last", ...
    int ci = 0, co = 0, k = 3, stride = 1;
    int hi = 0, wi = 0, ho = 0, wo = 0;
    int pad_t = 0, pad_l = 0;
    bool transposed = false;
};

struct Weight {
    std::string name;
    std::vector<int64_t> shape;
    std::vector<float> host;
    float* dev = nullptr;
    bool set = false;
    int64_t elems() const {
        int64_t n = 1;
        for (auto d : shape) n *= d;
        return n;
    }
};

struct MbBlock {            // MobileNetV3Small inverted residual block
    std::string prefix;
    int block_id = 0, cin = 0, cexp = 0, cout = 0, k = 3, s = 1, se = 0, act = ACT_RELU;
};

struct Tap {
    float* dev = nullptr;
    int64_t elems = 0;
};

struct ProfileGroup {
    double total_ms = 0.0;
    int64_t launches = 0;
};

struct Bf16State;           // defined in transfer_bf16.cu
struct Tf32Conv3x3;         // halo_gemm.cuh

}  // namespace rst

struct rst_ctx {
    rst_config cfg{};
    int device = 0;
    mutable std::string err;
    bool committed = false;

    // ---- plan (styleTransfer.py:213-332) ----
    int n_contract = 0, n_expand = 0, num_style_params = 0;
    int bott_h = 0, bott_w = 0;
    std::vector<rst::LayerDesc> contract, residual, expand;
    std::vector<rst::MbBlock> mb_blocks;
    int mb_last = 0;            // channels of Conv_1 (576)
    int feat_c = 0;             // predictor feature channels

    // ---- weights ----
    std::vector<rst::Weight> weights;
    std::map<std::string, int> weight_index;
    float* weight_arena = nullptr;          // when set (trainer), every Weight::dev points into this one allocation
    std::map<std::string, float*> folded;   // "<bn prefix>/scale", "<bn prefix>/shift" device arrays

    // ---- workspaces (fp32 path) ----
    float* act[3] = {nullptr, nullptr, nullptr};
    int64_t act_elems = 0;
    double* stats = nullptr;                // (max_batch, max C, 2)
    float* w_pyramid = nullptr;             // concat + mips storage
    std::vector<std::pair<int, float*>> mips;   // (width, device ptr) of the current forward
    // predictor
    float* pact[4] = {nullptr, nullptr, nullptr, nullptr};
    int64_t pact_elems = 0;
    float* pvec[3] = {nullptr, nullptr, nullptr};   // small (B, C) vectors (GAP / SE)
    // host-API staging
    float* st_content = nullptr; float* st_params = nullptr; float* st_weights = nullptr;
    float* st_out = nullptr; float* st_style = nullptr;
    float* cvt_content = nullptr; float* cvt_out = nullptr;    // fp32 path only: typed (fp16 in / uint8 out) calls convert through these
    cudaStream_t own_stream = nullptr;
    // double-buffered asynchronous host pipeline (rst_transfer_submit_host / rst_transfer_wait)
    struct Pipe {
        bool ready = false;
        float *content[2] = {nullptr, nullptr}, *params[2] = {nullptr, nullptr}, *weights[2] = {nullptr, nullptr},
              *out[2] = {nullptr, nullptr};
        cudaStream_t s_in = nullptr, s_out = nullptr;
        cudaEvent_t in_done[2] = {nullptr, nullptr}, comp_done[2] = {nullptr, nullptr}, out_done[2] = {nullptr, nullptr};
        bool busy[2] = {false, false};
        int64_t next = 0;
    } pipe;

    // ---- fp32 path: the residual 3x3 convolutions as error-compensated split-tf32 tcgen05 GEMMs (conv_tf32.cu) ----
    std::vector<std::shared_ptr<rst::Tf32Conv3x3>> res_tf32;   // one per residual conv; empty = CUDA-core kernels
    float* tf32_scratch = nullptr;                              // [x_hi | x_lo] expansion of one conv input
    int num_sms = 0;

    // ---- bf16 tensor-core path ----
    std::shared_ptr<rst::Bf16State> bf16;

    // ---- CUDA-graph cache of whole forwards (launch-bound inner loop: ~30 kernels per batch) ----
    struct GraphEntry {
        int batch = 0; const void *content = nullptr, *params = nullptr, *weights = nullptr, *out = nullptr;
        int content_dtype = 0, out_dtype = 0;
        cudaGraphExec_t exec = nullptr; int64_t launches = 0;
    };
    std::vector<GraphEntry> graphs;
    std::vector<GraphEntry> graph_candidates;   // buffer sets seen once (captured when they come back)
    bool use_graphs = true;

    // ---- debug / accounting ----
    bool keep_taps = false;
    std::map<std::string, rst::Tap> taps;
    int64_t launches = 0;
    bool profiling = false;
    std::map<std::string, rst::ProfileGroup> profile;
    std::vector<std::tuple<std::string, cudaEvent_t, cudaEvent_t>> pending_events;
    std::vector<cudaEvent_t> event_pool;
    cudaEvent_t take_event() {
        if (!event_pool.empty()) { cudaEvent_t e = event_pool.back(); event_pool.pop_back(); return e; }
        cudaEvent_t e = nullptr;
        cudaEventCreate(&e);
        return e;
    }

    const rst::Weight* find_weight(const std::string& n) const {
        auto it = weight_index.find(n);
        return it == weight_index.end() ? nullptr : &weights[it->second];
    }
    float* wdev(const std::string& n) const {
        const rst::Weight* w = find_weight(n);
        return w ? w->dev : nullptr;
    }
};

namespace rst {

// RAII helper used around each kernel launch: counts launches and (optionally) times the group.
struct LaunchScope {
    rst_ctx* ctx; cudaStream_t s; cudaEvent_t e0 = nullptr, e1 = nullptr; const char* group;
    LaunchScope(rst_ctx* c, cudaStream_t st, const char* g, int n = 1) : ctx(c), s(st), group(g) {
        ctx->launches += n;
        if (ctx->profiling) {
            e0 = ctx->take_event();
            e1 = ctx->take_event();
            cudaEventRecord(e0, s);
        }
    }
    ~LaunchScope() {
        if (e0) {
            cudaEventRecord(e1, s);
            ctx->pending_events.emplace_back(std::string(group), e0, e1);
        }
    }
};

int fail(rst_ctx* ctx, int code, const std::string& msg);
int cuda_fail(rst_ctx* ctx, cudaError_t e, const char* what);
void record_tap(rst_ctx* ctx, const std::string& name, const void* dev, int64_t elems, bool is_bf16, cudaStream_t s);

// implemented in rst_api.cu, shared with the bf16 path
int build_mips(rst_ctx* c, const float* d_style_weights, int batch, cudaStream_t s);
const float* mip_for_width(rst_ctx* c, int width);
int fp32_contract_stage(rst_ctx* c, const float* d_content, int batch, cudaStream_t s, float** out, int* free_idx);
int fp32_expand_stage(rst_ctx* c, float* x, float* t1, const float* d_style_params, int cursor, float* d_out, int batch,
                      cudaStream_t s);

// implemented in transfer_bf16.cu
int bf16_create(rst_ctx* ctx);
int bf16_commit(rst_ctx* ctx);
int bf16_transfer_forward(rst_ctx* ctx, const void* d_content, int content_dtype, const float* d_style_params,
                          const float* d_style_weights, void* d_out, int out_dtype, int batch, cudaStream_t s);
cudaError_t launch_bf16_to_f32(const void* src, float* dst, long long n, cudaStream_t s);

}  // namespace rst

#define RST_CUDA(ctx, expr)                                                    \
    do {                                                                       \
        cudaError_t _e = (expr);                                               \
        if (_e != cudaSuccess) return rst::cuda_fail((ctx), _e, #expr);        \
    } while (0)
