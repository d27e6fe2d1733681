// C ABI (include/rst_b200.h): context, layer plan, weight registry, fp32 forward, style predictor.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <tuple>

#include "train_kernels.cuh"
#include "rst_ctx.h"
#include "halo_gemm.cuh"

using namespace rst;

static thread_local std::string g_create_error;

namespace rst {

int fail(rst_ctx* ctx, int code, const std::string& msg) {
    if (ctx) ctx->err = msg;
    else g_create_error = msg;
    return code;
}
int cuda_fail(rst_ctx* ctx, cudaError_t e, const char* what) {
    return fail(ctx, RST_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

__global__ void bf16_to_f32_kernel(const __nv_bfloat16* __restrict__ src, float* __restrict__ dst, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = __bfloat162float(src[i]);
}
cudaError_t launch_bf16_to_f32(const void* src, float* dst, long long n, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    bf16_to_f32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>((const __nv_bfloat16*)src, dst, n);
    return cudaGetLastError();
}

void record_tap(rst_ctx* ctx, const std::string& name, const void* dev, int64_t elems, bool is_bf16,
                cudaStream_t s) {
    if (!ctx->keep_taps) return;
    Tap& t = ctx->taps[name];
    if (t.elems != elems) {
        if (t.dev) cudaFree(t.dev);
        t.dev = nullptr;
        cudaMalloc(&t.dev, (size_t)elems * sizeof(float));
        t.elems = elems;
    }
    if (is_bf16) launch_bf16_to_f32(dev, t.dev, elems, s);
    else cudaMemcpyAsync(t.dev, dev, (size_t)elems * sizeof(float), cudaMemcpyDeviceToDevice, s);
}

}  // namespace rst

// ------------------------------------------------------------------------------------------------
// plan
// ------------------------------------------------------------------------------------------------
static const int kContractFilters[4] = {16, 32, 32, 32};               // styleTransfer.py:218-223
static const int kExpandFilters[8] = {32, 16, 8, 4, 3, 3, 3, 3};       // styleTransfer.py:247-256

static void add_weight(rst_ctx* c, const std::string& name, std::vector<int64_t> shape) {
    Weight w;
    w.name = name;
    w.shape = std::move(shape);
    c->weight_index[name] = (int)c->weights.size();
    c->weights.push_back(std::move(w));
}
static void add_bn(rst_ctx* c, const std::string& prefix, int ch) {
    for (const char* v : {"gamma", "beta", "moving_mean", "moving_variance"})
        add_weight(c, prefix + "/" + v, {ch});
}

static int mb_depth(double v) {   // keras mobilenet_v3._depth, divisor 8
    int nv = std::max(8, (int)(v + 4.0) / 8 * 8);
    if (nv < 0.9 * v) nv += 8;
    return nv;
}

static int build_plan(rst_ctx* c) {
    const rst_config& g = c->cfg;
    const bool predictor_only = g.in_h == 0;
    if (predictor_only) {
        if (g.extractor == RST_EXTRACTOR_NONE || g.predictor_num_params <= 0 || g.max_batch < 1)
            return fail(nullptr, RST_ERR_INVALID, "rst_create: predictor-only context needs an extractor and predictor_num_params");
        c->num_style_params = g.predictor_num_params;
    } else {
    if (g.in_h <= 0 || g.in_w <= 0 || g.in_c <= 0 || g.out_h <= 0 || g.out_w <= 0 || g.bottleneck_res_y <= 0 ||
        g.bottleneck_num_filters <= 0 || g.num_styles < 1 || g.max_batch < 1)
        return fail(nullptr, RST_ERR_INVALID, "rst_create: non-positive dimension in rst_config");
    if (g.num_styles > 2)
        return fail(nullptr, RST_ERR_INVALID,
                    "rst_create: only 1 or 2 styles blend (styleTransfer.py:38 passes >2 styles through unblended)");
    // The reference's exact expression, including its floating-point rounding.
    c->n_contract = (int)std::ceil(std::log2((double)g.in_h) - std::log2((double)g.bottleneck_res_y));
    if (c->n_contract < 1 || c->n_contract > 4)
        return fail(nullptr, RST_ERR_INVALID, "rst_create: contract block count outside 1..4");
    int bh = (int)(g.in_h * std::pow(2.0, -c->n_contract));
    int bw = (int)(g.in_w * std::pow(2.0, -c->n_contract));
    c->n_expand = (int)std::ceil(std::log2((double)g.out_h) - std::log2((double)bh));
    if (c->n_expand < 1 || c->n_expand > 8)
        return fail(nullptr, RST_ERR_INVALID, "rst_create: expand block count outside 1..8");
    const int F = g.bottleneck_num_filters;

    int h = g.in_h, w = g.in_w, ci = g.in_c;
    auto conv_layer = [&](const std::string& name, int co, int k, int s) {
        LayerDesc L;
        L.name = name; L.ci = ci; L.co = co; L.k = k; L.stride = s; L.hi = h; L.wi = w;
        L.ho = tf_same(h, k, s, &L.pad_t);
        L.wo = tf_same(w, k, s, &L.pad_l);
        h = L.ho; w = L.wo; ci = co;
        return L;
    };
    c->contract.push_back(conv_layer("contract_start", 32, 9, 1));
    for (int i = 0; i < c->n_contract; ++i)
        c->contract.push_back(conv_layer("contract_" + std::to_string(i), kContractFilters[i], 3, 2));
    if (h != bh || w != bw)
        return fail(nullptr, RST_ERR_INVALID, "rst_create: input size is not divisible down to the bottleneck");
    c->bott_h = h; c->bott_w = w;
    for (int b = 0; b < 5; ++b)
        for (int i = 0; i < 2; ++i)
            c->residual.push_back(conv_layer("residual_block_" + std::to_string(b) + "/conv" + std::to_string(i), F, 3, 1));
    auto convt_layer = [&](const std::string& name, int co, int k, int s) {
        LayerDesc L;
        L.name = name; L.ci = ci; L.co = co; L.k = k; L.stride = s; L.hi = h; L.wi = w; L.transposed = true;
        L.ho = h * s; L.wo = w * s;
        tf_same(L.ho, k, s, &L.pad_t);
        tf_same(L.wo, k, s, &L.pad_l);
        h = L.ho; w = L.wo; ci = co;
        return L;
    };
    for (int i = 0; i < c->n_expand; ++i)
        c->expand.push_back(convt_layer("expand_" + std::to_string(i), kExpandFilters[i], 3, 2));
    c->expand.push_back(convt_layer("expand_last", 3, 9, 1));
    if (h != g.out_h || w != g.out_w)
        return fail(nullptr, RST_ERR_INVALID, "rst_create: expand blocks do not reach output_shape");
    c->num_style_params = 5 * 4 * F;
    for (auto& L : c->expand) c->num_style_params += 2 * L.co;
    }

    // ---- weight registry ----
    for (auto& L : c->contract) {
        add_weight(c, L.name + "/conv/kernel", {L.k, L.k, L.ci, L.co});
        add_weight(c, L.name + "/conv/bias", {L.co});
        add_bn(c, L.name + "/bn", L.co);
    }
    for (auto& L : c->residual) {
        add_weight(c, L.name + "/kernel", {L.k, L.k, L.ci, L.co});
        add_weight(c, L.name + "/bias", {L.co});
    }
    for (auto& L : c->expand) {
        add_weight(c, L.name + "/conv/kernel", {L.k, L.k, L.co, L.ci});   // Conv2DTranspose (kh,kw,out,in)
        add_weight(c, L.name + "/conv/bias", {L.co});
    }

    // ---- style predictor (stylePrediction.py:25-75) ----
    if (g.extractor != RST_EXTRACTOR_NONE) {
        if (g.style_h <= 0 || g.style_w <= 0)
            return fail(nullptr, RST_ERR_INVALID, "rst_create: style_h/style_w required with an extractor");
        if (g.extractor == RST_EXTRACTOR_DUMMY) {
            add_weight(c, "dummy_conv/kernel", {9, 9, 3, 1});
            add_weight(c, "dummy_conv/bias", {1});
            c->feat_c = 1;
        } else if (g.extractor == RST_EXTRACTOR_MOBILE_NET) {
            struct Row { double e; int co, k, s; bool se; int act; };
            static const Row rows[11] = {
                {1, 16, 3, 2, true, ACT_RELU},        {72. / 16, 24, 3, 2, false, ACT_RELU},
                {88. / 24, 24, 3, 1, false, ACT_RELU}, {4, 40, 5, 2, true, ACT_HSWISH},
                {6, 40, 5, 1, true, ACT_HSWISH},      {6, 40, 5, 1, true, ACT_HSWISH},
                {3, 48, 5, 1, true, ACT_HSWISH},      {3, 48, 5, 1, true, ACT_HSWISH},
                {6, 96, 5, 2, true, ACT_HSWISH},      {6, 96, 5, 1, true, ACT_HSWISH},
                {6, 96, 5, 1, true, ACT_HSWISH}};
            add_weight(c, "mobilenet/Conv/kernel", {3, 3, 3, 16});
            add_bn(c, "mobilenet/Conv/BatchNorm", 16);
            int cin = 16;
            for (int b = 0; b < 11; ++b) {
                MbBlock m;
                m.block_id = b;
                m.prefix = b == 0 ? "mobilenet/expanded_conv" : "mobilenet/expanded_conv_" + std::to_string(b);
                m.cin = cin; m.cexp = mb_depth(cin * rows[b].e); m.cout = rows[b].co;
                m.k = rows[b].k; m.s = rows[b].s; m.act = rows[b].act;
                m.se = rows[b].se ? mb_depth(m.cexp * 0.25) : 0;
                if (b) {
                    add_weight(c, m.prefix + "/expand/kernel", {1, 1, m.cin, m.cexp});
                    add_bn(c, m.prefix + "/expand/BatchNorm", m.cexp);
                }
                add_weight(c, m.prefix + "/depthwise/depthwise_kernel", {m.k, m.k, m.cexp, 1});
                add_bn(c, m.prefix + "/depthwise/BatchNorm", m.cexp);
                if (m.se) {
                    add_weight(c, m.prefix + "/squeeze_excite/Conv/kernel", {1, 1, m.cexp, m.se});
                    add_weight(c, m.prefix + "/squeeze_excite/Conv/bias", {m.se});
                    add_weight(c, m.prefix + "/squeeze_excite/Conv_1/kernel", {1, 1, m.se, m.cexp});
                    add_weight(c, m.prefix + "/squeeze_excite/Conv_1/bias", {m.cexp});
                }
                add_weight(c, m.prefix + "/project/kernel", {1, 1, m.cexp, m.cout});
                add_bn(c, m.prefix + "/project/BatchNorm", m.cout);
                cin = m.cout;
                c->mb_blocks.push_back(m);
            }
            c->mb_last = mb_depth(cin * 6);
            add_weight(c, "mobilenet/Conv_1/kernel", {1, 1, cin, c->mb_last});
            add_bn(c, "mobilenet/Conv_1/BatchNorm", c->mb_last);
            c->feat_c = c->mb_last;
        } else {
            return fail(nullptr, RST_ERR_INVALID, "rst_create: not a valid value for feature_extractor");
        }
        add_weight(c, "StylePredictor/kernel", {1, 1, c->feat_c, 100});
        add_weight(c, "StylePredictor/bias", {100});
        add_weight(c, "StyleNormPredictor/kernel", {1, 1, 100, c->num_style_params});
        add_weight(c, "StyleNormPredictor/bias", {c->num_style_params});
    }
    return RST_OK;
}

// ------------------------------------------------------------------------------------------------
// lifetime
// ------------------------------------------------------------------------------------------------
extern "C" const char* rst_version(void) { return "rst_b200 0.2 (sm_100a)"; }

// Page-locked host memory for frame / image buffers handed to the *_host entry points (asynchronous DMA needs it).
// write_combined: uncached on the CPU side -- right for buffers the host only WRITES (captured G-buffers on their way to the GPU).
extern "C" int rst_host_alloc(void** h_ptr, uint64_t bytes, int write_combined) {
    if (!h_ptr) return RST_ERR_INVALID;
    *h_ptr = nullptr;
    cudaError_t e = cudaHostAlloc(h_ptr, bytes ? bytes : 16, cudaHostAllocPortable | (write_combined ? cudaHostAllocWriteCombined : 0));
    if (e != cudaSuccess) { g_create_error = std::string("rst_host_alloc: ") + cudaGetErrorString(e); return RST_ERR_CUDA; }
    return RST_OK;
}
extern "C" int rst_host_free(void* h_ptr) {
    if (!h_ptr) return RST_OK;
    return cudaFreeHost(h_ptr) == cudaSuccess ? RST_OK : RST_ERR_CUDA;
}

// crc32c (Castagnoli, reflected 0x82F63B78) of a host buffer, slicing-by-8: the checksum of TensorFlow's tensor bundles
// (tensorflow/core/lib/hash/crc32c) for the checkpoint reader / writer, whose pure-Python loop is too slow for 100 MB shards.
extern "C" uint32_t rst_host_crc32c(const void* data, uint64_t n) {
    static uint32_t T[8][256];
    static bool ready = false;
    if (!ready) {
        for (uint32_t i = 0; i < 256; ++i) {
            uint32_t c = i;
            for (int k = 0; k < 8; ++k) c = (c >> 1) ^ (0x82F63B78u & (0u - (c & 1u)));
            T[0][i] = c;
        }
        for (uint32_t i = 0; i < 256; ++i)
            for (int t = 1; t < 8; ++t) T[t][i] = (T[t - 1][i] >> 8) ^ T[0][T[t - 1][i] & 0xFF];
        ready = true;
    }
    const uint8_t* p = static_cast<const uint8_t*>(data);
    uint32_t c = 0xFFFFFFFFu;
    while (n >= 8) {
        uint32_t lo, hi;
        memcpy(&lo, p, 4); memcpy(&hi, p + 4, 4);
        lo ^= c;
        c = T[7][lo & 0xFF] ^ T[6][(lo >> 8) & 0xFF] ^ T[5][(lo >> 16) & 0xFF] ^ T[4][lo >> 24] ^
            T[3][hi & 0xFF] ^ T[2][(hi >> 8) & 0xFF] ^ T[1][(hi >> 16) & 0xFF] ^ T[0][hi >> 24];
        p += 8; n -= 8;
    }
    while (n--) c = T[0][(c ^ *p++) & 0xFF] ^ (c >> 8);
    return c ^ 0xFFFFFFFFu;
}

static void free_ctx(rst_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->weight_arena) cudaFree(c->weight_arena);
    else for (auto& w : c->weights) if (w.dev) cudaFree(w.dev);
    for (auto& kv : c->folded) if (kv.second) cudaFree(kv.second);
    for (auto& a : c->act) if (a) cudaFree(a);
    for (auto& a : c->pact) if (a) cudaFree(a);
    for (auto& a : c->pvec) if (a) cudaFree(a);
    if (c->stats) cudaFree(c->stats);
    if (c->w_pyramid) cudaFree(c->w_pyramid);
    c->res_tf32.clear();
    if (c->tf32_scratch) cudaFree(c->tf32_scratch);
    for (float* p : {c->st_content, c->st_params, c->st_weights, c->st_out, c->st_style, c->cvt_content, c->cvt_out}) if (p) cudaFree(p);
    for (auto& kv : c->taps) if (kv.second.dev) cudaFree(kv.second.dev);
    for (auto e : c->event_pool) cudaEventDestroy(e);
    for (auto& g : c->graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
    c->bf16.reset();
    if (c->pipe.ready) {
        for (int i = 0; i < 2; ++i) {
            for (float* q : {c->pipe.content[i], c->pipe.params[i], c->pipe.weights[i], c->pipe.out[i]}) if (q) cudaFree(q);
            for (cudaEvent_t e : {c->pipe.in_done[i], c->pipe.comp_done[i], c->pipe.out_done[i]}) if (e) cudaEventDestroy(e);
        }
        if (c->pipe.s_in) cudaStreamDestroy(c->pipe.s_in);
        if (c->pipe.s_out) cudaStreamDestroy(c->pipe.s_out);
    }
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
}

extern "C" int rst_create(const rst_config* cfg, int device, rst_ctx** out_ctx) {
    if (!cfg || !out_ctx) return fail(nullptr, RST_ERR_INVALID, "rst_create: null argument");
    *out_ctx = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, RST_ERR_CUDA, "rst_create: no CUDA device is visible; this library has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(nullptr, RST_ERR_INVALID, "rst_create: device index out of range");
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    if (prop.major != 10)
        return fail(nullptr, RST_ERR_CUDA, "rst_create: built for sm_100a (B200) only, found another architecture");
    rst_ctx* c = new rst_ctx();
    c->cfg = *cfg;
    if (c->cfg.in_h == 0 && c->cfg.num_styles < 1) c->cfg.num_styles = 1;
    cfg = &c->cfg;
    c->device = device;
    int rc = build_plan(c);
    if (rc != RST_OK) { delete c; return rc; }
    cudaSetDevice(device);

    auto alloc = [&](void** p, size_t bytes) -> bool {
        cudaError_t ee = cudaMalloc(p, bytes ? bytes : 16);
        if (ee != cudaSuccess) { g_create_error = std::string("rst_create: cudaMalloc: ") + cudaGetErrorString(ee); return false; }
        return true;
    };
    const int B = cfg->max_batch;
    bool ok = true;
    int64_t max_act = 0;
    int max_c = 3;
    for (auto* v : {&c->contract, &c->residual, &c->expand})
        for (auto& L : *v) {
            max_act = std::max<int64_t>(max_act, (int64_t)L.ho * L.wo * L.co);
            max_c = std::max(max_c, L.co);
        }
    c->act_elems = max_act * B;
    for (int i = 0; i < 3 && ok; ++i) ok = alloc((void**)&c->act[i], (size_t)c->act_elems * sizeof(float));
    ok = ok && alloc((void**)&c->stats, (size_t)B * std::max(max_c, 1024) * 2 * sizeof(double));
    if (cfg->num_styles > 1)   // concat level + pyramid (< 1/3 extra)
        ok = ok && alloc((void**)&c->w_pyramid, (size_t)B * cfg->out_h * cfg->out_w * 2 * sizeof(float) * 4 / 3 + 4096);
    if (cfg->extractor != RST_EXTRACTOR_NONE) {
        // largest predictor activation: MobileNet block 1 expand at 1/4 resolution x 72 ch (or stem at 1/2 x 16)
        int64_t sh = cfg->style_h, sw = cfg->style_w;
        int64_t m = std::max<int64_t>(((sh + 1) / 2) * ((sw + 1) / 2) * 16, ((sh + 3) / 4) * ((sw + 3) / 4) * 72);
        m = std::max<int64_t>(m, ((sh + 4) / 5) * ((sw + 4) / 5));
        c->pact_elems = m * B;
        for (int i = 0; i < 4 && ok; ++i) ok = alloc((void**)&c->pact[i], (size_t)c->pact_elems * sizeof(float));
        for (int i = 0; i < 3 && ok; ++i)
            ok = alloc((void**)&c->pvec[i], (size_t)B * std::max(1024, c->num_style_params) * sizeof(float));
        ok = ok && alloc((void**)&c->st_style, (size_t)B * (cfg->num_styles + 1) * sh * sw * 3 * sizeof(float));
    }
    ok = ok && alloc((void**)&c->st_content, (size_t)B * cfg->in_h * cfg->in_w * cfg->in_c * sizeof(float));
    ok = ok && alloc((void**)&c->st_params, (size_t)B * cfg->num_styles * c->num_style_params * sizeof(float));
    ok = ok && alloc((void**)&c->st_out, (size_t)B * cfg->out_h * cfg->out_w * 3 * sizeof(float));
    if (cfg->num_styles > 1)
        ok = ok && alloc((void**)&c->st_weights, (size_t)B * cfg->out_h * cfg->out_w * (cfg->num_styles - 1) * sizeof(float));
    if (ok && cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
        g_create_error = "rst_create: cudaStreamCreate failed";
        ok = false;
    }
    if (!ok) { free_ctx(c); return RST_ERR_CUDA; }
    if (cfg->precision != RST_PRECISION_FP32 && cfg->precision != RST_PRECISION_BF16) {
        free_ctx(c);
        return fail(nullptr, RST_ERR_INVALID, "rst_create: unknown precision");
    }
    if (cfg->precision == RST_PRECISION_BF16 && cfg->in_h != 0) {
        rc = bf16_create(c);
        if (rc != RST_OK) { g_create_error = c->err; free_ctx(c); return rc; }
    }
    *out_ctx = c;
    return RST_OK;
}

extern "C" int rst_destroy(rst_ctx* ctx) {
    free_ctx(ctx);
    return RST_OK;
}

extern "C" const char* rst_last_error(const rst_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }
extern "C" int rst_num_style_params(const rst_ctx* ctx) { return ctx ? ctx->num_style_params : -1; }
extern "C" int rst_num_contract_blocks(const rst_ctx* ctx) { return ctx ? ctx->n_contract : -1; }
extern "C" int rst_num_expand_blocks(const rst_ctx* ctx) { return ctx ? ctx->n_expand : -1; }
extern "C" int rst_weight_count(const rst_ctx* ctx) { return ctx ? (int)ctx->weights.size() : -1; }
extern "C" const char* rst_weight_name(const rst_ctx* ctx, int i) {
    if (!ctx || i < 0 || i >= (int)ctx->weights.size()) return nullptr;
    return ctx->weights[i].name.c_str();
}
extern "C" int rst_weight_shape(const rst_ctx* ctx, int i, int64_t* shape4, int* ndim) {
    if (!ctx || i < 0 || i >= (int)ctx->weights.size() || !shape4 || !ndim) return RST_ERR_INVALID;
    const Weight& w = ctx->weights[i];
    *ndim = (int)w.shape.size();
    for (size_t k = 0; k < w.shape.size() && k < 4; ++k) shape4[k] = w.shape[k];
    return RST_OK;
}

extern "C" int rst_set_weight(rst_ctx* ctx, const char* name, const float* h_data, const int64_t* shape, int ndim) {
    if (!ctx || !name || !h_data || !shape) return fail(ctx, RST_ERR_INVALID, "rst_set_weight: null argument");
    auto it = ctx->weight_index.find(name);
    if (it == ctx->weight_index.end()) return fail(ctx, RST_ERR_INVALID, std::string("rst_set_weight: unknown variable ") + name);
    Weight& w = ctx->weights[it->second];
    bool same = ndim == (int)w.shape.size();
    for (int k = 0; same && k < ndim; ++k) same = shape[k] == w.shape[k];
    if (!same) return fail(ctx, RST_ERR_INVALID, std::string("rst_set_weight: shape mismatch for ") + name);
    w.host.assign(h_data, h_data + w.elems());
    w.set = true;
    ctx->committed = false;
    return RST_OK;
}

extern "C" int rst_get_weight(const rst_ctx* ctx, const char* name, float* h_data, int64_t capacity) {
    if (!ctx || !name || !h_data) return RST_ERR_INVALID;
    const Weight* w = ctx->find_weight(name);
    if (!w || !w->set) return fail(const_cast<rst_ctx*>(ctx), RST_ERR_INVALID, std::string("rst_get_weight: not set: ") + name);
    if (capacity < w->elems()) return fail(const_cast<rst_ctx*>(ctx), RST_ERR_INVALID, "rst_get_weight: buffer too small");
    std::memcpy(h_data, w->host.data(), (size_t)w->elems() * sizeof(float));
    return RST_OK;
}

static int fold_bn(rst_ctx* c, const std::string& prefix, int ch, float eps) {
    const Weight* g = c->find_weight(prefix + "/gamma");
    const Weight* b = c->find_weight(prefix + "/beta");
    const Weight* m = c->find_weight(prefix + "/moving_mean");
    const Weight* v = c->find_weight(prefix + "/moving_variance");
    std::vector<float> scale(ch), shift(ch);
    for (int i = 0; i < ch; ++i) {
        // double arithmetic then one rounding: closest to gamma*(x-mean)*rsqrt(var+eps)+beta in fp32
        double s = (double)g->host[i] / std::sqrt((double)v->host[i] + (double)eps);
        scale[i] = (float)s;
        shift[i] = (float)((double)b->host[i] - (double)m->host[i] * s);
    }
    for (auto kv : {std::make_pair("/scale", &scale), std::make_pair("/shift", &shift)}) {
        float*& d = c->folded[prefix + kv.first];
        if (!d) RST_CUDA(c, cudaMalloc(&d, ch * sizeof(float)));
        RST_CUDA(c, cudaMemcpy(d, kv.second->data(), ch * sizeof(float), cudaMemcpyHostToDevice));
    }
    return RST_OK;
}

// RST_PRECISION_FP32 inference: the residual blocks' 3x3 convolutions (62 % of the FLOPs of rst-960-120-128-17, 28 % of
// rst-960-120-32-3) run on the tensor cores as error-compensated split-tf32 GEMMs -- x = x_hi + x_lo, w = w_hi + w_lo on the tf32
// grid, y = x_hi w_hi + x_lo w_hi + x_hi w_lo accumulated in fp32: the arithmetic of RST_PRECISION_TF32X3 in rst_op_conv2d, <= 7e-7
// relative per conv (tests/test_gpu_fp32.py) -- the 9x9, strided and transposed layers keep the fp32 CUDA-core kernels.
// RST_FP32_TENSOR=0 selects the CUDA-core kernels for the residual blocks as well (A/B).
static int fp32_tensor_commit(rst_ctx* c) {
    c->res_tf32.clear();
    const char* env = ab_env("RST_FP32_TENSOR");
    if (env && env[0] == '0') return RST_OK;
    int max_ci = 0;
    for (auto& L : c->residual) {
        if (L.k != 3 || L.stride != 1 || L.transposed || L.ci % 32 != 0 || (L.co % 64 != 0 && L.co != 32)) return RST_OK;
        max_ci = std::max(max_ci, L.ci);
    }
    std::string err;
    if (!umma_init(&err)) return RST_OK;                       // no tensor-map encoder in this driver: keep the CUDA-core path
    if (!c->num_sms) RST_CUDA(c, cudaDeviceGetAttribute(&c->num_sms, cudaDevAttrMultiProcessorCount, c->device));
    std::vector<std::shared_ptr<Tf32Conv3x3>> convs;
    for (auto& L : c->residual) {
        auto t = std::make_shared<Tf32Conv3x3>();
        if (!t->setup_shape(L.ci, L.co, /*relu=*/true, /*input_gradient=*/false, &err, /*split=*/true))
            return fail(c, RST_ERR_CUDA, "fp32 tensor-core path: " + err);
        RST_CUDA(c, t->repack(c->wdev(L.name + "/kernel"), c->wdev(L.name + "/bias"), nullptr));
        convs.push_back(t);
    }
    RST_CUDA(c, cudaDeviceSynchronize());
    if (!c->tf32_scratch)
        RST_CUDA(c, cudaMalloc(&c->tf32_scratch, (size_t)c->cfg.max_batch * c->bott_h * c->bott_w * max_ci * 2 * sizeof(float)));
    c->res_tf32 = std::move(convs);
    return RST_OK;
}

extern "C" int rst_commit_weights(rst_ctx* ctx) {
    if (!ctx) return RST_ERR_INVALID;
    cudaSetDevice(ctx->device);
    for (auto& w : ctx->weights) {
        if (!w.set) return fail(ctx, RST_ERR_STATE, "rst_commit_weights: variable not set: " + w.name);
        if (!w.dev) RST_CUDA(ctx, cudaMalloc(&w.dev, (size_t)w.elems() * sizeof(float)));
        RST_CUDA(ctx, cudaMemcpy(w.dev, w.host.data(), (size_t)w.elems() * sizeof(float), cudaMemcpyHostToDevice));
    }
    for (auto& L : ctx->contract) {
        int rc = fold_bn(ctx, L.name + "/bn", L.co, 1e-3f);
        if (rc) return rc;
    }
    if (ctx->cfg.extractor == RST_EXTRACTOR_MOBILE_NET) {
        int rc = fold_bn(ctx, "mobilenet/Conv/BatchNorm", 16, 1e-3f);
        for (auto& m : ctx->mb_blocks) {
            if (!rc && m.block_id) rc = fold_bn(ctx, m.prefix + "/expand/BatchNorm", m.cexp, 1e-3f);
            if (!rc) rc = fold_bn(ctx, m.prefix + "/depthwise/BatchNorm", m.cexp, 1e-3f);
            if (!rc) rc = fold_bn(ctx, m.prefix + "/project/BatchNorm", m.cout, 1e-3f);
        }
        if (!rc) rc = fold_bn(ctx, "mobilenet/Conv_1/BatchNorm", ctx->mb_last, 1e-3f);
        if (rc) return rc;
    }
    if (ctx->cfg.precision == RST_PRECISION_BF16 && ctx->cfg.in_h != 0) {
        int rc = bf16_commit(ctx);
        if (rc) return rc;
    }
    if (ctx->cfg.precision == RST_PRECISION_FP32 && ctx->cfg.in_h != 0 && !ctx->weight_arena) {
        int rc = fp32_tensor_commit(ctx);
        if (rc) return rc;
    }
    for (auto& g : ctx->graphs) if (g.exec) cudaGraphExecDestroy(g.exec);     // captured graphs hold the old operands
    ctx->graphs.clear();
    ctx->committed = true;
    return RST_OK;
}

// ------------------------------------------------------------------------------------------------
// style-weight pyramid (styleTransfer.py:297-303, :335-345): returns the level with the given width
// ------------------------------------------------------------------------------------------------
namespace rst {
int build_mips(rst_ctx* c, const float* d_style_weights, int batch, cudaStream_t s) {
    c->mips.clear();
    if (c->cfg.num_styles < 2) return RST_OK;
    if (!d_style_weights) return fail(c, RST_ERR_INVALID, "style_weights required when num_styles > 1");
    int h = c->cfg.out_h, w = c->cfg.out_w;
    float* cur = c->w_pyramid;
    int first = 0;
    if (c->cfg.num_styles == 2 && h % 4 == 0 && w % 4 == 0 && c->n_expand + 1 >= 2 &&
        (reinterpret_cast<uintptr_t>(d_style_weights) & 15) == 0) {
        // the three levels the network reads (widths W, W/2, W/4) in one pass
        float* l1 = cur + (size_t)batch * h * w * 2;
        float* l2 = l1 + (size_t)batch * (h / 2) * (w / 2) * 2;
        {
            LaunchScope ls(c, s, "weights_pyramid");
            RST_CUDA(c, launch_weights_pyramid3(d_style_weights, cur, l1, l2, batch, h, w, s));
        }
        c->mips.emplace_back(w, cur);
        c->mips.emplace_back(w / 2, l1);
        c->mips.emplace_back(w / 4, l2);
        cur = l2; h /= 4; w /= 4; first = 2;
    } else {
        {
            LaunchScope ls(c, s, "weights_pyramid");
            RST_CUDA(c, launch_weights_concat(d_style_weights, cur, (long long)batch * h * w, c->cfg.num_styles - 1, s));
        }
        c->mips.emplace_back(w, cur);
    }
    for (int i = first; i < c->n_expand + 1; ++i) {
        if (h < 2 || w < 2) break;
        float* nxt = cur + (size_t)batch * h * w * 2;
        LaunchScope ls(c, s, "weights_pyramid");
        RST_CUDA(c, launch_avgpool2_f32(cur, nxt, batch, h, w, 2, s));
        h /= 2; w /= 2;
        c->mips.emplace_back(w, nxt);
        cur = nxt;
    }
    return RST_OK;
}
const float* mip_for_width(rst_ctx* c, int width) {
    for (auto& m : c->mips) if (m.first == width) return m.second;
    return nullptr;
}
}  // namespace rst

// ------------------------------------------------------------------------------------------------
// fp32 transfer forward (styleTransfer.py:305-329)
// ------------------------------------------------------------------------------------------------
static ConvF32 conv_params(rst_ctx* c, const LayerDesc& L, const float* x, float* y, int batch,
                           const std::string& kname, const std::string& bname) {
    ConvF32 p;
    p.x = x; p.y = y; p.w = c->wdev(kname); p.bias = c->wdev(bname);
    p.B = batch; p.Hi = L.hi; p.Wi = L.wi; p.Ci = L.ci; p.Ho = L.ho; p.Wo = L.wo; p.Co = L.co;
    p.kh = L.k; p.kw = L.k; p.stride = L.stride; p.pad_t = L.pad_t; p.pad_l = L.pad_l;
    p.transposed = L.transposed ? 1 : 0;
    if (!L.transposed) { p.w_tap = (long long)L.ci * L.co; p.w_ci = L.co; p.w_co = 1; }
    else { p.w_tap = (long long)L.ci * L.co; p.w_ci = 1; p.w_co = L.ci; }
    return p;
}

static int cin_layer(rst_ctx* c, float* x, float* y, const float* residual, int batch, int P, int C, int width,
                     const float* d_style_params, int param_off, int act, cudaStream_t s) {
    {
        LaunchScope ls(c, s, "cin_stats", 2);
        RST_CUDA(c, launch_zero_f64(c->stats, (long long)batch * C * 2, s));
        RST_CUDA(c, launch_moments_f32(x, c->stats, batch, P, C, s));
    }
    CinApply a;
    a.x = x; a.y = y; a.residual = residual; a.stats = c->stats; a.params = d_style_params;
    a.param_bstride = (long long)c->cfg.num_styles * c->num_style_params;
    a.param_sstride = c->num_style_params;
    a.scale_off = param_off; a.bias_off = param_off + C;
    a.B = batch; a.P = P; a.C = C; a.num_styles = c->cfg.num_styles; a.act = act;
    if (c->cfg.num_styles == 2) {
        a.weights = mip_for_width(c, width);
        if (!a.weights) return fail(c, RST_ERR_STATE, "no style-weight mip for this layer width");
    }
    LaunchScope ls(c, s, "cin_apply");
    RST_CUDA(c, launch_cin_apply(a, s));
    return RST_OK;
}

namespace rst {

// contract stage (styleTransfer.py:188-205, :224-232): returns the fp32 bottleneck input and the index of the
// ping-pong buffer that is free afterwards.
int fp32_contract_stage(rst_ctx* c, const float* d_content, int batch, cudaStream_t s, float** out, int* free_idx) {
    const float* cur = d_content;
    int which = 0;
    for (auto& L : c->contract) {
        float* y = c->act[which];
        ConvF32 p = conv_params(c, L, cur, y, batch, L.name + "/conv/kernel", L.name + "/conv/bias");
        p.act1 = ACT_RELU;
        p.post_scale = c->folded[L.name + "/bn/scale"];
        p.post_shift = c->folded[L.name + "/bn/shift"];
        p.act2 = ACT_RELU;
        {
            LaunchScope ls(c, s, "conv_fp32");
            RST_CUDA(c, launch_conv_f32(p, s));
        }
        record_tap(c, L.name, y, (int64_t)batch * L.ho * L.wo * L.co, false, s);
        cur = y;
        which ^= 1;
    }
    *out = const_cast<float*>(cur);
    *free_idx = which;
    return RST_OK;
}

// expand stage (styleTransfer.py:95-141, :260-276): x is consumed, t1 is scratch, result lands in d_out.
int fp32_expand_stage(rst_ctx* c, float* x, float* t1, const float* d_style_params, int cursor, float* d_out, int batch,
                      cudaStream_t s) {
    for (size_t i = 0; i < c->expand.size(); ++i) {
        const LayerDesc& L = c->expand[i];
        const bool last = i + 1 == c->expand.size();
        ConvF32 p = conv_params(c, L, x, t1, batch, L.name + "/conv/kernel", L.name + "/conv/bias");
        { LaunchScope ls(c, s, "conv_fp32"); RST_CUDA(c, launch_conv_f32(p, s)); }
        record_tap(c, L.name + "/conv", t1, (int64_t)batch * L.ho * L.wo * L.co, false, s);
        float* y = last ? d_out : t1;
        int rc = cin_layer(c, t1, y, nullptr, batch, L.ho * L.wo, L.co, L.wo, d_style_params, cursor,
                           last ? ACT_SIGMOID : ACT_RELU, s);
        if (rc) return rc;
        record_tap(c, L.name, y, (int64_t)batch * L.ho * L.wo * L.co, false, s);
        cursor += 2 * L.co;
        std::swap(x, t1);
    }
    return RST_OK;
}

}  // namespace rst

static int fp32_transfer_forward(rst_ctx* c, const float* d_content, const float* d_style_params,
                                 const float* d_style_weights, float* d_out, int batch, cudaStream_t s) {
    int rc = build_mips(c, d_style_weights, batch, s);
    if (rc) return rc;
    float* x = nullptr;
    int which = 0;
    rc = fp32_contract_stage(c, d_content, batch, s, &x, &which);
    if (rc) return rc;
    float* t1 = c->act[which];
    float* t2 = c->act[2];
    const int F = c->cfg.bottleneck_num_filters;
    const int P = c->bott_h * c->bott_w;
    int cursor = 0;
    const bool tensor = c->res_tf32.size() == c->residual.size() && !c->res_tf32.empty();
    for (int b = 0; b < 5; ++b) {                                          // residual_block, :144-185
        const LayerDesc& L0 = c->residual[2 * b];
        const LayerDesc& L1 = c->residual[2 * b + 1];
        if (tensor) {
            LaunchScope ls(c, s, "conv_tf32x3", 2);
            RST_CUDA(c, c->res_tf32[2 * b]->run_split(x, c->tf32_scratch, t1, batch, c->bott_h, c->bott_w, c->num_sms, s, &c->err));
        } else {
            ConvF32 p0 = conv_params(c, L0, x, t1, batch, L0.name + "/kernel", L0.name + "/bias");
            p0.act1 = ACT_RELU;
            LaunchScope ls(c, s, "conv_fp32");
            RST_CUDA(c, launch_conv_f32(p0, s));
        }
        record_tap(c, L0.name + "/relu", t1, (int64_t)batch * P * F, false, s);
        rc = cin_layer(c, t1, t1, nullptr, batch, P, F, c->bott_w, d_style_params, cursor, ACT_RELU, s);
        if (rc) return rc;
        record_tap(c, L0.name + "/cin", t1, (int64_t)batch * P * F, false, s);
        if (tensor) {
            LaunchScope ls(c, s, "conv_tf32x3", 2);
            RST_CUDA(c, c->res_tf32[2 * b + 1]->run_split(t1, c->tf32_scratch, t2, batch, c->bott_h, c->bott_w, c->num_sms, s, &c->err));
        } else {
            ConvF32 p1 = conv_params(c, L1, t1, t2, batch, L1.name + "/kernel", L1.name + "/bias");
            p1.act1 = ACT_RELU;
            LaunchScope ls(c, s, "conv_fp32");
            RST_CUDA(c, launch_conv_f32(p1, s));
        }
        record_tap(c, L1.name + "/relu", t2, (int64_t)batch * P * F, false, s);
        rc = cin_layer(c, t2, t2, b == 0 ? nullptr : x, batch, P, F, c->bott_w, d_style_params, cursor + 2 * F,
                       ACT_NONE, s);
        if (rc) return rc;
        record_tap(c, "residual_block_" + std::to_string(b), t2, (int64_t)batch * P * F, false, s);
        cursor += 4 * F;
        std::swap(x, t2);
    }
    return fp32_expand_stage(c, x, t1, d_style_params, cursor, d_out, batch, s);
}

static int collect_profile(rst_ctx* c) {
    for (auto& t : c->pending_events) {
        cudaEventSynchronize(std::get<2>(t));
        float ms = 0.f;
        cudaEventElapsedTime(&ms, std::get<1>(t), std::get<2>(t));
        ProfileGroup& g = c->profile[std::get<0>(t)];
        g.total_ms += ms;
        g.launches += 1;
        c->event_pool.push_back(std::get<1>(t));
        c->event_pool.push_back(std::get<2>(t));
    }
    c->pending_events.clear();
    return RST_OK;
}

static size_t dtype_size(int dtype) { return dtype == RST_DTYPE_F32 ? 4 : dtype == RST_DTYPE_F16 ? 2 : 1; }

__global__ void f16_to_f32_kernel(const __half2* __restrict__ x, float2* __restrict__ y, long long n2) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n2) y[i] = __half22float2(x[i]);
}
__global__ void f32_to_u8_kernel(const float4* __restrict__ x, uint32_t* __restrict__ y, long long n4) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    const float4 v = x[i];
    auto q = [](float f) { return min(__float2uint_rz(fmaxf(f, 0.f) * 255.f), 255u); };
    y[i] = q(v.x) | (q(v.y) << 8) | (q(v.z) << 16) | (q(v.w) << 24);
}

extern "C" int rst_transfer_forward(rst_ctx* ctx, const float* d_content, const float* d_style_params,
                                    const float* d_style_weights, float* d_out, int batch, void* stream) {
    return rst_transfer_forward_typed(ctx, d_content, RST_DTYPE_F32, d_style_params, d_style_weights, d_out, RST_DTYPE_F32, batch,
                                      stream);
}

extern "C" int rst_transfer_forward_typed(rst_ctx* ctx, const void* d_content, int content_dtype, const float* d_style_params,
                                          const float* d_style_weights, void* d_out, int out_dtype, int batch, void* stream) {
    if (!ctx) return RST_ERR_INVALID;
    if (content_dtype != RST_DTYPE_F32 && content_dtype != RST_DTYPE_F16)
        return fail(ctx, RST_ERR_INVALID, "rst_transfer_forward: content dtype must be RST_DTYPE_F32 or RST_DTYPE_F16");
    if (out_dtype != RST_DTYPE_F32 && out_dtype != RST_DTYPE_U8)
        return fail(ctx, RST_ERR_INVALID, "rst_transfer_forward: output dtype must be RST_DTYPE_F32 or RST_DTYPE_U8");
    if (ctx->cfg.in_h == 0) return fail(ctx, RST_ERR_STATE, "rst_transfer_forward: predictor-only context");
    if (!ctx->committed) return fail(ctx, RST_ERR_STATE, "rst_transfer_forward: weights not committed");
    if (!d_content || !d_style_params || !d_out) return fail(ctx, RST_ERR_INVALID, "rst_transfer_forward: null tensor");
    if (batch < 0 || batch > ctx->cfg.max_batch) return fail(ctx, RST_ERR_INVALID, "rst_transfer_forward: batch exceeds max_batch");
    if (batch == 0) return RST_OK;
    cudaSetDevice(ctx->device);
    ctx->launches = 0;
    cudaStream_t s = (cudaStream_t)stream;
    const rst_config& cg = ctx->cfg;
    const long long n_in = (long long)batch * cg.in_h * cg.in_w * cg.in_c, n_out = (long long)batch * cg.out_h * cg.out_w * 3;
    if (cg.precision != RST_PRECISION_BF16 && (content_dtype != RST_DTYPE_F32 || out_dtype != RST_DTYPE_F32)) {
        // the fp32 path computes on fp32 tensors: typed calls convert through two buffers allocated on first use
        if ((n_in & 1) || (n_out & 3)) return fail(ctx, RST_ERR_UNSUPPORTED, "typed fp32-path call: odd tensor sizes");
        if (content_dtype == RST_DTYPE_F16 && !ctx->cvt_content)
            RST_CUDA(ctx, cudaMalloc(&ctx->cvt_content, (size_t)cg.max_batch * cg.in_h * cg.in_w * cg.in_c * sizeof(float)));
        if (out_dtype == RST_DTYPE_U8 && !ctx->cvt_out)
            RST_CUDA(ctx, cudaMalloc(&ctx->cvt_out, (size_t)cg.max_batch * cg.out_h * cg.out_w * 3 * sizeof(float)));
    }
    auto run = [&]() -> int {
        if (cg.precision == RST_PRECISION_BF16)
            return bf16_transfer_forward(ctx, d_content, content_dtype, d_style_params, d_style_weights, d_out, out_dtype, batch, s);
        const float* x = (const float*)d_content;
        float* y = out_dtype == RST_DTYPE_U8 ? ctx->cvt_out : (float*)d_out;
        if (content_dtype == RST_DTYPE_F16) {
            LaunchScope ls(ctx, s, "convert");
            f16_to_f32_kernel<<<(unsigned)((n_in / 2 + 255) / 256), 256, 0, s>>>((const __half2*)d_content, (float2*)ctx->cvt_content, n_in / 2);
            x = ctx->cvt_content;
        }
        int rc2 = fp32_transfer_forward(ctx, x, d_style_params, d_style_weights, y, batch, s);
        if (rc2 == RST_OK && out_dtype == RST_DTYPE_U8) {
            LaunchScope ls(ctx, s, "convert");
            f32_to_u8_kernel<<<(unsigned)((n_out / 4 + 255) / 256), 256, 0, s>>>((const float4*)y, (uint32_t*)d_out, n_out / 4);
            RST_CUDA(ctx, cudaGetLastError());
        }
        return rc2;
    };
    // Replay a captured CUDA graph when the same buffers come back (video loop / pipelined host API); the legacy default
    // stream cannot be captured, profiling and taps need the eager path.
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    const bool graphable = ctx->use_graphs && s != nullptr && s != cudaStreamLegacy && s != cudaStreamPerThread &&
                           !ctx->profiling && !ctx->keep_taps &&
                           cudaStreamIsCapturing(s, &cap) == cudaSuccess && cap == cudaStreamCaptureStatusNone;
    if (!graphable) return run();
    for (auto& g : ctx->graphs)
        if (g.batch == batch && g.content == d_content && g.params == d_style_params && g.weights == d_style_weights &&
            g.out == d_out && g.content_dtype == content_dtype && g.out_dtype == out_dtype) {
            RST_CUDA(ctx, cudaGraphLaunch(g.exec, s));
            ctx->launches = g.launches;
            return RST_OK;
        }
    // capture only when this exact set of buffers comes back: callers that pass fresh device tensors on every call would
    // otherwise pay a capture + instantiate per call
    {
        bool seen = false;
        for (auto& g : ctx->graph_candidates)
            seen = seen || (g.batch == batch && g.content == d_content && g.params == d_style_params && g.weights == d_style_weights &&
                            g.out == d_out && g.content_dtype == content_dtype && g.out_dtype == out_dtype);
        if (!seen) {
            rst_ctx::GraphEntry k;
            k.batch = batch; k.content = d_content; k.params = d_style_params; k.weights = d_style_weights; k.out = d_out;
            k.content_dtype = content_dtype; k.out_dtype = out_dtype;
            if (ctx->graph_candidates.size() >= 16) ctx->graph_candidates.erase(ctx->graph_candidates.begin());
            ctx->graph_candidates.push_back(k);
            return run();
        }
    }
    if (cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
        cudaGetLastError();
        return run();
    }
    int rc = run();
    cudaGraph_t graph = nullptr;
    cudaError_t e = cudaStreamEndCapture(s, &graph);
    if (rc != RST_OK || e != cudaSuccess || !graph) {
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        ctx->use_graphs = false;                 // do not retry capture on this context
        return rc != RST_OK ? rc : run();
    }
    rst_ctx::GraphEntry ge;
    ge.batch = batch; ge.content = d_content; ge.params = d_style_params; ge.weights = d_style_weights; ge.out = d_out;
    ge.content_dtype = content_dtype; ge.out_dtype = out_dtype;
    ge.launches = ctx->launches;
    e = cudaGraphInstantiate(&ge.exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) { cudaGetLastError(); ctx->use_graphs = false; return run(); }
    if (ctx->graphs.size() >= 8) { cudaGraphExecDestroy(ctx->graphs.front().exec); ctx->graphs.erase(ctx->graphs.begin()); }
    ctx->graphs.push_back(ge);
    RST_CUDA(ctx, cudaGraphLaunch(ge.exec, s));
    return RST_OK;
}

extern "C" int rst_transfer_forward_host(rst_ctx* ctx, const float* h_content, const float* h_style_params,
                                         const float* h_style_weights, float* h_out, int batch) {
    return rst_transfer_forward_host_typed(ctx, h_content, RST_DTYPE_F32, h_style_params, h_style_weights, h_out, RST_DTYPE_F32, batch);
}

extern "C" int rst_transfer_forward_host_typed(rst_ctx* ctx, const void* h_content, int content_dtype, const float* h_style_params,
                                               const float* h_style_weights, void* h_out, int out_dtype, int batch) {
    if (!ctx) return RST_ERR_INVALID;
    if ((content_dtype != RST_DTYPE_F32 && content_dtype != RST_DTYPE_F16) || (out_dtype != RST_DTYPE_F32 && out_dtype != RST_DTYPE_U8))
        return fail(ctx, RST_ERR_INVALID, "rst_transfer_forward_host: content dtype F32|F16, output dtype F32|U8");
    if (!h_content || !h_style_params || !h_out) return fail(ctx, RST_ERR_INVALID, "rst_transfer_forward_host: null tensor");
    if (batch < 0 || batch > ctx->cfg.max_batch) return fail(ctx, RST_ERR_INVALID, "rst_transfer_forward_host: batch exceeds max_batch");
    if (batch == 0) return RST_OK;
    cudaSetDevice(ctx->device);
    const rst_config& g = ctx->cfg;
    cudaStream_t s = ctx->own_stream;
    RST_CUDA(ctx, cudaMemcpyAsync(ctx->st_content, h_content, (size_t)batch * g.in_h * g.in_w * g.in_c * dtype_size(content_dtype),
                                  cudaMemcpyHostToDevice, s));
    RST_CUDA(ctx, cudaMemcpyAsync(ctx->st_params, h_style_params,
                                  (size_t)batch * g.num_styles * ctx->num_style_params * sizeof(float),
                                  cudaMemcpyHostToDevice, s));
    if (g.num_styles > 1) {
        if (!h_style_weights) return fail(ctx, RST_ERR_INVALID, "style_weights required when num_styles > 1");
        RST_CUDA(ctx, cudaMemcpyAsync(ctx->st_weights, h_style_weights,
                                      (size_t)batch * g.out_h * g.out_w * (g.num_styles - 1) * sizeof(float),
                                      cudaMemcpyHostToDevice, s));
    }
    int rc = rst_transfer_forward_typed(ctx, ctx->st_content, content_dtype, ctx->st_params,
                                        g.num_styles > 1 ? ctx->st_weights : nullptr, ctx->st_out, out_dtype, batch, (void*)s);
    if (rc) return rc;
    RST_CUDA(ctx, cudaMemcpyAsync(h_out, ctx->st_out, (size_t)batch * g.out_h * g.out_w * 3 * dtype_size(out_dtype),
                                  cudaMemcpyDeviceToHost, s));
    RST_CUDA(ctx, cudaStreamSynchronize(s));
    if (ctx->profiling) collect_profile(ctx);
    return RST_OK;
}

// ------------------------------------------------------------------------------------------------
// Asynchronous, double-buffered host pipeline: the frame loop of predict_video_using_checkpoint.py:90-98
// (`for frame in dataset.prefetch(5): transfer.predict(...)`) with the H2D copy of batch i+1 and the D2H copy of
// batch i-1 overlapping the forward of batch i.  Host buffers should be pinned for the copies to be asynchronous.
// ------------------------------------------------------------------------------------------------
static int pipe_init(rst_ctx* c) {
    if (c->pipe.ready) return RST_OK;
    const rst_config& g = c->cfg;
    const size_t B = g.max_batch;
    for (int i = 0; i < 2; ++i) {
        RST_CUDA(c, cudaMalloc(&c->pipe.content[i], B * g.in_h * g.in_w * g.in_c * sizeof(float)));
        RST_CUDA(c, cudaMalloc(&c->pipe.params[i], B * g.num_styles * c->num_style_params * sizeof(float)));
        RST_CUDA(c, cudaMalloc(&c->pipe.out[i], B * g.out_h * g.out_w * 3 * sizeof(float)));
        if (g.num_styles > 1)
            RST_CUDA(c, cudaMalloc(&c->pipe.weights[i], B * g.out_h * g.out_w * (g.num_styles - 1) * sizeof(float)));
        RST_CUDA(c, cudaEventCreateWithFlags(&c->pipe.in_done[i], cudaEventDisableTiming));
        RST_CUDA(c, cudaEventCreateWithFlags(&c->pipe.comp_done[i], cudaEventDisableTiming));
        RST_CUDA(c, cudaEventCreateWithFlags(&c->pipe.out_done[i], cudaEventDisableTiming));
    }
    RST_CUDA(c, cudaStreamCreateWithFlags(&c->pipe.s_in, cudaStreamNonBlocking));
    RST_CUDA(c, cudaStreamCreateWithFlags(&c->pipe.s_out, cudaStreamNonBlocking));
    c->pipe.ready = true;
    return RST_OK;
}

extern "C" int rst_transfer_submit_host(rst_ctx* ctx, const float* h_content, const float* h_style_params,
                                        const float* h_style_weights, float* h_out, int batch, int64_t* ticket) {
    return rst_transfer_submit_host_typed(ctx, h_content, RST_DTYPE_F32, h_style_params, h_style_weights, h_out, RST_DTYPE_F32, batch,
                                          ticket);
}

extern "C" int rst_transfer_submit_host_typed(rst_ctx* ctx, const void* h_content, int content_dtype, const float* h_style_params,
                                              const float* h_style_weights, void* h_out, int out_dtype, int batch, int64_t* ticket) {
    if (!ctx) return RST_ERR_INVALID;
    if ((content_dtype != RST_DTYPE_F32 && content_dtype != RST_DTYPE_F16) || (out_dtype != RST_DTYPE_F32 && out_dtype != RST_DTYPE_U8))
        return fail(ctx, RST_ERR_INVALID, "rst_transfer_submit_host: content dtype F32|F16, output dtype F32|U8");
    if (ctx->cfg.in_h == 0) return fail(ctx, RST_ERR_STATE, "rst_transfer_submit_host: predictor-only context");
    if (!ctx->committed) return fail(ctx, RST_ERR_STATE, "rst_transfer_submit_host: weights not committed");
    if (!h_content || !h_style_params || !h_out || !ticket) return fail(ctx, RST_ERR_INVALID, "rst_transfer_submit_host: null argument");
    if (batch < 1 || batch > ctx->cfg.max_batch) return fail(ctx, RST_ERR_INVALID, "rst_transfer_submit_host: batch out of range");
    cudaSetDevice(ctx->device);
    int rc = pipe_init(ctx);
    if (rc) return rc;
    const rst_config& g = ctx->cfg;
    auto& P = ctx->pipe;
    const int slot = (int)(P.next & 1);
    if (P.busy[slot]) RST_CUDA(ctx, cudaEventSynchronize(P.out_done[slot]));     // the slot's previous batch has fully drained
    RST_CUDA(ctx, cudaMemcpyAsync(P.content[slot], h_content, (size_t)batch * g.in_h * g.in_w * g.in_c * dtype_size(content_dtype),
                                  cudaMemcpyHostToDevice, P.s_in));
    RST_CUDA(ctx, cudaMemcpyAsync(P.params[slot], h_style_params,
                                  (size_t)batch * g.num_styles * ctx->num_style_params * sizeof(float), cudaMemcpyHostToDevice, P.s_in));
    if (g.num_styles > 1) {
        if (!h_style_weights) return fail(ctx, RST_ERR_INVALID, "style_weights required when num_styles > 1");
        RST_CUDA(ctx, cudaMemcpyAsync(P.weights[slot], h_style_weights,
                                      (size_t)batch * g.out_h * g.out_w * (g.num_styles - 1) * sizeof(float),
                                      cudaMemcpyHostToDevice, P.s_in));
    }
    RST_CUDA(ctx, cudaEventRecord(P.in_done[slot], P.s_in));
    RST_CUDA(ctx, cudaStreamWaitEvent(ctx->own_stream, P.in_done[slot], 0));
    rc = rst_transfer_forward_typed(ctx, P.content[slot], content_dtype, P.params[slot],
                                    g.num_styles > 1 ? P.weights[slot] : nullptr, P.out[slot], out_dtype, batch, (void*)ctx->own_stream);
    if (rc) return rc;
    RST_CUDA(ctx, cudaEventRecord(P.comp_done[slot], ctx->own_stream));
    RST_CUDA(ctx, cudaStreamWaitEvent(P.s_out, P.comp_done[slot], 0));
    RST_CUDA(ctx, cudaMemcpyAsync(h_out, P.out[slot], (size_t)batch * g.out_h * g.out_w * 3 * dtype_size(out_dtype),
                                  cudaMemcpyDeviceToHost, P.s_out));
    RST_CUDA(ctx, cudaEventRecord(P.out_done[slot], P.s_out));
    P.busy[slot] = true;
    *ticket = P.next++;
    return RST_OK;
}

extern "C" int rst_transfer_wait(rst_ctx* ctx, int64_t ticket) {
    if (!ctx || !ctx->pipe.ready) return RST_ERR_INVALID;
    if (ticket < 0 || ticket >= ctx->pipe.next) return fail(ctx, RST_ERR_INVALID, "rst_transfer_wait: unknown ticket");
    if (ticket + 2 < ctx->pipe.next) return RST_OK;          // slot already recycled => that batch completed long ago
    cudaSetDevice(ctx->device);
    RST_CUDA(ctx, cudaEventSynchronize(ctx->pipe.out_done[ticket & 1]));
    if (ctx->profiling) collect_profile(ctx);
    return RST_OK;
}

// ------------------------------------------------------------------------------------------------
// style predictor (stylePrediction.py:25-75), fp32 CUDA-core kernels
// ------------------------------------------------------------------------------------------------
static int conv1x1(rst_ctx* c, const float* x, float* y, int batch, int h, int w, int ci, int co, const float* kernel,
                   const float* bias, const float* ps, const float* pshift, int act1, int act2, const float* residual,
                   cudaStream_t s) {
    ConvF32 p;
    p.x = x; p.y = y; p.w = kernel; p.bias = bias;
    p.B = batch; p.Hi = h; p.Wi = w; p.Ci = ci; p.Ho = h; p.Wo = w; p.Co = co;
    p.w_tap = 0; p.w_ci = co; p.w_co = 1;
    p.act1 = act1; p.post_scale = ps; p.post_shift = pshift; p.act2 = act2; p.residual = residual;
    LaunchScope ls(c, s, "predictor");
    RST_CUDA(c, launch_conv_f32(p, s));
    return RST_OK;
}

static int global_mean(rst_ctx* c, const float* x, float* mean, int batch, int P, int C, cudaStream_t s) {
    LaunchScope ls(c, s, "predictor", 3);
    RST_CUDA(c, launch_zero_f64(c->stats, (long long)batch * C * 2, s));
    RST_CUDA(c, launch_moments_f32(x, c->stats, batch, P, C, s));
    RST_CUDA(c, launch_stats_to_mean(c->stats, mean, batch, P, C, s));
    return RST_OK;
}

static int predictor_forward(rst_ctx* c, const float* d_style, float* d_params, int batch, cudaStream_t s) {
    const rst_config& g = c->cfg;
    if (g.extractor == RST_EXTRACTOR_NONE) return fail(c, RST_ERR_STATE, "context was created without a style predictor");
    int h = g.style_h, w = g.style_w, rc = 0;
    float* feat = nullptr;
    int fc = 0;
    if (g.extractor == RST_EXTRACTOR_DUMMY) {                       // Conv2D(1, 9, 5, 'same'), stylePrediction.py:30-31
        ConvF32 p;
        p.x = d_style; p.y = c->pact[0]; p.w = c->wdev("dummy_conv/kernel"); p.bias = c->wdev("dummy_conv/bias");
        p.B = batch; p.Hi = h; p.Wi = w; p.Ci = 3; p.Co = 1; p.kh = 9; p.kw = 9; p.stride = 5;
        p.Ho = tf_same(h, 9, 5, &p.pad_t); p.Wo = tf_same(w, 9, 5, &p.pad_l);
        p.w_tap = 3; p.w_ci = 1; p.w_co = 1;
        { LaunchScope ls(c, s, "predictor"); RST_CUDA(c, launch_conv_f32(p, s)); }
        feat = c->pact[0]; fc = 1; h = p.Ho; w = p.Wo;
    } else {
        // stem: Rescaling(2,-1) -> Conv 3x3 s2 'same' (no bias) -> BN -> hard-swish
        ConvF32 p;
        p.x = d_style; p.y = c->pact[0]; p.w = c->wdev("mobilenet/Conv/kernel");
        p.B = batch; p.Hi = h; p.Wi = w; p.Ci = 3; p.Co = 16; p.kh = 3; p.kw = 3; p.stride = 2;
        p.Ho = tf_same(h, 3, 2, &p.pad_t); p.Wo = tf_same(w, 3, 2, &p.pad_l);
        p.w_tap = 3 * 16; p.w_ci = 16; p.w_co = 1;
        p.in_scale = 2.f; p.in_shift = -1.f;
        p.post_scale = c->folded["mobilenet/Conv/BatchNorm/scale"];
        p.post_shift = c->folded["mobilenet/Conv/BatchNorm/shift"];
        p.act2 = ACT_HSWISH;
        { LaunchScope ls(c, s, "predictor"); RST_CUDA(c, launch_conv_f32(p, s)); }
        h = p.Ho; w = p.Wo;
        float* x = c->pact[0];
        float* t1 = c->pact[1];
        float* t2 = c->pact[2];
        float* t3 = c->pact[3];
        for (auto& m : c->mb_blocks) {
            const float* in = x;
            const float* e = x;
            if (m.block_id) {   // expand 1x1 -> BN -> act
                rc = conv1x1(c, x, t1, batch, h, w, m.cin, m.cexp, c->wdev(m.prefix + "/expand/kernel"), nullptr,
                             c->folded[m.prefix + "/expand/BatchNorm/scale"], c->folded[m.prefix + "/expand/BatchNorm/shift"],
                             ACT_NONE, m.act, nullptr, s);
                if (rc) return rc;
                e = t1;
            }
            DepthwiseF32 d;
            d.x = e; d.y = t2; d.w = c->wdev(m.prefix + "/depthwise/depthwise_kernel");
            d.B = batch; d.Hi = h; d.Wi = w; d.C = m.cexp; d.k = m.k; d.stride = m.s;
            if (m.s == 2) {     // ZeroPadding2D(correct_pad) + 'valid'
                int pt = m.k / 2 - (1 - h % 2), pl = m.k / 2 - (1 - w % 2);
                d.pad_t = pt; d.pad_l = pl;
                d.Ho = (h + pt + m.k / 2 - m.k) / 2 + 1;
                d.Wo = (w + pl + m.k / 2 - m.k) / 2 + 1;
            } else {
                d.Ho = tf_same(h, m.k, 1, &d.pad_t);
                d.Wo = tf_same(w, m.k, 1, &d.pad_l);
            }
            d.post_scale = c->folded[m.prefix + "/depthwise/BatchNorm/scale"];
            d.post_shift = c->folded[m.prefix + "/depthwise/BatchNorm/shift"];
            d.act = m.act;
            { LaunchScope ls(c, s, "predictor"); RST_CUDA(c, launch_depthwise_f32(d, s)); }
            int ho = d.Ho, wo = d.Wo;
            float* dw = t2;
            if (m.se) {         // squeeze-excite: GAP -> 1x1+b -> relu -> 1x1+b -> hard-sigmoid -> multiply
                rc = global_mean(c, dw, c->pvec[0], batch, ho * wo, m.cexp, s);
                if (!rc) rc = conv1x1(c, c->pvec[0], c->pvec[1], batch, 1, 1, m.cexp, m.se,
                                      c->wdev(m.prefix + "/squeeze_excite/Conv/kernel"),
                                      c->wdev(m.prefix + "/squeeze_excite/Conv/bias"), nullptr, nullptr, ACT_RELU, ACT_NONE,
                                      nullptr, s);
                if (!rc) rc = conv1x1(c, c->pvec[1], c->pvec[2], batch, 1, 1, m.se, m.cexp,
                                      c->wdev(m.prefix + "/squeeze_excite/Conv_1/kernel"),
                                      c->wdev(m.prefix + "/squeeze_excite/Conv_1/bias"), nullptr, nullptr, ACT_HSIGMOID,
                                      ACT_NONE, nullptr, s);
                if (rc) return rc;
                LaunchScope ls(c, s, "predictor");
                RST_CUDA(c, launch_scale_channels(dw, c->pvec[2], dw, batch, ho * wo, m.cexp, s));
            }
            const bool add = m.s == 1 && m.cin == m.cout;
            rc = conv1x1(c, dw, t3, batch, ho, wo, m.cexp, m.cout, c->wdev(m.prefix + "/project/kernel"), nullptr,
                         c->folded[m.prefix + "/project/BatchNorm/scale"], c->folded[m.prefix + "/project/BatchNorm/shift"],
                         ACT_NONE, ACT_NONE, add ? in : nullptr, s);
            if (rc) return rc;
            std::swap(x, t3);
            h = ho; w = wo;
        }
        rc = conv1x1(c, x, t1, batch, h, w, c->mb_blocks.back().cout, c->mb_last, c->wdev("mobilenet/Conv_1/kernel"), nullptr,
                     c->folded["mobilenet/Conv_1/BatchNorm/scale"], c->folded["mobilenet/Conv_1/BatchNorm/shift"],
                     ACT_NONE, ACT_HSWISH, nullptr, s);
        if (rc) return rc;
        feat = t1; fc = c->mb_last;
    }
    record_tap(c, "predictor/features", feat, (int64_t)batch * h * w * fc, false, s);
    rc = global_mean(c, feat, c->pvec[0], batch, h * w, fc, s);                                    // avg_pool, :54
    if (!rc) rc = conv1x1(c, c->pvec[0], c->pvec[1], batch, 1, 1, fc, 100, c->wdev("StylePredictor/kernel"),
                          c->wdev("StylePredictor/bias"), nullptr, nullptr, ACT_NONE, ACT_NONE, nullptr, s);   // :59-63
    if (!rc) rc = conv1x1(c, c->pvec[1], d_params, batch, 1, 1, 100, c->num_style_params,
                          c->wdev("StyleNormPredictor/kernel"), c->wdev("StyleNormPredictor/bias"), nullptr, nullptr,
                          ACT_NONE, ACT_NONE, nullptr, s);                                                   // :66-70
    return rc;
}

extern "C" int rst_predict_style(rst_ctx* ctx, const float* d_style, float* d_params, int batch, void* stream) {
    if (!ctx) return RST_ERR_INVALID;
    if (!ctx->committed) return fail(ctx, RST_ERR_STATE, "rst_predict_style: weights not committed");
    if (!d_style || !d_params) return fail(ctx, RST_ERR_INVALID, "rst_predict_style: null tensor");
    if (batch < 0 || batch > ctx->cfg.max_batch) return fail(ctx, RST_ERR_INVALID, "rst_predict_style: batch exceeds max_batch");
    if (batch == 0) return RST_OK;
    cudaSetDevice(ctx->device);
    ctx->launches = 0;
    return predictor_forward(ctx, d_style, d_params, batch, (cudaStream_t)stream);
}

extern "C" int rst_predict_style_host(rst_ctx* ctx, const float* h_style, float* h_params, int batch) {
    if (!ctx) return RST_ERR_INVALID;
    if (!h_style || !h_params) return fail(ctx, RST_ERR_INVALID, "rst_predict_style_host: null tensor");
    if (batch < 0 || batch > ctx->cfg.max_batch) return fail(ctx, RST_ERR_INVALID, "rst_predict_style_host: batch exceeds max_batch");
    if (batch == 0) return RST_OK;
    if (ctx->cfg.extractor == RST_EXTRACTOR_NONE) return fail(ctx, RST_ERR_STATE, "context was created without a style predictor");
    cudaSetDevice(ctx->device);
    const rst_config& g = ctx->cfg;
    cudaStream_t s = ctx->own_stream;
    RST_CUDA(ctx, cudaMemcpyAsync(ctx->st_style, h_style, (size_t)batch * g.style_h * g.style_w * 3 * sizeof(float),
                                  cudaMemcpyHostToDevice, s));
    int rc = rst_predict_style(ctx, ctx->st_style, ctx->st_params, batch, (void*)s);
    if (rc) return rc;
    RST_CUDA(ctx, cudaMemcpyAsync(h_params, ctx->st_params, (size_t)batch * ctx->num_style_params * sizeof(float),
                                  cudaMemcpyDeviceToHost, s));
    RST_CUDA(ctx, cudaStreamSynchronize(s));
    return RST_OK;
}

// styleTransferInferenceModel.py:24-39: unstack styles, predictor per style, stack -> (B,S,P), transfer.
__global__ void interleave_params_kernel(const float* __restrict__ src, float* __restrict__ dst, int B, int S, int P,
                                         int s_idx) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)B * P) return;
    long long b = i / P;
    int j = (int)(i % P);
    dst[(b * S + s_idx) * P + j] = src[i];
}
__global__ void gather_style_kernel(const float* __restrict__ src, float* __restrict__ dst, long long img, int S, int s_idx,
                                    long long total) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    long long b = i / img;
    long long r = i % img;
    dst[i] = src[(b * S + s_idx) * img + r];
}

extern "C" int rst_inference_forward_host(rst_ctx* ctx, const float* h_content, const float* h_style,
                                          const float* h_style_weights, float* h_out, int batch) {
    if (!ctx) return RST_ERR_INVALID;
    if (!ctx->committed) return fail(ctx, RST_ERR_STATE, "rst_inference_forward_host: weights not committed");
    if (!h_content || !h_style || !h_out) return fail(ctx, RST_ERR_INVALID, "rst_inference_forward_host: null tensor");
    if (batch < 0 || batch > ctx->cfg.max_batch) return fail(ctx, RST_ERR_INVALID, "batch exceeds max_batch");
    if (batch == 0) return RST_OK;
    if (ctx->cfg.extractor == RST_EXTRACTOR_NONE) return fail(ctx, RST_ERR_STATE, "context was created without a style predictor");
    cudaSetDevice(ctx->device);
    const rst_config& g = ctx->cfg;
    cudaStream_t s = ctx->own_stream;
    const int S = g.num_styles, P = ctx->num_style_params;
    const long long img = (long long)g.style_h * g.style_w * 3;
    RST_CUDA(ctx, cudaMemcpyAsync(ctx->st_style, h_style, (size_t)batch * S * img * sizeof(float), cudaMemcpyHostToDevice, s));
    int64_t total_launches = 0;
    for (int si = 0; si < S; ++si) {
        const float* style_s = ctx->st_style;
        if (S > 1) {   // gather style si of every sample into a dense (B,H,W,3) batch (extra slot behind st_style)
            long long total = (long long)batch * img;
            float* dense = ctx->st_style + (size_t)batch * S * img;
            gather_style_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(ctx->st_style, dense, img, S, si, total);
            RST_CUDA(ctx, cudaGetLastError());
            style_s = dense;
            total_launches += 1;
        }
        int rc = rst_predict_style(ctx, style_s, ctx->pvec[2], batch, (void*)s);
        if (rc) return rc;
        total_launches += ctx->launches + 1;
        interleave_params_kernel<<<(unsigned)(((long long)batch * P + 255) / 256), 256, 0, s>>>(ctx->pvec[2], ctx->st_params,
                                                                                               batch, S, P, si);
        RST_CUDA(ctx, cudaGetLastError());
    }
    RST_CUDA(ctx, cudaMemcpyAsync(ctx->st_content, h_content, (size_t)batch * g.in_h * g.in_w * g.in_c * sizeof(float),
                                  cudaMemcpyHostToDevice, s));
    if (S > 1) {
        if (!h_style_weights) return fail(ctx, RST_ERR_INVALID, "style_weights required when num_styles > 1");
        RST_CUDA(ctx, cudaMemcpyAsync(ctx->st_weights, h_style_weights,
                                      (size_t)batch * g.out_h * g.out_w * (S - 1) * sizeof(float), cudaMemcpyHostToDevice, s));
    }
    int rc = rst_transfer_forward(ctx, ctx->st_content, ctx->st_params, S > 1 ? ctx->st_weights : nullptr, ctx->st_out, batch,
                                  (void*)s);
    if (rc) return rc;
    ctx->launches += total_launches;
    RST_CUDA(ctx, cudaMemcpyAsync(h_out, ctx->st_out, (size_t)batch * g.out_h * g.out_w * 3 * sizeof(float),
                                  cudaMemcpyDeviceToHost, s));
    RST_CUDA(ctx, cudaStreamSynchronize(s));
    return RST_OK;
}

// ------------------------------------------------------------------------------------------------
// debug / accounting
// ------------------------------------------------------------------------------------------------
extern "C" int rst_debug_enable_taps(rst_ctx* ctx, int enable) {
    if (!ctx) return RST_ERR_INVALID;
    ctx->keep_taps = enable != 0;
    return RST_OK;
}

extern "C" int rst_debug_tap(rst_ctx* ctx, const char* name, float* h_out, int64_t capacity, int64_t* elems) {
    if (!ctx || !name) return RST_ERR_INVALID;
    auto it = ctx->taps.find(name);
    if (it == ctx->taps.end()) return fail(ctx, RST_ERR_INVALID, std::string("rst_debug_tap: no such tap (enable taps first): ") + name);
    if (elems) *elems = it->second.elems;
    if (!h_out) return RST_OK;
    if (capacity < it->second.elems) return fail(ctx, RST_ERR_INVALID, "rst_debug_tap: buffer too small");
    cudaSetDevice(ctx->device);
    RST_CUDA(ctx, cudaDeviceSynchronize());
    RST_CUDA(ctx, cudaMemcpy(h_out, it->second.dev, (size_t)it->second.elems * sizeof(float), cudaMemcpyDeviceToHost));
    return RST_OK;
}

extern "C" int64_t rst_last_launch_count(const rst_ctx* ctx) { return ctx ? ctx->launches : -1; }

extern "C" int rst_profile_enable(rst_ctx* ctx, int enable) {
    if (!ctx) return RST_ERR_INVALID;
    ctx->profiling = enable != 0;
    return RST_OK;
}
extern "C" int rst_profile_reset(rst_ctx* ctx) {
    if (!ctx) return RST_ERR_INVALID;
    collect_profile(ctx);
    ctx->profile.clear();
    return RST_OK;
}
extern "C" int rst_profile_get(rst_ctx* ctx, const char* group, double* total_ms, int64_t* launches) {
    if (!ctx || !group) return RST_ERR_INVALID;
    cudaSetDevice(ctx->device);
    collect_profile(ctx);
    auto it = ctx->profile.find(group);
    if (total_ms) *total_ms = it == ctx->profile.end() ? 0.0 : it->second.total_ms;
    if (launches) *launches = it == ctx->profile.end() ? 0 : it->second.launches;
    return RST_OK;
}
extern "C" int rst_profile_group_count(rst_ctx* ctx) {
    if (!ctx) return -1;
    collect_profile(ctx);
    return (int)ctx->profile.size();
}
extern "C" const char* rst_profile_group_name(rst_ctx* ctx, int i) {
    if (!ctx) return nullptr;
    int k = 0;
    for (auto& kv : ctx->profile) if (k++ == i) return kv.first.c_str();
    return nullptr;
}

// ------------------------------------------------------------------------------------------------
// stand-alone operators
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_op_error;
extern "C" const char* rst_op_last_error(void) { return g_op_error.c_str(); }
static int op_fail(int code, const std::string& m) { g_op_error = m; return code; }
#define OP_CUDA(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) return op_fail(RST_ERR_CUDA, std::string(#expr ": ") + cudaGetErrorString(_e)); } while (0)

namespace rst {
int op_conv2d_bf16(const float* d_x, const float* d_kernel, const float* d_bias, float* d_y, int batch, int h, int w, int ci,
                   int co, int kh, int kw, int stride, int transposed, int act, cudaStream_t s, std::string* err);
}

extern "C" int rst_op_conv2d(const float* d_x, const float* d_kernel, const float* d_bias, float* d_y, int batch, int h, int w,
                             int ci, int co, int kh, int kw, int stride, int transposed, int act, int precision, void* stream) {
    if (batch == 0) return RST_OK;
    if (!d_x || !d_kernel || !d_y) return op_fail(RST_ERR_INVALID, "rst_op_conv2d: null tensor");
    if (batch < 0 || h <= 0 || w <= 0 || ci <= 0 || co <= 0 || kh <= 0 || kw <= 0 || stride <= 0)
        return op_fail(RST_ERR_INVALID, "rst_op_conv2d: bad dimension");
    cudaStream_t s = (cudaStream_t)stream;
    if (precision == RST_PRECISION_BF16) {
        std::string err;
        int rc = op_conv2d_bf16(d_x, d_kernel, d_bias, d_y, batch, h, w, ci, co, kh, kw, stride, transposed, act, s, &err);
        if (rc) return op_fail(rc, err);
        return RST_OK;
    }
    if (precision == RST_PRECISION_TF32 || precision == RST_PRECISION_TF32X3) {
        const bool split = precision == RST_PRECISION_TF32X3;
        if (kh != 3 || kw != 3 || stride != 1 || ci % 32 || co % 64 || (act != ACT_RELU && act != ACT_NONE))
            return op_fail(RST_ERR_UNSUPPORTED, "rst_op_conv2d(tf32): 3x3 stride-1 convs with Cin % 32 == 0, Cout % 64 == 0 only");
        std::vector<float> hk((size_t)9 * ci * co), hb(co, 0.f);
        OP_CUDA(cudaMemcpy(hk.data(), d_kernel, hk.size() * 4, cudaMemcpyDeviceToHost));
        if (d_bias) OP_CUDA(cudaMemcpy(hb.data(), d_bias, co * 4, cudaMemcpyDeviceToHost));
        Tf32Conv3x3 conv;
        std::string err;
        // a stride-1 'same' Conv2DTranspose with kernel (3,3,co,ci) is the input gradient of the conv layer co -> ci with that kernel
        const bool ok = transposed ? conv.setup(co, ci, hk.data(), d_bias ? hb.data() : nullptr, act == ACT_RELU, true, &err, split)
                                   : conv.setup(ci, co, hk.data(), d_bias ? hb.data() : nullptr, act == ACT_RELU, false, &err, split);
        if (!ok) return op_fail(RST_ERR_CUDA, err);
        float* scratch = nullptr;
        if (split) OP_CUDA(cudaMalloc(&scratch, conv.scratch_floats(batch, h, w) * sizeof(float)));
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaError_t e = conv.run_split(d_x, scratch, d_y, batch, h, w, sms, s, &err);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        if (scratch) cudaFree(scratch);
        if (e != cudaSuccess) return op_fail(RST_ERR_CUDA, err.empty() ? cudaGetErrorString(e) : err);
        return RST_OK;
    }
    ConvF32 p;
    p.x = d_x; p.w = d_kernel; p.bias = d_bias; p.y = d_y;
    p.B = batch; p.Hi = h; p.Wi = w; p.Ci = ci; p.Co = co; p.kh = kh; p.kw = kw; p.stride = stride;
    p.transposed = transposed ? 1 : 0;
    if (!transposed) {
        p.Ho = tf_same(h, kh, stride, &p.pad_t);
        p.Wo = tf_same(w, kw, stride, &p.pad_l);
        p.w_tap = (long long)ci * co; p.w_ci = co; p.w_co = 1;
    } else {
        p.Ho = h * stride; p.Wo = w * stride;
        tf_same(p.Ho, kh, stride, &p.pad_t);
        tf_same(p.Wo, kw, stride, &p.pad_l);
        p.w_tap = (long long)ci * co; p.w_ci = 1; p.w_co = ci;
    }
    p.act1 = act;
    OP_CUDA(launch_conv_f32(p, s));
    return RST_OK;
}

extern "C" int rst_op_cin(const float* d_x, const float* d_params, const float* d_weights, float* d_y, int batch, int h, int w,
                          int f, int num_styles, int act, void* stream) {
    if (!d_x || !d_params || !d_y) return op_fail(RST_ERR_INVALID, "rst_op_cin: null tensor");
    if (num_styles < 1 || num_styles > 2) return op_fail(RST_ERR_INVALID, "rst_op_cin: 1 or 2 styles");
    if (num_styles == 2 && !d_weights) return op_fail(RST_ERR_INVALID, "rst_op_cin: weights required for 2 styles");
    cudaStream_t s = (cudaStream_t)stream;
    double* stats = nullptr;
    OP_CUDA(cudaMalloc(&stats, (size_t)batch * f * 2 * sizeof(double)));
    cudaError_t e = launch_zero_f64(stats, (long long)batch * f * 2, s);
    if (e == cudaSuccess) e = launch_moments_f32(d_x, stats, batch, h * w, f, s);
    CinApply a;
    a.x = d_x; a.y = d_y; a.stats = stats; a.params = d_params;
    a.param_bstride = (long long)num_styles * 2 * f; a.param_sstride = 2 * f; a.scale_off = 0; a.bias_off = f;
    a.weights = d_weights; a.B = batch; a.P = h * w; a.C = f; a.num_styles = num_styles; a.act = act;
    if (e == cudaSuccess) e = launch_cin_apply(a, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    cudaFree(stats);
    OP_CUDA(e);
    return RST_OK;
}

extern "C" int rst_op_apply_style_weights(const float* d_weights, const float* d_params, float* d_out, int batch, int h, int w,
                                          int f, void* stream) {
    if (!d_weights || !d_params || !d_out) return op_fail(RST_ERR_INVALID, "rst_op_apply_style_weights: null tensor");
    OP_CUDA(launch_apply_style_weights(d_weights, d_params, d_out, batch, (long long)h * w, f, (cudaStream_t)stream));
    return RST_OK;
}

extern "C" int rst_op_gram(const float* d_x, float* d_gram, int batch, int h, int w, int c, void* stream) {
    if (!d_x || !d_gram) return op_fail(RST_ERR_INVALID, "rst_op_gram: null tensor");
    cudaStream_t s = (cudaStream_t)stream;
    if (gram_tf32_supported(c) && !ab_env("RST_GRAM_CUDA_CORE")) {
        // tensor cores, error-compensated split tf32 (fp32-level accuracy); a stand-alone operator may allocate its scratch
        std::string err;
        if (!umma_init(&err)) return op_fail(RST_ERR_CUDA, err);
        float* scratch = nullptr;
        OP_CUDA(cudaMalloc(&scratch, gram_tf32_scratch_floats(batch, h * w, c) * sizeof(float)));
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaError_t e = launch_gram_tf32(d_x, d_gram, scratch, batch, h * w, c, true, sms, s, &err);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        cudaFree(scratch);
        if (e != cudaSuccess) return op_fail(RST_ERR_CUDA, err.empty() ? cudaGetErrorString(e) : err);
        return RST_OK;
    }
    OP_CUDA(launch_gram_f32(d_x, d_gram, batch, h * w, c, s));
    return RST_OK;
}
