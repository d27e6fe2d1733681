// Training step (styleTransferTrainingModel.py:11-70 + Keras train_step): predictor and transfer network
// in training mode (BatchNorm on batch statistics, moving statistics updated with momentum 0.99), the VGG loss model of loss.cu,
// back-propagation through every layer and a Keras-RMSprop update.  fp32 CUDA-core kernels; the tensor-core training path is
// later work.
//
// The step is recorded on a small tape: every forward op launches its kernels and pushes the closure that back-propagates
// through it; the backward pass runs the closures in reverse.  A gradient buffer is written by its first producer and
// accumulated into by later ones (skip connections, MobileNet residuals, the style-parameter vector read by every CIN).
#include <cstring>
#include <deque>
#include <functional>

#include "rst_ctx.h"
#include "train_kernels.cuh"
#include "halo_gemm.cuh"
#include <memory>

using namespace rst;

namespace {

struct Tensor {
    std::string name;         // set for the tensors rst_train_debug_read exposes
    float* d = nullptr;
    float* g = nullptr;
    bool g_init = false, needs_grad = true;
    int B = 0, H = 0, W = 0, C = 0;
    long long n() const { return (long long)B * H * W * C; }
    int P() const { return H * W; }
};

struct Var {
    float* w = nullptr;
    float* g = nullptr;       // null: not trainable (BatchNorm moving statistics)
    float* slot = nullptr;    // RMSprop accumulator
    int64_t offset = -1, elems = 0;
};

struct Slot { void* p = nullptr; size_t bytes = 0; };

}  // namespace

struct rst_trainer {
    rst_ctx* m = nullptr;
    rst_loss* loss = nullptr;
    int device = 0;
    std::string err;
    float *weights_flat = nullptr, *grads_flat = nullptr, *slots_flat = nullptr;
    int64_t n_train = 0, n_total = 0;     // padded lengths (floats): trainable prefix / whole arena
    std::map<std::string, Var> vars;
    std::vector<Slot> slots;              // activations and scratch: step k reuses the allocations of step k-1
    size_t slot_cursor = 0;
    std::deque<Tensor> tensors;
    std::vector<std::function<int()>> tape;
    cudaStream_t s = nullptr;
    Tensor* pred = nullptr;
    Tensor* style_params = nullptr;
    int rc = RST_OK;                      // sticky error of the op being recorded
    // optional tf32 tensor-core path for the 3x3 stride-1 convolutions of the transfer network (Cin % 32 == 0, Cout % 64 == 0:
    // the residual trunk), forward and input gradient; weights are re-packed from the live variables every step
    int math = RST_PRECISION_FP32;
    int split_tf32 = 1;                   // fp32 math: 1 = split-tf32 tensor-core convs for the trunk, 0 = CUDA-core fp32 (RST_TRAIN_SPLIT_TF32=0)
    int num_sms = 148;
    std::map<std::string, std::unique_ptr<Tf32Conv3x3>> tf32_fwd, tf32_bwd;
};

static Tf32Conv3x3* tf32_conv(rst_trainer* t, std::map<std::string, std::unique_ptr<Tf32Conv3x3>>& cache, const std::string& name,
                              int ci, int co, bool relu, bool input_gradient, bool split) {
    auto it = cache.find(name);
    if (it != cache.end() && it->second->split == split) return it->second.get();
    std::unique_ptr<Tf32Conv3x3> c(new Tf32Conv3x3());
    std::string err;
    if (!c->setup_shape(ci, co, relu, input_gradient, &err, split)) {
        if (t->rc == RST_OK) t->rc = RST_ERR_CUDA, t->err = "training tf32 conv " + name + ": " + err;
        return nullptr;
    }
    return (cache[name] = std::move(c)).get();
}

static thread_local std::string g_train_create_error;

static int tfail(rst_trainer* t, int code, const std::string& msg) {
    if (t) t->err = msg; else g_train_create_error = msg;
    return code;
}
#define TCUDA(t, expr)                                                                                        \
    do {                                                                                                      \
        cudaError_t _e = (expr);                                                                              \
        if (_e != cudaSuccess) return tfail((t), RST_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    } while (0)
// inside recording ops (which return Tensor*): remember the first error and keep going with harmless pointers
#define OPCUDA(t, expr)                                                                                       \
    do {                                                                                                      \
        cudaError_t _e = (expr);                                                                              \
        if (_e != cudaSuccess && (t)->rc == RST_OK)                                                           \
            (t)->rc = tfail((t), RST_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));           \
    } while (0)

static void* salloc(rst_trainer* t, size_t bytes) {
    bytes = (bytes + 255) & ~(size_t)255;
    if (bytes == 0) bytes = 256;
    if (t->slot_cursor == t->slots.size()) t->slots.push_back(Slot());
    Slot& sl = t->slots[t->slot_cursor++];
    if (sl.bytes < bytes) {
        if (sl.p) { cudaStreamSynchronize(t->s); cudaFree(sl.p); }
        sl.p = nullptr; sl.bytes = 0;
        cudaError_t e = cudaMalloc(&sl.p, bytes);
        if (e != cudaSuccess) {
            if (t->rc == RST_OK) t->rc = tfail(t, RST_ERR_CUDA, std::string("training workspace cudaMalloc: ") + cudaGetErrorString(e));
            return nullptr;
        }
        sl.bytes = bytes;
    }
    return sl.p;
}
static float* falloc(rst_trainer* t, long long n) { return (float*)salloc(t, (size_t)n * sizeof(float)); }
static double* dalloc(rst_trainer* t, long long n) { return (double*)salloc(t, (size_t)n * sizeof(double)); }

static Tensor* new_tensor(rst_trainer* t, int B, int H, int W, int C, bool with_data = true) {
    t->tensors.emplace_back();
    Tensor* x = &t->tensors.back();
    x->B = B; x->H = H; x->W = W; x->C = C;
    if (with_data) x->d = falloc(t, x->n());
    x->g = falloc(t, x->n());
    return x;
}
static Tensor* input_tensor(rst_trainer* t, const float* d, int B, int H, int W, int C) {
    t->tensors.emplace_back();
    Tensor* x = &t->tensors.back();
    x->B = B; x->H = H; x->W = W; x->C = C;
    x->d = const_cast<float*>(d);
    x->needs_grad = false;
    return x;
}
static Var* var(rst_trainer* t, const std::string& name) {
    auto it = t->vars.find(name);
    if (it == t->vars.end()) {
        if (t->rc == RST_OK) t->rc = tfail(t, RST_ERR_STATE, "training: unknown variable " + name);
        return nullptr;
    }
    return &it->second;
}

// sum over (n, pixels) of g per channel, accumulated into a bias gradient
static int bias_grad(rst_trainer* t, const float* g, int B, int P, int C, float* db, double* st, double* st2) {
    TCUDA(t, launch_zero_f64(st, (long long)B * C * 2, t->s));
    TCUDA(t, launch_moments_f32(g, st, B, P, C, t->s));
    TCUDA(t, launch_reduce_over_batch(st, st2, B, C * 2, t->s));
    TCUDA(t, launch_gather_f64(st2, db, C, 2, 0, 1, t->s));
    t->m->launches += 4;
    return RST_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// ops
// ---------------------------------------------------------------------------------------------------------------------
struct ConvSpec {
    int co = 0, k = 1, stride = 1, act = ACT_NONE;
    bool transposed = false;       // Conv2DTranspose, kernel (kh,kw,out,in)
    float in_scale = 1.f, in_shift = 0.f;
};

// Conv2D / Conv2DTranspose, padding 'same', optional bias, optional activation
static Tensor* op_conv(rst_trainer* t, Tensor* x, const std::string& kname, const std::string& bname, const ConvSpec& cs) {
    Var* vk = var(t, kname);
    Var* vb = bname.empty() ? nullptr : var(t, bname);
    if (!vk) return x;
    int pad_t = 0, pad_l = 0, ho, wo;
    if (!cs.transposed) { ho = tf_same(x->H, cs.k, cs.stride, &pad_t); wo = tf_same(x->W, cs.k, cs.stride, &pad_l); }
    else { ho = x->H * cs.stride; wo = x->W * cs.stride; tf_same(ho, cs.k, cs.stride, &pad_t); tf_same(wo, cs.k, cs.stride, &pad_l); }
    Tensor* y = new_tensor(t, x->B, ho, wo, cs.co);
    const int ci = x->C, co = cs.co;
    ConvF32 p;
    p.x = x->d; p.y = y->d; p.w = vk->w; p.bias = vb ? vb->w : nullptr;
    p.B = x->B; p.Hi = x->H; p.Wi = x->W; p.Ci = ci; p.Ho = ho; p.Wo = wo; p.Co = co;
    p.kh = cs.k; p.kw = cs.k; p.stride = cs.stride; p.pad_t = pad_t; p.pad_l = pad_l;
    p.transposed = cs.transposed ? 1 : 0;
    p.w_tap = (long long)ci * co;
    if (!cs.transposed) { p.w_ci = co; p.w_co = 1; } else { p.w_ci = 1; p.w_co = ci; }
    p.in_scale = cs.in_scale; p.in_shift = cs.in_shift;
    p.act1 = cs.act;
    // 3x3 stride-1 convs with Cin % 32 == 0, Cout % 64 == 0 (the residual trunk) run on the tensor cores: plain tf32 operands
    // under RST_PRECISION_TF32, error-compensated split tf32 (fp32-level accuracy) otherwise
    const bool tensor_core = t->split_tf32 >= 0 && !cs.transposed && cs.k == 3 && cs.stride == 1 && ci % 32 == 0 && co % 64 == 0 &&
                             (cs.act == ACT_RELU || cs.act == ACT_NONE) && cs.in_scale == 1.f && cs.in_shift == 0.f &&
                             (t->math == RST_PRECISION_TF32 || t->split_tf32 == 1);
    const bool split = tensor_core && t->math != RST_PRECISION_TF32;
    Tf32Conv3x3* fwd = (tensor_core && !exp_env("RST_EXP_TRUNK_FWD_FP32")) ? tf32_conv(t, t->tf32_fwd, kname, ci, co, cs.act == ACT_RELU, false, split) : nullptr;
    if (fwd) {
        std::string err;
        float* scratch = split ? falloc(t, (long long)fwd->scratch_floats(x->B, x->H, x->W)) : nullptr;
        OPCUDA(t, fwd->repack(vk->w, vb ? vb->w : nullptr, t->s));
        OPCUDA(t, fwd->run_split(x->d, scratch, y->d, x->B, x->H, x->W, t->num_sms, t->s, &err));
        t->m->launches += 1 + fwd->nblk + (split ? 1 : 0);
    } else {
        OPCUDA(t, launch_conv_f32(p, t->s));
        t->m->launches += 1;
    }
    float* dgrad_tmp = (tensor_core && x->needs_grad) ? falloc(t, x->n()) : nullptr;
    float* dgrad_scratch = (split && x->needs_grad) ? falloc(t, 2LL * y->n()) : nullptr;   // [g_hi | g_lo] for the input gradient; [x_lo | g_lo] for the weight gradient before it
    // 128 -> 128: the weight gradient runs on the tensor cores as well (wgrad_tf32.cu); its scratch is free until the input gradient
    const char* wgrad_env = ab_env("RST_WGRAD_TF32");                  // read per step: the tests toggle it
    const bool wgrad_tc_off = wgrad_env && wgrad_env[0] == '0';
    const bool wgrad_tc = tensor_core && ci == 128 && co == 128 && !wgrad_tc_off;
    float* wgrad_scratch = (wgrad_tc && split) ? (dgrad_scratch ? dgrad_scratch : falloc(t, 2LL * y->n())) : nullptr;
    double* st = dalloc(t, (long long)x->B * co * 2);
    double* st2 = dalloc(t, (long long)co * 2);
    const ConvSpec spec = cs;
    t->tape.push_back([=]() -> int {
        if (!y->g_init) return RST_OK;
        TCUDA(t, launch_act_bwd(y->g, y->d, spec.act, y->n(), t->s));
        WgradF32 wg;
        wg.x = x->d; wg.g = y->g; wg.dw = vk->g;
        wg.B = x->B; wg.Hx = x->H; wg.Wx = x->W; wg.Ci = ci; wg.Hg = ho; wg.Wg = wo; wg.Co = co;
        wg.kh = spec.k; wg.kw = spec.k; wg.stride = spec.stride; wg.pad_t = pad_t; wg.pad_l = pad_l;
        wg.transposed = spec.transposed ? 1 : 0;
        wg.Hb = spec.transposed ? x->H : ho; wg.Wb = spec.transposed ? x->W : wo;
        wg.in_scale = spec.in_scale; wg.in_shift = spec.in_shift;
        if (wgrad_tc) {
            std::string err;
            TCUDA(t, launch_wgrad_tf32(x->d, y->g, vk->g, wgrad_scratch, x->B, x->H, x->W, split, t->num_sms, t->s, &err));
            t->m->launches += split ? 2 : 0;
        } else {
            TCUDA(t, launch_wgrad_f32(wg, t->s));
        }
        t->m->launches += 2;
        if (vb) { int rc = bias_grad(t, y->g, y->B, y->P(), co, vb->g, st, st2); if (rc) return rc; }
        Tf32Conv3x3* bwd = (tensor_core && x->needs_grad && !exp_env("RST_EXP_TRUNK_BWD_FP32")) ? tf32_conv(t, t->tf32_bwd, kname, ci, co, false, true, split) : nullptr;
        if (bwd) {
            std::string err;
            float* dst = x->g_init ? dgrad_tmp : x->g;
            TCUDA(t, bwd->repack(vk->w, nullptr, t->s));
            TCUDA(t, bwd->run_split(y->g, dgrad_scratch, dst, x->B, x->H, x->W, t->num_sms, t->s, &err));
            if (x->g_init) TCUDA(t, launch_add_inplace(x->g, dgrad_tmp, x->n(), t->s));
            t->m->launches += 2 + bwd->nblk;
            x->g_init = true;
        } else if (x->needs_grad) {
            ConvF32 q;     // input gradient: the adjoint map with the channel roles of the same kernel swapped
            q.x = y->g; q.y = x->g; q.w = vk->w;
            q.B = x->B; q.Hi = ho; q.Wi = wo; q.Ci = co; q.Ho = x->H; q.Wo = x->W; q.Co = ci;
            q.kh = spec.k; q.kw = spec.k; q.stride = spec.stride; q.pad_t = pad_t; q.pad_l = pad_l;
            q.transposed = spec.transposed ? 0 : 1;
            q.w_tap = (long long)ci * co;
            if (!spec.transposed) { q.w_ci = 1; q.w_co = co; } else { q.w_ci = ci; q.w_co = 1; }
            q.residual = x->g_init ? x->g : nullptr;
            TCUDA(t, launch_conv_f32(q, t->s));
            t->m->launches += 1;
            x->g_init = true;
        }
        return RST_OK;
    });
    return y;
}

// shared tail of the two normalisations: y = act(x*a + b) (+ residual), and the backward closure
static Tensor* norm_apply(rst_trainer* t, Tensor* x, bool per_sample, float* mean, float* inv, float* a, float* b, int act,
                          Tensor* residual, std::function<int(const double* r)> param_grads) {
    Tensor* y = new_tensor(t, x->B, x->H, x->W, x->C);
    OPCUDA(t, launch_affine_act(x->d, y->d, a, b, residual ? residual->d : nullptr, x->B, x->P(), x->C, per_sample ? 1 : 0, act, t->s));
    t->m->launches += 1;
    const int B = x->B, C = x->C;
    double* r = dalloc(t, (long long)B * C * 2);
    double* r2 = per_sample ? nullptr : dalloc(t, (long long)C * 2);
    t->tape.push_back([=]() -> int {
        if (!y->g_init) return RST_OK;
        if (residual && residual->needs_grad) {
            if (residual->g_init) TCUDA(t, launch_add_inplace(residual->g, y->g, y->n(), t->s));
            else TCUDA(t, cudaMemcpyAsync(residual->g, y->g, (size_t)y->n() * sizeof(float), cudaMemcpyDeviceToDevice, t->s));
            residual->g_init = true;
            t->m->launches += 1;
        }
        TCUDA(t, launch_affine_act_bwd(y->g, x->d, a, b, B, x->P(), C, per_sample ? 1 : 0, act, t->s));
        TCUDA(t, launch_zero_f64(r, (long long)B * C * 2, t->s));
        TCUDA(t, launch_norm_bwd_reduce(y->g, x->d, mean, inv, r, B, x->P(), C, per_sample ? 1 : 0, t->s));
        const double* rr = r;
        if (!per_sample) { TCUDA(t, launch_reduce_over_batch(r, r2, B, C * 2, t->s)); rr = r2; }
        int rc = param_grads(rr);
        if (rc) return rc;
        if (x->needs_grad) {
            TCUDA(t, launch_norm_bwd_apply(y->g, x->d, mean, inv, a, rr, x->g, B, x->P(), C, per_sample ? 1 : 0,
                                           per_sample ? (double)x->P() : (double)B * x->P(), x->g_init ? 1 : 0, t->s));
            x->g_init = true;
        }
        t->m->launches += 5;
        return RST_OK;
    });
    return y;
}

// BatchNormalization in training mode (batch statistics; Keras momentum 0.99) + activation (+ residual)
static Tensor* op_bn(rst_trainer* t, Tensor* x, const std::string& prefix, float eps, float momentum, int act,
                     Tensor* residual = nullptr) {
    Var* vg = var(t, prefix + "/gamma");
    Var* vb = var(t, prefix + "/beta");
    Var* vm = var(t, prefix + "/moving_mean");
    Var* vv = var(t, prefix + "/moving_variance");
    if (!vg || !vb || !vm || !vv) return x;
    const int B = x->B, C = x->C;
    double* st = dalloc(t, (long long)B * C * 2);
    double* st2 = dalloc(t, (long long)C * 2);
    float* co = falloc(t, 4LL * C);
    float *mean = co, *inv = co + C, *a = co + 2 * C, *b = co + 3 * C;
    OPCUDA(t, launch_zero_f64(st, (long long)B * C * 2, t->s));
    OPCUDA(t, launch_moments_f32(x->d, st, B, x->P(), C, t->s));
    OPCUDA(t, launch_reduce_over_batch(st, st2, B, C * 2, t->s));
    OPCUDA(t, launch_norm_finalize(st2, 1, C, (double)B * x->P(), eps, vg->w, vb->w, 0, mean, inv, a, b, vm->w, vv->w, momentum, t->s));
    t->m->launches += 4;
    return norm_apply(t, x, false, mean, inv, a, b, act, residual, [=](const double* r) -> int {
        TCUDA(t, launch_gather_f64(r, vb->g, C, 2, 0, 1, t->s));     // d beta  = sum g
        TCUDA(t, launch_gather_f64(r, vg->g, C, 2, 1, 1, t->s));     // d gamma = sum g * xhat
        t->m->launches += 2;
        return RST_OK;
    });
}

// ConditionalInstanceNormalization with one style (styleTransfer.py:57-71): scale/bias are columns of the predictor output
static Tensor* op_cin(rst_trainer* t, Tensor* x, Tensor* params, int off, int act, Tensor* residual = nullptr) {
    const int B = x->B, C = x->C;
    const long long ptotal = params->C;
    double* st = dalloc(t, (long long)B * C * 2);
    float* co = falloc(t, 4LL * B * C);
    float *mean = co, *inv = co + (long long)B * C, *a = co + 2LL * B * C, *b = co + 3LL * B * C;
    OPCUDA(t, launch_zero_f64(st, (long long)B * C * 2, t->s));
    OPCUDA(t, launch_moments_f32(x->d, st, B, x->P(), C, t->s));
    OPCUDA(t, launch_norm_finalize(st, B, C, (double)x->P(), 1e-5f, params->d + off, params->d + off + C, ptotal, mean, inv, a, b,
                                   nullptr, nullptr, 0.f, t->s));
    t->m->launches += 3;
    return norm_apply(t, x, true, mean, inv, a, b, act, residual, [=](const double* r) -> int {
        if (!params->g_init) {
            TCUDA(t, cudaMemsetAsync(params->g, 0, (size_t)params->n() * sizeof(float), t->s));
            params->g_init = true;
        }
        TCUDA(t, launch_cin_param_grad(r, params->g, B, C, ptotal, off, t->s));
        t->m->launches += 1;
        return RST_OK;
    });
}

// global average pooling over H, W (keepdims)
static Tensor* op_gap(rst_trainer* t, Tensor* x) {
    Tensor* y = new_tensor(t, x->B, 1, 1, x->C);
    double* st = dalloc(t, (long long)x->B * x->C * 2);
    OPCUDA(t, launch_zero_f64(st, (long long)x->B * x->C * 2, t->s));
    OPCUDA(t, launch_moments_f32(x->d, st, x->B, x->P(), x->C, t->s));
    OPCUDA(t, launch_stats_to_mean(st, y->d, x->B, x->P(), x->C, t->s));
    t->m->launches += 3;
    t->tape.push_back([=]() -> int {
        if (!y->g_init || !x->needs_grad) return RST_OK;
        TCUDA(t, launch_gap_bwd(y->g, x->g, x->B, x->P(), x->C, x->g_init ? 1 : 0, t->s));
        x->g_init = true;
        t->m->launches += 1;
        return RST_OK;
    });
    return y;
}

// depthwise convolution (MobileNetV3 blocks), explicit padding
static Tensor* op_depthwise(rst_trainer* t, Tensor* x, const std::string& wname, int k, int stride, int pad_t, int pad_l, int ho, int wo) {
    Var* vw = var(t, wname);
    if (!vw) return x;
    Tensor* y = new_tensor(t, x->B, ho, wo, x->C);
    DepthwiseF32 d;
    d.x = x->d; d.y = y->d; d.w = vw->w;
    d.B = x->B; d.Hi = x->H; d.Wi = x->W; d.C = x->C; d.k = k; d.stride = stride; d.pad_t = pad_t; d.pad_l = pad_l; d.Ho = ho; d.Wo = wo;
    OPCUDA(t, launch_depthwise_f32(d, t->s));
    t->m->launches += 1;
    t->tape.push_back([=]() -> int {
        if (!y->g_init) return RST_OK;
        TCUDA(t, launch_depthwise_wgrad(d, y->g, vw->g, t->s));
        if (x->needs_grad) {
            TCUDA(t, launch_depthwise_dgrad(d, y->g, x->g, x->g_init ? 1 : 0, t->s));
            x->g_init = true;
        }
        t->m->launches += 2;
        return RST_OK;
    });
    return y;
}

// squeeze-excite multiply: y[n,p,c] = x[n,p,c] * z[n,c]
static Tensor* op_scale_channels(rst_trainer* t, Tensor* x, Tensor* z) {
    Tensor* y = new_tensor(t, x->B, x->H, x->W, x->C);
    OPCUDA(t, launch_scale_channels(x->d, z->d, y->d, x->B, x->P(), x->C, t->s));
    t->m->launches += 1;
    double* r = dalloc(t, (long long)x->B * x->C * 2);
    t->tape.push_back([=]() -> int {
        if (!y->g_init) return RST_OK;
        TCUDA(t, launch_zero_f64(r, (long long)x->B * x->C * 2, t->s));
        TCUDA(t, launch_norm_bwd_reduce(y->g, x->d, nullptr, nullptr, r, x->B, x->P(), x->C, 1, t->s));
        TCUDA(t, launch_gather_f64(r, z->g, x->B * x->C, 2, 1, z->g_init ? 1 : 0, t->s));     // dz = sum_p g * x
        z->g_init = true;
        TCUDA(t, launch_scale_channels_bwd(y->g, z->d, x->g, x->B, x->P(), x->C, x->g_init ? 1 : 0, t->s));
        x->g_init = true;
        t->m->launches += 4;
        return RST_OK;
    });
    return y;
}

// ---------------------------------------------------------------------------------------------------------------------
// networks
// ---------------------------------------------------------------------------------------------------------------------
// create_style_prediction_model (stylePrediction.py:25-75), training mode
static const float kMbMomentum = 0.999f;      // keras.applications MobileNetV3 BatchNormalization(momentum=0.999, epsilon=1e-3)
static Tensor* record_predictor(rst_trainer* t, Tensor* style) {
    rst_ctx* c = t->m;
    Tensor* x = style;
    if (c->cfg.extractor == RST_EXTRACTOR_DUMMY) {
        ConvSpec cs; cs.co = 1; cs.k = 9; cs.stride = 5;
        x = op_conv(t, x, "dummy_conv/kernel", "dummy_conv/bias", cs);
    } else {
        ConvSpec stem; stem.co = 16; stem.k = 3; stem.stride = 2; stem.in_scale = 2.f; stem.in_shift = -1.f;   // Rescaling(2, -1)
        x = op_conv(t, x, "mobilenet/Conv/kernel", "", stem);
        x = op_bn(t, x, "mobilenet/Conv/BatchNorm", 1e-3f, kMbMomentum, ACT_HSWISH);
        for (auto& m : c->mb_blocks) {
            Tensor* in = x;
            Tensor* e = x;
            if (m.block_id) {
                ConvSpec ex; ex.co = m.cexp;
                e = op_conv(t, e, m.prefix + "/expand/kernel", "", ex);
                e = op_bn(t, e, m.prefix + "/expand/BatchNorm", 1e-3f, kMbMomentum, m.act);
            }
            int pt, pl, ho, wo;
            if (m.s == 2) {     // ZeroPadding2D(correct_pad) + 'valid'
                pt = m.k / 2 - (1 - e->H % 2); pl = m.k / 2 - (1 - e->W % 2);
                ho = (e->H + pt + m.k / 2 - m.k) / 2 + 1;
                wo = (e->W + pl + m.k / 2 - m.k) / 2 + 1;
            } else {
                ho = tf_same(e->H, m.k, 1, &pt);
                wo = tf_same(e->W, m.k, 1, &pl);
            }
            Tensor* d = op_depthwise(t, e, m.prefix + "/depthwise/depthwise_kernel", m.k, m.s, pt, pl, ho, wo);
            d = op_bn(t, d, m.prefix + "/depthwise/BatchNorm", 1e-3f, kMbMomentum, m.act);
            if (m.se) {
                Tensor* z = op_gap(t, d);
                ConvSpec s1; s1.co = m.se; s1.act = ACT_RELU;
                z = op_conv(t, z, m.prefix + "/squeeze_excite/Conv/kernel", m.prefix + "/squeeze_excite/Conv/bias", s1);
                ConvSpec s2; s2.co = m.cexp; s2.act = ACT_HSIGMOID;
                z = op_conv(t, z, m.prefix + "/squeeze_excite/Conv_1/kernel", m.prefix + "/squeeze_excite/Conv_1/bias", s2);
                d = op_scale_channels(t, d, z);
            }
            ConvSpec pr; pr.co = m.cout;
            Tensor* p = op_conv(t, d, m.prefix + "/project/kernel", "", pr);
            const bool add = m.s == 1 && m.cin == m.cout;
            x = op_bn(t, p, m.prefix + "/project/BatchNorm", 1e-3f, kMbMomentum, ACT_NONE, add ? in : nullptr);
        }
        ConvSpec last; last.co = c->mb_last;
        x = op_conv(t, x, "mobilenet/Conv_1/kernel", "", last);
        x = op_bn(t, x, "mobilenet/Conv_1/BatchNorm", 1e-3f, kMbMomentum, ACT_HSWISH);
    }
    x = op_gap(t, x);
    ConvSpec d1; d1.co = 100;
    x = op_conv(t, x, "StylePredictor/kernel", "StylePredictor/bias", d1);
    ConvSpec d2; d2.co = c->num_style_params;
    x = op_conv(t, x, "StyleNormPredictor/kernel", "StyleNormPredictor/bias", d2);
    x->name = "style_params";
    return x;       // (B,1,1,P)
}

// create_style_transfer_model (styleTransfer.py:213-332) with num_styles == 1, training mode
static Tensor* record_transfer(rst_trainer* t, Tensor* content, Tensor* params) {
    rst_ctx* c = t->m;
    Tensor* x = content;
    for (auto& L : c->contract) {                                           // contract_block :188-205
        ConvSpec cs; cs.co = L.co; cs.k = L.k; cs.stride = L.stride; cs.act = ACT_RELU;
        x = op_conv(t, x, L.name + "/conv/kernel", L.name + "/conv/bias", cs);
        x->name = L.name + "/conv";
        x = op_bn(t, x, L.name + "/bn", 1e-3f, 0.99f, ACT_RELU);     // Keras BatchNormalization() defaults
        x->name = L.name + "/out";
    }
    const int F = c->cfg.bottleneck_num_filters;
    int cursor = 0;
    for (int b = 0; b < 5; ++b) {                                            // residual_block :144-185
        Tensor* skip = b == 0 ? nullptr : x;
        Tensor* fx = x;
        for (int i = 0; i < 2; ++i) {
            const LayerDesc& L = c->residual[2 * b + i];
            ConvSpec cs; cs.co = F; cs.k = 3; cs.act = ACT_RELU;
            fx = op_conv(t, fx, L.name + "/kernel", L.name + "/bias", cs);
            fx->name = L.name + "/conv";
            fx = op_cin(t, fx, params, cursor + 2 * F * i, i == 0 ? ACT_RELU : ACT_NONE, i == 1 ? skip : nullptr);
            fx->name = L.name + "/out";
        }
        cursor += 4 * F;
        x = fx;
    }
    for (size_t i = 0; i < c->expand.size(); ++i) {                          // expand_block :95-141
        const LayerDesc& L = c->expand[i];
        ConvSpec cs; cs.co = L.co; cs.k = L.k; cs.stride = L.stride; cs.transposed = true;
        x = op_conv(t, x, L.name + "/conv/kernel", L.name + "/conv/bias", cs);
        x->name = L.name + "/conv";
        x = op_cin(t, x, params, cursor, i + 1 == c->expand.size() ? ACT_SIGMOID : ACT_RELU);
        x->name = L.name + "/out";
        cursor += 2 * L.co;
    }
    return x;
}

// ---------------------------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------------------------
extern "C" const char* rst_train_last_error(const rst_trainer* t) { return t ? t->err.c_str() : g_train_create_error.c_str(); }

extern "C" int rst_train_destroy(rst_trainer* t) {
    if (!t) return RST_OK;
    cudaSetDevice(t->device);
    cudaDeviceSynchronize();
    for (auto& sl : t->slots) if (sl.p) cudaFree(sl.p);
    if (t->loss) rst_loss_destroy(t->loss);
    if (t->m) rst_destroy(t->m);              // frees weights_flat through ctx->weight_arena
    if (t->grads_flat) cudaFree(t->grads_flat);
    if (t->slots_flat) cudaFree(t->slots_flat);
    if (t->s) cudaStreamDestroy(t->s);
    delete t;
    return RST_OK;
}

extern "C" int rst_train_create(const rst_config* cfg, int device, rst_trainer** out) {
    if (!cfg || !out) return tfail(nullptr, RST_ERR_INVALID, "rst_train_create: null argument");
    *out = nullptr;
    if (cfg->num_styles != 1)     // make_style_transfer_training_model forces num_styles=1 (styleTransferTrainingModel.py:50-52)
        return tfail(nullptr, RST_ERR_INVALID, "rst_train_create: the training model has exactly one style");
    if (cfg->extractor == RST_EXTRACTOR_NONE)
        return tfail(nullptr, RST_ERR_INVALID, "rst_train_create: the training model needs a style predictor (extractor)");
    rst_config mc = *cfg;
    mc.precision = RST_PRECISION_FP32;
    rst_trainer* t = new rst_trainer();
    t->device = device;
    int rc = rst_create(&mc, device, &t->m);
    if (rc != RST_OK) { g_train_create_error = rst_last_error(nullptr); delete t; return rc; }
    rc = rst_loss_create(cfg->out_h, cfg->out_w, cfg->max_batch, device, &t->loss);
    if (rc != RST_OK) { g_train_create_error = rst_loss_last_error(nullptr); rst_train_destroy(t); return rc; }
    cudaSetDevice(device);
    cudaDeviceGetAttribute(&t->num_sms, cudaDevAttrMultiProcessorCount, device);
    if (const char* env = ab_env("RST_TRAIN_SPLIT_TF32")) t->split_tf32 = env[0] == '0' ? 0 : 1;
    // one arena for every variable: trainable ones first, so that gradients / RMSprop slots are flat arrays of the same layout
    auto padded = [](int64_t n) { return (n + 63) / 64 * 64; };
    auto trainable = [](const std::string& n) {
        auto ends = [&](const char* sfx) { size_t l = strlen(sfx); return n.size() >= l && n.compare(n.size() - l, l, sfx) == 0; };
        return !ends("/moving_mean") && !ends("/moving_variance");
    };
    int64_t off = 0;
    for (auto& w : t->m->weights) if (trainable(w.name)) { t->vars[w.name].offset = off; t->vars[w.name].elems = w.elems(); off += padded(w.elems()); }
    t->n_train = off;
    for (auto& w : t->m->weights) if (!trainable(w.name)) { t->vars[w.name].offset = off; t->vars[w.name].elems = w.elems(); off += padded(w.elems()); }
    t->n_total = off;
    bool ok = cudaMalloc(&t->weights_flat, (size_t)t->n_total * sizeof(float)) == cudaSuccess;
    ok = ok && cudaMalloc(&t->grads_flat, (size_t)t->n_train * sizeof(float)) == cudaSuccess;
    ok = ok && cudaMalloc(&t->slots_flat, (size_t)t->n_train * sizeof(float)) == cudaSuccess;
    ok = ok && cudaStreamCreateWithFlags(&t->s, cudaStreamNonBlocking) == cudaSuccess;
    if (!ok) { rst_train_destroy(t); return tfail(nullptr, RST_ERR_CUDA, "rst_train_create: device allocation failed"); }
    cudaMemset(t->weights_flat, 0, (size_t)t->n_total * sizeof(float));
    cudaMemset(t->grads_flat, 0, (size_t)t->n_train * sizeof(float));
    cudaMemset(t->slots_flat, 0, (size_t)t->n_train * sizeof(float));
    t->m->weight_arena = t->weights_flat;
    for (auto& w : t->m->weights) {
        Var& v = t->vars[w.name];
        v.w = t->weights_flat + v.offset;
        w.dev = v.w;
        if (v.offset < t->n_train) { v.g = t->grads_flat + v.offset; v.slot = t->slots_flat + v.offset; }
    }
    *out = t;
    return RST_OK;
}

// RST_PRECISION_FP32 (default) or RST_PRECISION_TF32: tensor-core tf32 arithmetic for the loss model (rst_loss_set_math) AND for
// the 3x3 convolutions of the transfer network's residual trunk, forward and input gradient (weight gradients stay fp32).
extern "C" int rst_train_set_math(rst_trainer* t, int precision) {
    if (!t) return RST_ERR_INVALID;
    if (precision != RST_PRECISION_FP32 && precision != RST_PRECISION_TF32)
        return tfail(t, RST_ERR_INVALID, "rst_train_set_math: RST_PRECISION_FP32 or RST_PRECISION_TF32");
    t->math = precision;
    return rst_loss_set_math(t->loss, precision);
}

extern "C" rst_ctx* rst_train_model(rst_trainer* t) { return t ? t->m : nullptr; }
extern "C" rst_loss* rst_train_loss(rst_trainer* t) { return t ? t->loss : nullptr; }
extern "C" int64_t rst_train_num_gradient_elements(const rst_trainer* t) { return t ? t->n_train : -1; }
extern "C" float* rst_train_gradients(rst_trainer* t) { return t ? t->grads_flat : nullptr; }
extern "C" const float* rst_train_prediction(const rst_trainer* t) { return t && t->pred ? t->pred->d : nullptr; }

extern "C" int rst_train_variable_range(const rst_trainer* t, const char* name, int64_t* offset, int64_t* elems) {
    if (!t || !name) return RST_ERR_INVALID;
    auto it = t->vars.find(name);
    if (it == t->vars.end() || it->second.offset >= t->n_train)
        return tfail(const_cast<rst_trainer*>(t), RST_ERR_INVALID, std::string("rst_train_variable_range: not a trainable variable: ") + name);
    if (offset) *offset = it->second.offset;
    if (elems) *elems = it->second.elems;
    return RST_OK;
}

extern "C" int rst_train_forward_backward(rst_trainer* t, const float* d_content, const float* d_style, const float* d_gt_content,
                                          const float* d_gt_style, float* d_losses, int batch) {
    if (!t) return RST_ERR_INVALID;
    rst_ctx* c = t->m;
    if (!c->committed) return tfail(t, RST_ERR_STATE, "rst_train_forward_backward: weights not committed");
    if (!d_content || !d_style || !d_gt_content || !d_gt_style || !d_losses)
        return tfail(t, RST_ERR_INVALID, "rst_train_forward_backward: null tensor");
    if (batch < 1 || batch > c->cfg.max_batch) return tfail(t, RST_ERR_INVALID, "rst_train_forward_backward: batch outside 1..max_batch");
    cudaSetDevice(t->device);
    t->slot_cursor = 0;
    t->tensors.clear();
    t->tape.clear();
    t->rc = RST_OK;
    t->pred = nullptr;
    c->launches = 0;
    TCUDA(t, cudaMemsetAsync(t->grads_flat, 0, (size_t)t->n_train * sizeof(float), t->s));
    Tensor* style = input_tensor(t, d_style, batch, c->cfg.style_h, c->cfg.style_w, 3);
    Tensor* content = input_tensor(t, d_content, batch, c->cfg.in_h, c->cfg.in_w, c->cfg.in_c);
    Tensor* params = record_predictor(t, style);
    Tensor* y = record_transfer(t, content, params);
    if (t->rc != RST_OK) return t->rc;
    t->pred = y;
    t->style_params = params;
    int rc = rst_loss_forward(t->loss, y->d, d_gt_content, d_gt_style, d_losses, batch, t->s);
    if (!rc) rc = rst_loss_backward(t->loss, y->d, y->g, batch, t->s);
    if (rc) return tfail(t, rc, std::string("loss model: ") + rst_loss_last_error(t->loss));
    y->g_init = true;
    for (auto it = t->tape.rbegin(); it != t->tape.rend(); ++it) {
        rc = (*it)();
        if (rc) return rc;
    }
    TCUDA(t, cudaStreamSynchronize(t->s));
    return RST_OK;
}

extern "C" void* rst_train_stream(rst_trainer* t) { return t ? (void*)t->s : nullptr; }

extern "C" int rst_train_apply_gradients(rst_trainer* t, float learning_rate, float rho, float epsilon) {
    if (!t) return RST_ERR_INVALID;
    cudaSetDevice(t->device);
    TCUDA(t, launch_rmsprop(t->weights_flat, t->grads_flat, t->slots_flat, learning_rate, rho, epsilon, t->n_train, t->s));
    TCUDA(t, cudaStreamSynchronize(t->s));
    t->m->launches += 1;
    for (auto& g : t->m->graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
    t->m->graphs.clear();
    return RST_OK;
}

// device -> host registry, then re-commit: rst_get_weight, checkpoints and the inference entry points of rst_train_model()
// see the trained values (BatchNorm folding included)
extern "C" int rst_train_sync_weights(rst_trainer* t) {
    if (!t) return RST_ERR_INVALID;
    cudaSetDevice(t->device);
    TCUDA(t, cudaStreamSynchronize(t->s));
    for (auto& w : t->m->weights) {
        w.host.resize((size_t)w.elems());
        TCUDA(t, cudaMemcpy(w.host.data(), w.dev, (size_t)w.elems() * sizeof(float), cudaMemcpyDeviceToHost));
        w.set = true;
    }
    int rc = rst_commit_weights(t->m);
    if (rc) return tfail(t, rc, rst_last_error(t->m));
    return RST_OK;
}

extern "C" int rst_train_read_gradient(rst_trainer* t, const char* name, float* h_out, int64_t capacity) {
    if (!t || !name || !h_out) return RST_ERR_INVALID;
    auto it = t->vars.find(name);
    if (it == t->vars.end() || !it->second.g) return tfail(t, RST_ERR_INVALID, std::string("rst_train_read_gradient: not trainable: ") + name);
    if (capacity < it->second.elems) return tfail(t, RST_ERR_INVALID, "rst_train_read_gradient: buffer too small");
    cudaSetDevice(t->device);
    TCUDA(t, cudaMemcpy(h_out, it->second.g, (size_t)it->second.elems * sizeof(float), cudaMemcpyDeviceToHost));
    return RST_OK;
}

extern "C" int rst_train_read_prediction(rst_trainer* t, float* h_out, int64_t capacity) {
    if (!t || !h_out) return RST_ERR_INVALID;
    if (!t->pred) return tfail(t, RST_ERR_STATE, "rst_train_read_prediction: no step has run");
    if (capacity < t->pred->n()) return tfail(t, RST_ERR_INVALID, "rst_train_read_prediction: buffer too small");
    cudaSetDevice(t->device);
    TCUDA(t, cudaMemcpy(h_out, t->pred->d, (size_t)t->pred->n() * sizeof(float), cudaMemcpyDeviceToHost));
    return RST_OK;
}

// Debug: activation (want_grad = 0) or its gradient (want_grad = 1) of the last step, by name: "<layer>/conv" (after the
// convolution's own activation), "<layer>/out" (after normalisation, activation and skip), "style_params".
// Note the backward pass masks gradients in place, so a "/conv" gradient is the one w.r.t. the pre-activation.
extern "C" int rst_train_debug_read(rst_trainer* t, const char* name, int want_grad, float* h_out, int64_t capacity, int64_t* elems) {
    if (!t || !name) return RST_ERR_INVALID;
    for (auto& x : t->tensors) {
        if (x.name != name) continue;
        if (elems) *elems = x.n();
        if (!h_out) return RST_OK;
        if (capacity < x.n()) return tfail(t, RST_ERR_INVALID, "rst_train_debug_read: buffer too small");
        cudaSetDevice(t->device);
        TCUDA(t, cudaMemcpy(h_out, want_grad ? x.g : x.d, (size_t)x.n() * sizeof(float), cudaMemcpyDeviceToHost));
        return RST_OK;
    }
    return tfail(t, RST_ERR_INVALID, std::string("rst_train_debug_read: no tensor named ") + name);
}
