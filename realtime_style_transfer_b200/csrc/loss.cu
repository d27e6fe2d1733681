// Training loss of the reference: VGG16 feature extraction, Gram-matrix style loss, content loss and total variation
// (models/styleLoss.py:69-109 StyleLossModelVGG, :11-18 Gram, :290-292 mean_l2_loss_on_batch, :295-369
// make_style_loss_function with with_depth_loss=False), forward and backward w.r.t. the prediction.
//
// fp32 CUDA-core kernels (the bar on the Gram-loss scalar is 1e-3 relative, which bf16 features do not meet).
// The loss is a (B,) vector; Keras differentiates its batch SUM (styleTransferTrainingModel.py:26-29), so the backward
// here returns d(sum_b loss[b]) / d(prediction).
#include <cmath>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "rst_internal.cuh"
#include "halo_gemm.cuh"
#include "train_kernels.cuh"
#include <memory>

using namespace rst;

namespace {

struct VggLayer { std::string name; int ci, co; bool pool_after; int style_idx; bool content; };
const VggLayer kVgg[13] = {
    {"block1_conv1", 3, 64, false, -1, false},   {"block1_conv2", 64, 64, true, 0, false},
    {"block2_conv1", 64, 128, false, -1, false}, {"block2_conv2", 128, 128, true, 1, false},
    {"block3_conv1", 128, 256, false, -1, false}, {"block3_conv2", 256, 256, false, -1, false},
    {"block3_conv3", 256, 256, true, 2, false},
    {"block4_conv1", 256, 512, false, -1, false}, {"block4_conv2", 512, 512, false, -1, false},
    {"block4_conv3", 512, 512, true, 3, false},
    {"block5_conv1", 512, 512, false, -1, false}, {"block5_conv2", 512, 512, false, -1, false},
    {"block5_conv3", 512, 512, false, -1, true},
};

// ---- small kernels ---------------------------------------------------------------------------------------------
// x*255 -> RGB->BGR -> subtract the caffe mean (keras vgg16.preprocess_input, 'caffe' mode)
__global__ void vgg_preprocess_kernel(const float* __restrict__ x, float* __restrict__ y, long long pixels) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= pixels) return;
    const float r = x[i * 3] * 255.f, g = x[i * 3 + 1] * 255.f, b = x[i * 3 + 2] * 255.f;
    y[i * 3] = b - 103.939f; y[i * 3 + 1] = g - 116.779f; y[i * 3 + 2] = r - 123.68f;
}
// gradient of the preprocessing: dx[rgb] = 255 * dy[bgr]
__global__ void vgg_preprocess_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, long long pixels, int accumulate) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= pixels) return;
    const float r = 255.f * dy[i * 3 + 2], g = 255.f * dy[i * 3 + 1], b = 255.f * dy[i * 3];
    if (accumulate) { dx[i * 3] += r; dx[i * 3 + 1] += g; dx[i * 3 + 2] += b; }
    else { dx[i * 3] = r; dx[i * 3 + 1] = g; dx[i * 3 + 2] = b; }
}
// g *= (out > 0)   (ReLU backward on the post-activation tensor)
__global__ void relu_bwd_kernel(float* __restrict__ g, const float* __restrict__ out, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && !(out[i] > 0.f)) g[i] = 0.f;
}
// 2x2 max-pool backward: the gradient goes to the first maximal element of each window (TF MaxPoolGrad)
__global__ void maxpool2_bwd_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ gy,
                                    float* __restrict__ gx, int Hi, int Wi, int C, long long total_out) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total_out) return;
    const int Ho = Hi / 2, Wo = Wi / 2;
    int c = (int)(idx % C);
    long long t = idx / C;
    int ox = (int)(t % Wo);
    t /= Wo;
    int oy = (int)(t % Ho);
    long long n = t / Ho;
    const long long base = ((n * Hi + 2 * oy) * Wi + 2 * ox) * C + c;
    const long long off[4] = {0, C, (long long)Wi * C, (long long)Wi * C + C};
    const float m = y[idx], g = gy[idx];
    bool done = false;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const bool hit = !done && x[base + off[k]] == m;
        gx[base + off[k]] = hit ? g : 0.f;
        done = done || hit;
    }
}
// per-sample sum of (a-b)^2 -> out[n] (double atomics); count elements per sample = per
__global__ void sqdiff_kernel(const float* __restrict__ a, const float* __restrict__ b, double* __restrict__ out, long long per) {
    const int n = blockIdx.y;
    double acc = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < per; i += (long long)gridDim.x * blockDim.x) {
        const double d = (double)a[n * per + i] - (double)b[n * per + i];
        acc += d * d;
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(out + n, acc);
}
// g (+)= scale * (a - b)
__global__ void diff_scale_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ g, float scale,
                                  long long n, int accumulate) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float v = scale * (a[i] - b[i]);
    g[i] = accumulate ? g[i] + v : v;
}
// tf.image.total_variation: per-sample sum |dy| + |dx| over (H,W,C)
__global__ void tv_kernel(const float* __restrict__ x, double* __restrict__ out, int H, int W, int C) {
    const int n = blockIdx.y;
    const long long per = (long long)H * W * C;
    const float* xb = x + n * per;
    double acc = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < per; i += (long long)gridDim.x * blockDim.x) {
        const long long pix = i / C;
        const int w = (int)(pix % W), h = (int)(pix / W);
        const float v = xb[i];
        if (h + 1 < H) acc += fabs((double)xb[i + (long long)W * C] - (double)v);
        if (w + 1 < W) acc += fabs((double)xb[i + C] - (double)v);
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(out + n, acc);
}
__device__ __forceinline__ float sgn(float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); }
// gradient of scale * total_variation, accumulated into g
__global__ void tv_bwd_kernel(const float* __restrict__ x, float* __restrict__ g, int H, int W, int C, float scale, long long total) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const long long pix = i / C;
    const int w = (int)(pix % W), h = (int)((pix / W) % H);
    const float v = x[i];
    float d = 0.f;
    if (h + 1 < H) d -= sgn(x[i + (long long)W * C] - v);
    if (h > 0) d += sgn(v - x[i - (long long)W * C]);
    if (w + 1 < W) d -= sgn(x[i + C] - v);
    if (w > 0) d += sgn(v - x[i - C]);
    g[i] += scale * d;
}
__global__ void finalize_losses_kernel(const double* __restrict__ feat, const double* __restrict__ style4, const double* __restrict__ tv,
                                       float* __restrict__ out, int B, double feat_norm, const double* style_norm4,
                                       double cf, double sf, double tf) {
    int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= B) return;
    const double f = 0.5 * feat[n] / feat_norm * cf;
    double s = 0.0;
    for (int l = 0; l < 4; ++l) s += 0.5 * style4[l * B + n] / style_norm4[l];
    s = s / 4.0 * sf;
    const double t = tv[n] * tf;
    out[n * 4 + 0] = (float)(f + s + t);
    out[n * 4 + 1] = (float)f;
    out[n * 4 + 2] = (float)s;
    out[n * 4 + 3] = (float)t;
}

inline unsigned blocks_for(long long n) { return (unsigned)((n + 255) / 256); }

}  // namespace

struct rst_loss {
    int device = 0, H = 0, W = 0, max_batch = 0;
    mutable std::string err;
    bool committed = false;
    float cf = 1e4f, sf = 1e-3f, tf = 1e-1f;                 // styleLoss.py:101-104
    std::map<std::string, std::vector<float>> host_w;
    std::map<std::string, float*> dev_w;
    // per-layer geometry
    int lh[13], lw[13];
    // saved activations of the PREDICTION pass: input (preprocessed), conv outputs, pool outputs
    float* pre = nullptr;
    float* act[13] = {};
    float* pool[13] = {};
    // scratch for the content / style passes and for gradients
    float* sa = nullptr; float* sb = nullptr;
    float* content_feat = nullptr;               // block5_conv3 of the ground-truth content
    float* gram_style[4] = {}; float* gram_pred[4] = {}; float* gram_diff = nullptr;
    double* red = nullptr;                       // [feat B | style 4B | tv B]
    double* style_norm_dev = nullptr;
    int last_batch = 0;
    int64_t launches = 0;
    // tf32 tensor-core convolutions for the 12 layers with >= 64 input channels (conv_tf32.cu), forward and input gradient
    int math = RST_PRECISION_FP32;       // FP32: split tf32 on the tensor cores (fp32-level accuracy); TF32: plain tf32 operands
    float* split_scratch = nullptr;      // [x_hi | x_lo | x_hi] expansion of the largest conv input
    size_t split_scratch_floats = 0;
    int num_sms = 148;
    std::unique_ptr<Tf32Conv3x3> fwd[13], bwd[13];
};

static thread_local std::string g_loss_err;
static int lfail(rst_loss* c, int code, const std::string& m) { if (c) c->err = m; else g_loss_err = m; return code; }
#define LCUDA(c, expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) return lfail((c), RST_ERR_CUDA, std::string(#expr ": ") + cudaGetErrorString(_e)); } while (0)

extern "C" const char* rst_loss_last_error(const rst_loss* c) { return c ? c->err.c_str() : g_loss_err.c_str(); }

extern "C" int rst_loss_destroy(rst_loss* c) {
    if (!c) return RST_OK;
    cudaSetDevice(c->device);
    for (auto& kv : c->dev_w) if (kv.second) cudaFree(kv.second);
    for (float* p : {c->pre, c->sa, c->sb, c->content_feat, c->gram_diff}) if (p) cudaFree(p);
    for (int i = 0; i < 13; ++i) { if (c->act[i]) cudaFree(c->act[i]); if (c->pool[i]) cudaFree(c->pool[i]); }
    for (int i = 0; i < 4; ++i) { if (c->gram_style[i]) cudaFree(c->gram_style[i]); if (c->gram_pred[i]) cudaFree(c->gram_pred[i]); }
    if (c->red) cudaFree(c->red);
    if (c->style_norm_dev) cudaFree(c->style_norm_dev);
    if (c->split_scratch) cudaFree(c->split_scratch);
    delete c;
    return RST_OK;
}

// StyleLossModelVGG(input_shape) + make_style_loss_function(loss_model, output_shape, num_styles=1, with_depth_loss=False)
extern "C" int rst_loss_create(int h, int w, int max_batch, int device, rst_loss** out) {
    if (!out || h < 16 || w < 16 || h % 16 || w % 16 || max_batch < 1)
        return lfail(nullptr, RST_ERR_INVALID, "rst_loss_create: image size must be a positive multiple of 16");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return lfail(nullptr, RST_ERR_CUDA, "rst_loss_create: no CUDA device is visible; this library has no CPU fallback");
    if (device < 0 || device >= ndev) return lfail(nullptr, RST_ERR_INVALID, "rst_loss_create: bad device");
    cudaSetDevice(device);
    rst_loss* c = new rst_loss();
    c->device = device; c->H = h; c->W = w; c->max_batch = max_batch;
    const size_t B = max_batch;
    bool ok = true;
    auto alloc = [&](float** p, size_t elems) { if (ok && cudaMalloc(p, elems * sizeof(float)) != cudaSuccess) ok = false; };
    int ch = h, cw = w;
    size_t max_act = 0;
    for (int i = 0; i < 13; ++i) {
        c->lh[i] = ch; c->lw[i] = cw;
        const size_t e = B * ch * cw * kVgg[i].co;
        max_act = std::max(max_act, e);
        alloc(&c->act[i], e);
        if (kVgg[i].pool_after) { ch /= 2; cw /= 2; alloc(&c->pool[i], B * ch * cw * kVgg[i].co); }
    }
    alloc(&c->pre, B * h * w * 3);
    alloc(&c->sa, max_act);
    alloc(&c->sb, max_act);
    alloc(&c->content_feat, B * c->lh[12] * c->lw[12] * 512);
    const int sc[4] = {64, 128, 256, 512};
    for (int l = 0; l < 4; ++l) { alloc(&c->gram_style[l], B * sc[l] * sc[l]); alloc(&c->gram_pred[l], B * sc[l] * sc[l]); }
    alloc(&c->gram_diff, 512 * 512);
    if (ok && cudaMalloc(&c->red, 6 * B * sizeof(double)) != cudaSuccess) ok = false;
    if (ok && cudaMalloc(&c->style_norm_dev, 4 * sizeof(double)) != cudaSuccess) ok = false;
    if (!ok) { rst_loss_destroy(c); return lfail(nullptr, RST_ERR_CUDA, "rst_loss_create: out of device memory"); }
    double sn[4];
    for (int l = 0; l < 4; ++l) sn[l] = (double)sc[l] * sc[l];
    cudaMemcpy(c->style_norm_dev, sn, sizeof(sn), cudaMemcpyHostToDevice);
    *out = c;
    return RST_OK;
}

extern "C" int rst_loss_set_factors(rst_loss* c, float content, float style, float tv) {
    if (!c) return RST_ERR_INVALID;
    c->cf = content; c->sf = style; c->tf = tv;
    return RST_OK;
}

extern "C" int rst_loss_set_weight(rst_loss* c, const char* name, const float* h_data, const int64_t* shape, int ndim) {
    if (!c || !name || !h_data || !shape) return lfail(c, RST_ERR_INVALID, "rst_loss_set_weight: null argument");
    std::string n(name);
    for (int i = 0; i < 13; ++i) {
        if (n == kVgg[i].name + "/kernel") {
            if (ndim != 4 || shape[0] != 3 || shape[1] != 3 || shape[2] != kVgg[i].ci || shape[3] != kVgg[i].co)
                return lfail(c, RST_ERR_INVALID, "rst_loss_set_weight: shape mismatch for " + n);
            c->host_w[n].assign(h_data, h_data + (size_t)9 * kVgg[i].ci * kVgg[i].co);
            c->committed = false;
            return RST_OK;
        }
        if (n == kVgg[i].name + "/bias") {
            if (ndim != 1 || shape[0] != kVgg[i].co) return lfail(c, RST_ERR_INVALID, "rst_loss_set_weight: shape mismatch for " + n);
            c->host_w[n].assign(h_data, h_data + kVgg[i].co);
            c->committed = false;
            return RST_OK;
        }
    }
    return lfail(c, RST_ERR_INVALID, "rst_loss_set_weight: unknown variable " + n);
}

extern "C" int rst_loss_commit(rst_loss* c) {
    if (!c) return RST_ERR_INVALID;
    cudaSetDevice(c->device);
    for (int i = 0; i < 13; ++i)
        for (const char* suf : {"/kernel", "/bias"}) {
            const std::string n = kVgg[i].name + suf;
            auto it = c->host_w.find(n);
            if (it == c->host_w.end()) return lfail(c, RST_ERR_STATE, "rst_loss_commit: variable not set: " + n);
            float*& d = c->dev_w[n];
            if (!d) LCUDA(c, cudaMalloc(&d, it->second.size() * sizeof(float)));
            LCUDA(c, cudaMemcpy(d, it->second.data(), it->second.size() * sizeof(float), cudaMemcpyHostToDevice));
        }
    for (int i = 0; i < 13; ++i) { c->fwd[i].reset(); c->bwd[i].reset(); }
    const bool cuda_core_only = c->math == RST_PRECISION_FP32 && ab_env("RST_LOSS_CUDA_CORE") != nullptr;
    if (!cuda_core_only) {
        const bool split = c->math == RST_PRECISION_FP32;
        cudaDeviceGetAttribute(&c->num_sms, cudaDevAttrMultiProcessorCount, c->device);
        std::string err;
        size_t need = 0;
        for (int i = 1; i < 13; ++i) {
            const std::vector<float>& k = c->host_w[kVgg[i].name + "/kernel"];
            const std::vector<float>& b = c->host_w[kVgg[i].name + "/bias"];
            c->fwd[i].reset(new Tf32Conv3x3());
            c->bwd[i].reset(new Tf32Conv3x3());
            if (!c->fwd[i]->setup(kVgg[i].ci, kVgg[i].co, k.data(), b.data(), true, false, &err, split) ||
                !c->bwd[i]->setup(kVgg[i].ci, kVgg[i].co, k.data(), nullptr, false, true, &err, split))
                return lfail(c, RST_ERR_CUDA, "rst_loss_commit: " + err);
            need = std::max(need, c->fwd[i]->scratch_floats(c->max_batch, c->lh[i], c->lw[i]));
            need = std::max(need, c->bwd[i]->scratch_floats(c->max_batch, c->lh[i], c->lw[i]));
        }
        if (need > c->split_scratch_floats) {
            if (c->split_scratch) cudaFree(c->split_scratch);
            c->split_scratch = nullptr; c->split_scratch_floats = 0;
            LCUDA(c, cudaMalloc(&c->split_scratch, need * sizeof(float)));
            c->split_scratch_floats = need;
        }
    }
    c->committed = true;
    return RST_OK;
}

// The 12 convolutions with >= 64 input channels and their input gradients run on the tensor cores.  RST_PRECISION_FP32 (default):
// error-compensated split tf32 (x_hi*w_hi + x_lo*w_hi + x_hi*w_lo, fp32 accumulation) = fp32-level accuracy.
// RST_PRECISION_TF32: plain tf32 operands -- what TensorFlow does with float32 convolutions on Ampere-and-later GPUs unless
// tf.config.experimental.enable_tensor_float_32_execution(False); ~3x less tensor work, losses still within 1e-3.
// (Environment RST_LOSS_CUDA_CORE=1 with FP32 math: CUDA-core fp32 convolutions everywhere, the slow cross-check.)
extern "C" int rst_loss_set_math(rst_loss* c, int precision) {
    if (!c) return RST_ERR_INVALID;
    if (precision != RST_PRECISION_FP32 && precision != RST_PRECISION_TF32)
        return lfail(c, RST_ERR_INVALID, "rst_loss_set_math: RST_PRECISION_FP32 or RST_PRECISION_TF32");
    if (precision != c->math) { c->math = precision; c->committed = false; }
    return RST_OK;
}

static ConvF32 vgg_conv(rst_loss* c, int i, const float* x, float* y, int batch) {
    ConvF32 p;
    p.x = x; p.y = y; p.w = c->dev_w[kVgg[i].name + "/kernel"]; p.bias = c->dev_w[kVgg[i].name + "/bias"];
    p.B = batch; p.Hi = c->lh[i]; p.Wi = c->lw[i]; p.Ci = kVgg[i].ci; p.Ho = c->lh[i]; p.Wo = c->lw[i]; p.Co = kVgg[i].co;
    p.kh = 3; p.kw = 3; p.stride = 1; p.pad_t = 1; p.pad_l = 1;
    p.w_tap = (long long)kVgg[i].ci * kVgg[i].co; p.w_ci = kVgg[i].co; p.w_co = 1;
    p.act1 = ACT_RELU;
    return p;
}

// Gram matrix of one style tap (styleLoss.py:11-18).  On the tensor cores (gram_tf32.cu) whenever the convolutions are: split
// tf32 with RST_PRECISION_FP32 math, plain tf32 operands with RST_PRECISION_TF32; the scratch of the split convolutions is free
// between two layers and large enough (2 x the largest conv input >= the tap).
static int gram(rst_loss* c, const float* y, float* out, int B, int P, int C, cudaStream_t s) {
    const bool split = c->math == RST_PRECISION_FP32;
    if (c->fwd[1] && gram_tf32_supported(C) && (!split || gram_tf32_scratch_floats(B, P, C) <= c->split_scratch_floats)) {
        std::string err;
        cudaError_t e = launch_gram_tf32(y, out, c->split_scratch, B, P, C, split, c->num_sms, s, &err);
        if (e != cudaSuccess) return lfail(c, RST_ERR_CUDA, "gram (tensor cores): " + (err.empty() ? std::string(cudaGetErrorString(e)) : err));
        return RST_OK;
    }
    LCUDA(c, launch_gram_f32(y, out, B, P, C, s));
    return RST_OK;
}

// Runs VGG16 on `img` (B,H,W,3 in [0,1]).  keep: write every activation into c->act/pool (prediction pass);
// otherwise ping-pong through the scratch buffers and only hand the tap tensors to `on_tap`.
template <typename F>
static int vgg_forward(rst_loss* c, const float* img, int batch, bool keep, cudaStream_t s, F on_tap) {
    const long long px = (long long)batch * c->H * c->W;
    float* pre = keep ? c->pre : c->sb;
    vgg_preprocess_kernel<<<blocks_for(px), 256, 0, s>>>(img, pre, px);
    c->launches++;
    const float* cur = pre;
    int flip = 0;
    for (int i = 0; i < 13; ++i) {
        float* y = keep ? c->act[i] : (flip ? c->sb : c->sa);
        if (c->fwd[i] && !exp_env("RST_EXP_LOSS_FWD_FP32")) {
            std::string err;
            cudaError_t e = c->fwd[i]->run_split(cur, c->split_scratch, y, batch, c->lh[i], c->lw[i], c->num_sms, s, &err);
            if (e != cudaSuccess) return lfail(c, RST_ERR_CUDA, "tf32 conv " + kVgg[i].name + ": " + (err.empty() ? cudaGetErrorString(e) : err));
            c->launches += c->fwd[i]->nblk;
        } else {
            ConvF32 p = vgg_conv(c, i, cur, y, batch);
            p.out_tf32 = c->math == RST_PRECISION_TF32 ? 1 : 0;
            LCUDA(c, launch_conv_f32(p, s));
            c->launches++;
        }
        int rc = on_tap(i, y);
        if (rc) return rc;
        cur = y;
        flip ^= 1;
        if (kVgg[i].pool_after) {
            float* q = keep ? c->pool[i] : (flip ? c->sb : c->sa);
            LCUDA(c, launch_maxpool2_f32(cur, q, batch, c->lh[i], c->lw[i], kVgg[i].co, s));
            c->launches++;
            cur = q;
            flip ^= 1;
        }
    }
    return RST_OK;
}

// compute_loss(y_pred, y_true) -> (B,4) [loss, feature_loss, style_loss, total_variation_loss]   (styleLoss.py:316-352)
extern "C" int rst_loss_forward(rst_loss* c, const float* d_pred, const float* d_gt_content, const float* d_gt_style,
                                float* d_losses, int batch, void* stream) {
    if (!c) return RST_ERR_INVALID;
    if (!c->committed) return lfail(c, RST_ERR_STATE, "rst_loss_forward: weights not committed");
    if (!d_pred || !d_gt_content || !d_gt_style || !d_losses) return lfail(c, RST_ERR_INVALID, "rst_loss_forward: null tensor");
    if (batch < 1 || batch > c->max_batch) return lfail(c, RST_ERR_INVALID, "rst_loss_forward: batch out of range");
    cudaSetDevice(c->device);
    cudaStream_t s = (cudaStream_t)stream;
    c->launches = 0;
    const int B = batch;
    double* feat = c->red; double* style4 = c->red + c->max_batch; double* tv = c->red + 5 * c->max_batch;
    LCUDA(c, cudaMemsetAsync(c->red, 0, 6 * c->max_batch * sizeof(double), s));
    // ground-truth content: only block5_conv3 is needed
    const long long feat_elems = (long long)B * c->lh[12] * c->lw[12] * 512;
    int rc = vgg_forward(c, d_gt_content, B, false, s, [&](int i, float* y) -> int {
        if (kVgg[i].content) LCUDA(c, cudaMemcpyAsync(c->content_feat, y, feat_elems * sizeof(float), cudaMemcpyDeviceToDevice, s));
        return RST_OK;
    });
    if (rc) return rc;
    // style image: the four Gram matrices
    rc = vgg_forward(c, d_gt_style, B, false, s, [&](int i, float* y) -> int {
        if (kVgg[i].style_idx >= 0) {
            if (int grc = gram(c, y, c->gram_style[kVgg[i].style_idx], B, c->lh[i] * c->lw[i], kVgg[i].co, s)) return grc;
            c->launches++;
        }
        return RST_OK;
    });
    if (rc) return rc;
    // prediction: keep everything for the backward pass
    rc = vgg_forward(c, d_pred, B, true, s, [&](int i, float* y) -> int {
        if (kVgg[i].style_idx >= 0) {
            const int l = kVgg[i].style_idx, C = kVgg[i].co;
            if (int grc = gram(c, y, c->gram_pred[l], B, c->lh[i] * c->lw[i], C, s)) return grc;
            sqdiff_kernel<<<dim3(64, B), 256, 0, s>>>(c->gram_pred[l], c->gram_style[l], style4 + (size_t)l * B, (long long)C * C);
            c->launches += 2;
        }
        if (kVgg[i].content) {
            sqdiff_kernel<<<dim3(256, B), 256, 0, s>>>(y, c->content_feat, feat, feat_elems / B);
            c->launches++;
        }
        return RST_OK;
    });
    if (rc) return rc;
    tv_kernel<<<dim3(256, B), 256, 0, s>>>(d_pred, tv, c->H, c->W, 3);
    finalize_losses_kernel<<<1, 64, 0, s>>>(feat, style4, tv, d_losses, B, (double)(feat_elems / B), c->style_norm_dev,
                                            (double)c->cf, (double)c->sf, (double)c->tf);
    c->launches += 2;
    LCUDA(c, cudaGetLastError());
    c->last_batch = B;
    return RST_OK;
}

// d(sum_b loss[b]) / d(prediction) using the activations saved by the last rst_loss_forward on the same prediction.
extern "C" int rst_loss_backward(rst_loss* c, const float* d_pred, float* d_grad_pred, int batch, void* stream) {
    if (!c) return RST_ERR_INVALID;
    if (!c->committed || c->last_batch != batch) return lfail(c, RST_ERR_STATE, "rst_loss_backward: call rst_loss_forward on this batch first");
    if (!d_pred || !d_grad_pred) return lfail(c, RST_ERR_INVALID, "rst_loss_backward: null tensor");
    cudaSetDevice(c->device);
    cudaStream_t s = (cudaStream_t)stream;
    const int B = batch;
    float* g = c->sa;          // gradient w.r.t. the current layer's output
    float* gn = c->sb;         // gradient w.r.t. its input
    // seed: content loss at block5_conv3: cf * (Fp - Fc) / (H5*W5*C5)
    {
        const long long n = (long long)B * c->lh[12] * c->lw[12] * 512;
        diff_scale_kernel<<<blocks_for(n), 256, 0, s>>>(c->act[12], c->content_feat, g, c->cf / (float)(n / B), n, 0);
    }
    for (int i = 12; i >= 0; --i) {
        const int C = kVgg[i].co, P = c->lh[i] * c->lw[i];
        const long long n = (long long)B * P * C;
        if (kVgg[i].pool_after) {
            // g currently holds the gradient w.r.t. the pooled tensor -> route it to the conv output
            const long long nout = (long long)B * (c->lh[i] / 2) * (c->lw[i] / 2) * C;
            maxpool2_bwd_kernel<<<blocks_for(nout), 256, 0, s>>>(c->act[i], c->pool[i], g, gn, c->lh[i], c->lw[i], C, nout);
            std::swap(g, gn);
        }
        if (kVgg[i].style_idx >= 0) {
            // style loss: d/dF of sf/4 * mean_{c,d} 1/2 (G-Gs)^2, G = F^T F / P  ->  (sf / (4 C^2)) * (2/P) * F (G - Gs)
            const int l = kVgg[i].style_idx;
            for (int b = 0; b < B; ++b) {
                diff_scale_kernel<<<blocks_for((long long)C * C), 256, 0, s>>>(c->gram_pred[l] + (size_t)b * C * C,
                                                                               c->gram_style[l] + (size_t)b * C * C, c->gram_diff,
                                                                               c->sf / (4.f * C * C) * 2.f / (float)P, (long long)C * C, 0);
                ConvF32 p;     // (P x C) . (C x C) as a 1x1 convolution, accumulated into g through `residual`
                p.x = c->act[i] + (size_t)b * P * C; p.y = g + (size_t)b * P * C; p.w = c->gram_diff;
                p.residual = g + (size_t)b * P * C;
                p.B = 1; p.Hi = c->lh[i]; p.Wi = c->lw[i]; p.Ci = C; p.Ho = c->lh[i]; p.Wo = c->lw[i]; p.Co = C;
                p.w_tap = 0; p.w_ci = C; p.w_co = 1;
                LCUDA(c, launch_conv_f32(p, s));
            }
        }
        const bool tc_bwd = c->bwd[i] && !exp_env("RST_EXP_LOSS_BWD_FP32");      // (env: experiment switch, see profiles/r01_03_experiments.md)
        const bool fused_mask = tc_bwd && c->bwd[i]->split;                      // the split expansion applies the ReLU mask itself
        if (!fused_mask) relu_bwd_kernel<<<blocks_for(n), 256, 0, s>>>(g, c->act[i], n);
        if (tc_bwd) {
            std::string err;
            cudaError_t e = c->bwd[i]->run_split(g, c->split_scratch, gn, B, c->lh[i], c->lw[i], c->num_sms, s, &err,
                                                 fused_mask ? c->act[i] : nullptr);
            if (e != cudaSuccess) return lfail(c, RST_ERR_CUDA, "tf32 dgrad " + kVgg[i].name + ": " + (err.empty() ? cudaGetErrorString(e) : err));
            std::swap(g, gn);
            continue;
        }
        // dgrad of the 3x3 SAME conv: a "transposed" pass over g with the same kernel, input/output channel roles swapped
        ConvF32 p;
        p.x = g; p.y = gn; p.w = c->dev_w[kVgg[i].name + "/kernel"];
        p.B = B; p.Hi = c->lh[i]; p.Wi = c->lw[i]; p.Ci = C; p.Ho = c->lh[i]; p.Wo = c->lw[i]; p.Co = kVgg[i].ci;
        p.kh = 3; p.kw = 3; p.stride = 1; p.pad_t = 1; p.pad_l = 1; p.transposed = 1;
        p.w_tap = (long long)kVgg[i].ci * C; p.w_ci = 1; p.w_co = C;
        LCUDA(c, launch_conv_f32(p, s));
        std::swap(g, gn);
    }
    const long long px = (long long)B * c->H * c->W;
    vgg_preprocess_bwd_kernel<<<blocks_for(px), 256, 0, s>>>(g, d_grad_pred, px, 0);
    tv_bwd_kernel<<<blocks_for(px * 3), 256, 0, s>>>(d_pred, d_grad_pred, c->H, c->W, 3, c->tf, px * 3);
    LCUDA(c, cudaGetLastError());
    return RST_OK;
}
