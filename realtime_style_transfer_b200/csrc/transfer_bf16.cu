// RST_PRECISION_BF16 forward of the transfer network: residual bottleneck on tcgen05 tensor cores
// (conv_umma.cu), instance-norm statistics fused into the conv epilogues, normalisation + style affine +
// ReLU / skip-add as one bf16 pass per layer.  Encoder / decoder layers still run the fp32 kernels.
#include "conv_umma.cuh"
#include "rst_ctx.h"

namespace rst {

struct TrunkLayer {
    __nv_bfloat16* w_packed = nullptr;   // [half][tap][COUT][64] bf16, K-major
    CUtensorMap tmB;
    int nhalf = 1;
};

struct Bf16State {
    int num_sms = 148;
    int cin_pad = 64;                        // channels of the (zero-padded) bottleneck input
    __nv_bfloat16 *b_in = nullptr, *bx = nullptr, *by = nullptr, *bz = nullptr;
    CUtensorMap tm_in, tm_x, tm_z;
    double* stats = nullptr;                 // [10][max_batch][F][2]
    size_t stats_bytes = 0;
    TrunkLayer layers[10];
    ~Bf16State() {
        for (void* p : {(void*)b_in, (void*)bx, (void*)by, (void*)bz, (void*)stats}) if (p) cudaFree(p);
        for (auto& l : layers) if (l.w_packed) cudaFree(l.w_packed);
    }
};
struct TrainState {};

int bf16_create(rst_ctx* c) {
    const int F = c->cfg.bottleneck_num_filters;
    if (F != 64 && F != 128)
        return fail(c, RST_ERR_UNSUPPORTED, "bf16 path: bottleneck_num_filters must be 64 or 128 (use the fp32 path otherwise)");
    std::string err;
    if (!umma_init(&err)) return fail(c, RST_ERR_CUDA, err);
    auto st = std::make_shared<Bf16State>();
    cudaDeviceProp prop;
    RST_CUDA(c, cudaGetDeviceProperties(&prop, c->device));
    st->num_sms = prop.multiProcessorCount;
    const int B = c->cfg.max_batch, H = c->bott_h, W = c->bott_w;
    const int res_in = c->residual[0].ci;
    st->cin_pad = (res_in + 63) / 64 * 64;
    const int fpad = F;
    const size_t px = (size_t)B * H * W;
    RST_CUDA(c, cudaMalloc(&st->b_in, px * st->cin_pad * 2));
    RST_CUDA(c, cudaMalloc(&st->bx, px * fpad * 2));
    RST_CUDA(c, cudaMalloc(&st->by, px * F * 2));
    RST_CUDA(c, cudaMalloc(&st->bz, px * fpad * 2));
    RST_CUDA(c, cudaMemset(st->bx, 0, px * fpad * 2));
    RST_CUDA(c, cudaMemset(st->bz, 0, px * fpad * 2));
    st->stats_bytes = (size_t)10 * B * F * 2 * sizeof(double);
    RST_CUDA(c, cudaMalloc(&st->stats, st->stats_bytes));
    if (!umma_encode_activation_map(&st->tm_in, st->b_in, B, H, W, st->cin_pad, &err)) return fail(c, RST_ERR_CUDA, err);
    if (!umma_encode_activation_map(&st->tm_x, st->bx, B, H, W, fpad, &err)) return fail(c, RST_ERR_CUDA, err);
    if (!umma_encode_activation_map(&st->tm_z, st->bz, B, H, W, fpad, &err)) return fail(c, RST_ERR_CUDA, err);
    c->bf16 = st;
    return RST_OK;
}

// Keras Conv2D kernel (3,3,ci,co) fp32 -> [half][tap][co][64] bf16 (zero padded input channels)
static void pack_conv3x3(const std::vector<float>& k, int ci, int co, int nhalf, std::vector<__nv_bfloat16>* out) {
    out->assign((size_t)nhalf * 9 * co * 64, __float2bfloat16(0.f));
    for (int tap = 0; tap < 9; ++tap)
        for (int i = 0; i < ci; ++i)
            for (int o = 0; o < co; ++o) {
                int half = i / 64, il = i % 64;
                (*out)[(((size_t)half * 9 + tap) * co + o) * 64 + il] = __float2bfloat16(k[((size_t)tap * ci + i) * co + o]);
            }
}

int bf16_commit(rst_ctx* c) {
    Bf16State* st = c->bf16.get();
    const int F = c->cfg.bottleneck_num_filters;
    std::string err;
    for (int i = 0; i < 10; ++i) {
        const LayerDesc& L = c->residual[i];
        const Weight* k = c->find_weight(L.name + "/kernel");
        TrunkLayer& tl = st->layers[i];
        tl.nhalf = (L.ci + 63) / 64;
        std::vector<__nv_bfloat16> packed;
        pack_conv3x3(k->host, L.ci, F, tl.nhalf, &packed);
        if (!tl.w_packed) RST_CUDA(c, cudaMalloc(&tl.w_packed, packed.size() * 2));
        RST_CUDA(c, cudaMemcpy(tl.w_packed, packed.data(), packed.size() * 2, cudaMemcpyHostToDevice));
        if (!umma_encode_weight_map(&tl.tmB, tl.w_packed, tl.nhalf * 9 * F, F, &err)) return fail(c, RST_ERR_CUDA, err);
    }
    return RST_OK;
}

static int trunk_conv(rst_ctx* c, int layer, const CUtensorMap& tmA, int batch, cudaStream_t s) {
    Bf16State* st = c->bf16.get();
    const int F = c->cfg.bottleneck_num_filters;
    ConvUmmaParams p;
    p.y = st->by;
    p.bias = c->wdev(c->residual[layer].name + "/bias");
    p.stats = st->stats + (size_t)layer * c->cfg.max_batch * F * 2;
    p.B = batch; p.H = c->bott_h; p.W = c->bott_w;
    p.nhalf = st->layers[layer].nhalf;
    p.tiles_h = ceil_div(p.H, kUmmaTH);
    p.tiles_w = ceil_div(p.W, kUmmaTW);
    p.relu = 1;
    LaunchScope ls(c, s, "conv3x3_umma");
    RST_CUDA(c, launch_conv3x3_umma(F, tmA, st->layers[layer].tmB, p, st->num_sms, s));
    return RST_OK;
}

static int trunk_norm(rst_ctx* c, int layer, __nv_bfloat16* dst, const __nv_bfloat16* residual, int batch,
                      const float* d_style_params, int param_off, int act, cudaStream_t s) {
    Bf16State* st = c->bf16.get();
    const int F = c->cfg.bottleneck_num_filters;
    CinApplyBf16 a;
    a.x = st->by; a.y = dst; a.residual = residual;
    a.stats = st->stats + (size_t)layer * c->cfg.max_batch * F * 2;
    a.params = d_style_params;
    a.param_bstride = (long long)c->cfg.num_styles * c->num_style_params;
    a.param_sstride = c->num_style_params;
    a.scale_off = param_off; a.bias_off = param_off + F;
    a.weights = c->cfg.num_styles == 2 ? mip_for_width(c, c->bott_w) : nullptr;
    a.B = batch; a.P = c->bott_h * c->bott_w; a.C = F; a.num_styles = c->cfg.num_styles; a.act = act; a.eps = 1e-5f;
    if (c->cfg.num_styles == 2 && !a.weights) return fail(c, RST_ERR_STATE, "no style-weight mip for the bottleneck");
    LaunchScope ls(c, s, "cin_apply_bf16");
    RST_CUDA(c, launch_cin_apply_bf16(a, s));
    return RST_OK;
}

int bf16_transfer_forward(rst_ctx* c, const float* d_content, const float* d_style_params, const float* d_style_weights,
                          float* d_out, int batch, cudaStream_t s) {
    Bf16State* st = c->bf16.get();
    const int F = c->cfg.bottleneck_num_filters;
    const long long px = (long long)batch * c->bott_h * c->bott_w;
    int rc = build_mips(c, d_style_weights, batch, s);
    if (rc) return rc;
    float* enc = nullptr;
    int free_idx = 0;
    rc = fp32_contract_stage(c, d_content, batch, s, &enc, &free_idx);
    if (rc) return rc;
    {
        LaunchScope ls(c, s, "convert");
        RST_CUDA(c, launch_f32_to_bf16_pad(enc, st->b_in, px, c->residual[0].ci, st->cin_pad, s));
    }
    RST_CUDA(c, cudaMemsetAsync(st->stats, 0, st->stats_bytes, s));
    int cursor = 0;
    for (int b = 0; b < 5; ++b) {                                     // residual_block, styleTransfer.py:144-185
        const std::string name = "residual_block_" + std::to_string(b);
        rc = trunk_conv(c, 2 * b, b == 0 ? st->tm_in : st->tm_x, batch, s);
        if (rc) return rc;
        record_tap(c, name + "/conv0/relu", st->by, px * F, true, s);
        rc = trunk_norm(c, 2 * b, st->bz, nullptr, batch, d_style_params, cursor, ACT_RELU, s);
        if (rc) return rc;
        record_tap(c, name + "/conv0/cin", st->bz, px * F, true, s);
        rc = trunk_conv(c, 2 * b + 1, st->tm_z, batch, s);
        if (rc) return rc;
        record_tap(c, name + "/conv1/relu", st->by, px * F, true, s);
        rc = trunk_norm(c, 2 * b + 1, st->bx, b == 0 ? nullptr : st->bx, batch, d_style_params, cursor + 2 * F, ACT_NONE, s);
        if (rc) return rc;
        record_tap(c, name, st->bx, px * F, true, s);
        cursor += 4 * F;
    }
    float* x = c->act[free_idx];
    float* t1 = c->act[2];
    {
        LaunchScope ls(c, s, "convert");
        RST_CUDA(c, launch_bf16_to_f32_slice(st->bx, x, px, F, F, s));
    }
    return fp32_expand_stage(c, x, t1, d_style_params, cursor, d_out, batch, s);
}

// ------------------------------------------------------------------------------------------------
// stand-alone operator (tests): 3x3 stride-1 'same' conv through the tensor-core kernel, fp32 tensors at the
// boundary (converted to bf16 on the device).  Not a hot path: allocates and frees its scratch.
// ------------------------------------------------------------------------------------------------
int op_conv2d_bf16(const float* d_x, const float* d_kernel, const float* d_bias, float* d_y, int batch, int h, int w, int ci,
                   int co, int kh, int kw, int stride, int transposed, int act, cudaStream_t s, std::string* err) {
    if (kh != 3 || kw != 3 || stride != 1 || transposed || (co != 32 && co != 64 && co != 128) ||
        (act != ACT_NONE && act != ACT_RELU)) {
        *err = "rst_op_conv2d(bf16): only 3x3 stride-1 convs with 32/64/128 filters run on the tensor-core kernel";
        return RST_ERR_UNSUPPORTED;
    }
    if (!umma_init(err)) return RST_ERR_CUDA;
    const int nhalf = (ci + 63) / 64, cpad = nhalf * 64;
    const long long px = (long long)batch * h * w;
    std::vector<float> hk((size_t)9 * ci * co);
    cudaError_t e = cudaMemcpy(hk.data(), d_kernel, hk.size() * 4, cudaMemcpyDeviceToHost);
    std::vector<__nv_bfloat16> packed;
    pack_conv3x3(hk, ci, co, nhalf, &packed);
    __nv_bfloat16 *xb = nullptr, *yb = nullptr, *wb = nullptr;
    if (e == cudaSuccess) e = cudaMalloc(&xb, px * cpad * 2);
    if (e == cudaSuccess) e = cudaMalloc(&yb, px * co * 2);
    if (e == cudaSuccess) e = cudaMalloc(&wb, packed.size() * 2);
    if (e == cudaSuccess) e = cudaMemcpy(wb, packed.data(), packed.size() * 2, cudaMemcpyHostToDevice);
    CUtensorMap tmA, tmB;
    int rc = RST_OK;
    if (e == cudaSuccess) {
        if (!umma_encode_activation_map(&tmA, xb, batch, h, w, cpad, err) ||
            !umma_encode_weight_map(&tmB, wb, nhalf * 9 * co, co, err))
            rc = RST_ERR_CUDA;
    }
    if (e == cudaSuccess && rc == RST_OK) {
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        e = launch_f32_to_bf16_pad(d_x, xb, px, ci, cpad, s);
        ConvUmmaParams p;
        p.y = yb; p.bias = d_bias; p.stats = nullptr; p.B = batch; p.H = h; p.W = w; p.nhalf = nhalf;
        p.tiles_h = ceil_div(h, kUmmaTH); p.tiles_w = ceil_div(w, kUmmaTW); p.relu = act == ACT_RELU;
        if (e == cudaSuccess) e = launch_conv3x3_umma(co, tmA, tmB, p, sms, s);
        if (e == cudaSuccess) e = launch_bf16_to_f32_slice(yb, d_y, px, co, co, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    }
    for (void* p : {(void*)xb, (void*)yb, (void*)wb}) if (p) cudaFree(p);
    if (e != cudaSuccess) { *err = std::string("rst_op_conv2d(bf16): ") + cudaGetErrorString(e); return RST_ERR_CUDA; }
    return rc;
}

}  // namespace rst
