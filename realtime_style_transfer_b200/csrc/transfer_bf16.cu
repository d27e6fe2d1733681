// RST_PRECISION_BF16 forward of the transfer network on tcgen05 tensor cores.
//
//   content fp32 --pack--> [16 real | 9-tap window of channel 16] bf16 rows
//     -> stem 9x9 (halo GEMM, bias+ReLU+BatchNorm+ReLU epilogue)
//     -> contract_0 / contract_1 (3x3 stride 2)                      [fp32 CUDA-core kernels for now]
//     -> 10 bottleneck convs (halo GEMM; bias+ReLU+instance-norm statistics fused in the epilogue)
//        each followed by ONE bf16 pass: normalise + style affine (+ReLU | + skip add)
//     -> expand_0 / expand_1 (stride-2 transposed conv as a 2x2-tap GEMM over 4 output phases, stats fused)
//     -> expand_last 9x9 (4 pixels per GEMM row, 12 of 16 columns used) -> normalise + sigmoid -> fp32 image
#include <cstdlib>

#include "halo_gemm.cuh"
#include "rst_ctx.h"

namespace rst {

// ------------------------------------------------------------------------------------------------
// One convolution layer mapped onto the halo GEMM: packed weights, K-step table, tensor maps.
// ------------------------------------------------------------------------------------------------
struct HaloConv {
    HaloGemmLaunch launch;
    HaloGemmParams p;
    CUtensorMap tmA, tmB, tmB_half;
    CUtensorMap tmA_fuse;          // fused-norm mode: the RAW tensor the loader warps read (used for its L2 prefetch)
    bool stem_2cta = false;        // SCH_STEM2 / SCH_STEM2B: cta_group::2 kernel (halo_stem2cta.cu)
    bool two_cta = false;          // 128->128 3x3 convs: cta_group::2 kernel with resident weights
    __nv_bfloat16* w_packed = nullptr;
    float* col_bias = nullptr;     // [N] bias expanded to GEMM columns
    float* col_scale = nullptr;    // [N] optional post affine
    float* col_shift = nullptr;
    int total_ksteps = 0;          // n_groups * ksteps
    int in_C = 0;                  // channels (bf16 elements per row unit) of the bound input tensor

    ~HaloConv() {
        for (void* q : {(void*)w_packed, (void*)col_bias, (void*)col_scale, (void*)col_shift})
            if (q) cudaFree(q);
    }

    cudaError_t upload(const std::vector<__nv_bfloat16>& packed, const std::vector<float>& bias,
                       const std::vector<float>* scale, const std::vector<float>* shift) {
        cudaError_t e = cudaSuccess;
        auto up = [&](void** dst, const void* src, size_t bytes) {
            if (e != cudaSuccess) return;
            if (!*dst) e = cudaMalloc(dst, bytes);
            if (e == cudaSuccess) e = cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice);
        };
        up((void**)&w_packed, packed.data(), packed.size() * 2);
        up((void**)&col_bias, bias.data(), bias.size() * 4);
        if (scale) up((void**)&col_scale, scale->data(), scale->size() * 4);
        if (shift) up((void**)&col_shift, shift->data(), shift->size() * 4);
        p.bias = col_bias;
        p.post_scale = scale ? col_scale : nullptr;
        p.post_shift = shift ? col_shift : nullptr;
        return e;
    }

    // Bind the input tensor (B, H, WRU, in_C) and finish the launch plan.
    // For SCH_S2D the bound tensor is the un-strided input (B, 2H, 2W, in_C/2); H, WRU are the OUTPUT grid.
    bool bind_input(const void* x, int B, int H, int WRU, std::string* err) {
        p.B = B; p.H = H; p.WRU = WRU;
        p.tiles_h = ceil_div(H, 8); p.tiles_w = ceil_div(WRU, 16);
        if (launch.sched == SCH_S2D) {
            if (!encode_s2d_map(&tmA, x, B, 2 * H, 2 * WRU, in_C / 2, err)) return false;
        } else if (!encode_halo_map(&tmA, x, B, H, WRU * sched_a_unit_stride(launch.sched), in_C, launch.row_bytes / 2,
                                    sched_halo_h(launch.sched), sched_halo_w(launch.sched), err)) return false;
        if (launch.sched == SCH_HEAD8) {
            if (!encode_weight_unit_map(&tmB, w_packed, kHead8Boxes * 256, err)) return false;
        } else if (sched_b_units(launch.sched)) {
            if (!encode_weight_unit_map(&tmB, w_packed, (sched_b_boxes(launch.sched) + 1) * 256, err)) return false;
            const char* env2 = ab_env("RST_STEM_2CTA");
            stem_2cta = !(env2 && env2[0] == '0') && launch.mode == (HALO_MODE_RELU | HALO_MODE_POST);
        } else if (!encode_weight_map(&tmB, w_packed, total_ksteps / 4, launch.N, err)) return false;
        const char* env = ab_env("RST_TRUNK_2CTA");
        two_cta = launch.sched == SCH_C3 && launch.N == 128 && launch.row_bytes == 128 && launch.epi == EPI_NHWC &&
                  launch.mode == HALO_MODE_RELU && !(env && env[0] == '0') &&
                  halo_gemm2_smem_bytes(p.n_groups) <= 227 * 1024;
        if (two_cta && !encode_weight_map(&tmB_half, w_packed, total_ksteps / 4, launch.N, err, launch.N / 2)) return false;
        return halo_gemm_plan(&launch, &p, err);
    }

    int l2_hints = 0;               // HaloGemmParams::l2_hints (RST_L2_HINTS, read when the plan is built)
    bool next_is_conv = false;      // the kernel launched after this one in the forward is a convolution (HaloGemmParams::pdl_trigger)
    // fin (optional): fused input transform of the 2-CTA trunk kernel, fields fuse / fin_* of HaloGemmParams
    cudaError_t run(void* y, bool y_f32, double* stats, int batch, int num_sms, cudaStream_t s, const HaloGemmParams* fin = nullptr) {
        HaloGemmParams q = p;
        q.B = batch; q.y = y; q.y_f32 = y_f32 ? 1 : 0; q.stats = stats;
        q.pdl = 1;                      // inference: weights / bias / folded BN were packed at commit time, long before this launch
        q.pdl_trigger = next_is_conv ? 1 : 0;
        q.l2_hints = l2_hints;
        if (fin) {
            if (!two_cta || y_f32) return cudaErrorInvalidValue;
            q.fuse = fin->fuse; q.fin_x = fin->fin_x; q.fin_skip = fin->fin_skip; q.fin_out = fin->fin_out; q.fin_stats = fin->fin_stats;
            q.fin_params = fin->fin_params; q.fin_param_bstride = fin->fin_param_bstride;
            q.fin_scale_off = fin->fin_scale_off; q.fin_bias_off = fin->fin_bias_off;
#ifdef RST_EXPERIMENTS
            if (exp_env("RST_EXP_FUSE_DBG")) q.fuse_dbg = atoi(exp_env("RST_EXP_FUSE_DBG"));
#endif
            return launch_halo_gemm2(tmA_fuse, tmB_half, q, num_sms, s);
        }
#ifdef RST_EXPERIMENTS
        if (exp_env("RST_EXP_NOSTATS")) q.stats = nullptr;     // timing experiments only (wrong results)
        if (exp_env("RST_EXP_NOSTORE")) q.H = 0;
#endif
        if (two_cta && !y_f32) return launch_halo_gemm2(tmA, tmB_half, q, num_sms, s);
        if (stem_2cta && !y_f32 && !q.stats) return launch_halo_stem2cta(launch.sched, tmA, tmB, q, num_sms, s);
        return launch_halo_gemm(launch, tmA, tmB, q, num_sms, s);
    }
};

static void use_sched(HaloConv* c, int sch) {
    c->launch.sched = sch;
    c->p.a_w_mul = sched_a_unit_stride(sch);
    c->p.ksteps = sched_ksteps(sch, c->launch.row_bytes);
    c->p.halo_h = sched_halo_h(sch); c->p.halo_w = sched_halo_w(sch);
    c->p.oy = sched_oy(sch); c->p.ox = sched_ox(sch);
    c->total_ksteps = c->p.n_groups * c->p.ksteps;
}

// 3x3 stride-1 'same' conv, Keras kernel (3,3,ci,co).  ci <= 32 uses 64-byte rows, otherwise 64-channel groups.
static void setup_conv3x3(HaloConv* c, int ci, int co, const float* k, const float* bias, int act,
                          std::vector<__nv_bfloat16>* packed, std::vector<float>* col_bias) {
    const int row_elems = ci <= 32 ? 32 : 64, rowb = row_elems * 2;
    c->launch.N = co; c->launch.row_bytes = rowb; c->launch.epi = EPI_NHWC;
    c->in_C = (ci + row_elems - 1) / row_elems * row_elems;
    c->p.n_groups = c->in_C / row_elems;
    const int kper = row_elems / 16;
    use_sched(c, SCH_C3);
    const int real_ksteps = 9 * kper;
    c->launch.mode = HALO_MODE_RELU; (void)act;
    c->p.out_C = co; c->p.stats_c = co;
    const int ksteps = c->p.ksteps;
    pack_b_blocks(c->total_ksteps, co, [&](int ks, int n, int e) -> float {
        const int g = ks / ksteps, l = ks % ksteps;
        if (l >= real_ksteps) return 0.f;
        const int tap = l / kper, cin = g * row_elems + (l % kper) * 16 + e;
        return cin < ci ? k[((size_t)tap * ci + cin) * co + n] : 0.f;
    }, packed);
    col_bias->assign(bias, bias + co);
}

struct StemLayout { int n_real, n_virtual, row_elems; };
static bool stem_layout(int C, StemLayout* L) {
    if (C >= 16) { L->n_real = 16; L->n_virtual = C - 16; }
    else if (C <= 3) { L->n_real = 0; L->n_virtual = C; }
    else { L->n_real = C; L->n_virtual = 0; }
    const int elems = (L->n_real ? 16 : 0) + 16 * L->n_virtual;
    if (elems > 64) return false;
    L->row_elems = elems <= 32 ? 32 : 64;
    // schedules instantiated in halo_gemm.cu: (real, windowed) = (16,0) (16,1) (16,2) (0,3)
    const int key = (L->n_real ? 4 : 0) + L->n_virtual;
    return key == 4 || key == 5 || key == 6 || key == 3;
}

// 9x9 stride-1 'same' conv over packed rows [16 real | 16 per virtual channel], Keras kernel (9,9,C,co).
static bool setup_stem(HaloConv* c, int C, int co, const float* k, const float* bias, const float* bn_scale,
                       const float* bn_shift, std::vector<__nv_bfloat16>* packed, std::vector<float>* col_bias,
                       std::vector<float>* col_scale, std::vector<float>* col_shift) {
    StemLayout L;
    if (!stem_layout(C, &L)) return false;
    const int rowb = L.row_elems * 2;
    c->launch.N = co; c->launch.row_bytes = rowb; c->launch.epi = EPI_NHWC;
    c->in_C = L.row_elems; c->p.n_groups = 1;
    const int real_ksteps = L.n_real ? 81 : 0;
    use_sched(c, SCH_STEM + 4 * (L.n_real ? 1 : 0) + L.n_virtual);
    const int used = real_ksteps + 9 * L.n_virtual;
    c->launch.mode = HALO_MODE_RELU | HALO_MODE_POST;
    c->p.out_C = co; c->p.stats_c = co;
    pack_b_blocks(c->total_ksteps, co, [&](int ks, int n, int e) -> float {
        if (ks >= used) return 0.f;
        if (ks < real_ksteps) {
            return e < L.n_real ? k[((size_t)ks * C + e) * co + n] : 0.f;      // ks == ky*9 + kx
        }
        const int v = (ks - real_ksteps) / 9, ky = (ks - real_ksteps) % 9, ch = L.n_real + v;
        return e < 9 ? k[((size_t)(ky * 9 + e) * C + ch) * co + n] : 0.f;      // e == kx
    }, packed);
    col_bias->assign(bias, bias + co);
    col_scale->assign(bn_scale, bn_scale + co);
    col_shift->assign(bn_shift, bn_shift + co);
    return true;
}

// 17-channel 9x9 stem with two pixels per GEMM row (SCH_STEM2): input rows are pixel PAIRS of the packed tensor
// (B, H, W/2, 64), output is (B, H, W/2, 2 x 32) = the same bytes as NHWC-32.  Weights are stored as 32-row units.
static void setup_stem2(HaloConv* c, const float* k, const float* bias, const float* bn_scale, const float* bn_shift,
                        std::vector<__nv_bfloat16>* packed, std::vector<float>* col_bias, std::vector<float>* col_scale,
                        std::vector<float>* col_shift) {
    const int C = 17, co = 32;
    c->launch.N = 64; c->launch.row_bytes = 128; c->launch.epi = EPI_NHWC;
    c->launch.mode = HALO_MODE_RELU | HALO_MODE_POST;
    c->in_C = 64; c->p.n_groups = 1;
    use_sched(c, SCH_STEM2);
    c->p.out_C = 64; c->p.stats_c = 64;
    packed->assign((size_t)(kStem2Boxes + 1) * 256 * 16, __float2bfloat16(0.f));   // + one zero box: the 2-CTA peer reads one unit further
    auto put = [&](int unit, int row, int e, float w) { (*packed)[((size_t)unit * 32 + row) * 16 + e] = __float2bfloat16(w); };
    for (int ky = 0; ky < 9; ++ky) {
        for (int s = 1; s <= 9; ++s) {                       // unit s of the sequence holds column tap kx = 9 - s
            const int kx = 9 - s;
            for (int o = 0; o < co; ++o)
                for (int e = 0; e < 16; ++e) put(ky * 11 + s, o, e, k[((size_t)(ky * 9 + kx) * C + e) * co + o]);
        }
        for (int o = 0; o < co; ++o)                         // windowed 17th channel: window slot e holds column x0 - 4 + e
            for (int e = 0; e < 9; ++e) {
                const float w = k[((size_t)(ky * 9 + e) * C + 16) * co + o];
                put(99 + ky * 2, o, e, w);                   // even pixel x0: tap kx = e
                put(99 + ky * 2 + 1, o, e + 1, w);           // odd pixel x0 + 1: tap kx = e at slot e + 1
            }
    }
    col_bias->resize(64); col_scale->resize(64); col_shift->resize(64);
    for (int n = 0; n < 64; ++n) {
        (*col_bias)[n] = bias[n % co]; (*col_scale)[n] = bn_scale[n % co]; (*col_shift)[n] = bn_shift[n % co];
    }
}

// 18-channel variant (SCH_STEM2B): pair rows [even pixel 16 real | odd pixel 16 real | window ch 16 | window ch 17]; the unit
// sequences of consecutive row taps share their zero unit: [0, W(ky,8) .. W(ky,0)] x 9, one final 0, then 2 units per (ch, ky).
static void setup_stem2b(HaloConv* c, const float* k, const float* bias, const float* bn_scale, const float* bn_shift,
                         std::vector<__nv_bfloat16>* packed, std::vector<float>* col_bias, std::vector<float>* col_scale,
                         std::vector<float>* col_shift) {
    const int C = 18, co = 32;
    c->launch.N = 64; c->launch.row_bytes = 128; c->launch.epi = EPI_NHWC;
    c->launch.mode = HALO_MODE_RELU | HALO_MODE_POST;
    c->in_C = 64; c->p.n_groups = 1;
    use_sched(c, SCH_STEM2B);
    c->p.out_C = 64; c->p.stats_c = 64;
    packed->assign((size_t)(kStem2BBoxes + 1) * 256 * 16, __float2bfloat16(0.f));  // + one zero box for the 2-CTA peer's shifted copy
    auto put = [&](int unit, int row, int e, float w) { (*packed)[((size_t)unit * 32 + row) * 16 + e] = __float2bfloat16(w); };
    for (int ky = 0; ky < 9; ++ky) {
        for (int s = 1; s <= 9; ++s) {                       // unit s of the sequence holds column tap kx = 9 - s
            const int kx = 9 - s;
            for (int o = 0; o < co; ++o)
                for (int e = 0; e < 16; ++e) put(ky * 10 + s, o, e, k[((size_t)(ky * 9 + kx) * C + e) * co + o]);
        }
        for (int g = 0; g < 2; ++g)
            for (int o = 0; o < co; ++o)                     // window slot e holds column x0 - 4 + e of channel 16 + g
                for (int e = 0; e < 9; ++e) {
                    const float w = k[((size_t)(ky * 9 + e) * C + 16 + g) * co + o];
                    put(91 + (g * 9 + ky) * 2, o, e, w);             // even pixel x0: tap kx = e
                    put(91 + (g * 9 + ky) * 2 + 1, o, e + 1, w);     // odd pixel x0 + 1: tap kx = e at slot e + 1
                }
    }
    col_bias->resize(64); col_scale->resize(64); col_shift->resize(64);
    for (int n = 0; n < 64; ++n) {
        (*col_bias)[n] = bias[n % co]; (*col_scale)[n] = bn_scale[n % co]; (*col_shift)[n] = bn_shift[n % co];
    }
}

// 3x3 stride-2 'same' conv on even-sized inputs (TF pads 0 before / 1 after), Keras kernel (3,3,ci,co), followed by
// ReLU -> BatchNorm -> ReLU (contract blocks).  Runs over the space-to-depth view: a row holds [col parity 0 | col parity 1].
static bool setup_conv_s2(HaloConv* c, int ci, int co, const float* k, const float* bias, const float* bn_scale,
                          const float* bn_shift, std::vector<__nv_bfloat16>* packed, std::vector<float>* col_bias,
                          std::vector<float>* col_scale, std::vector<float>* col_shift) {
    if ((ci != 16 && ci != 32) || (co != 16 && co != 32)) return false;
    const int row_elems = 2 * ci, rowb = row_elems * 2;
    c->launch.N = co; c->launch.row_bytes = rowb; c->launch.epi = EPI_NHWC;
    c->launch.mode = HALO_MODE_RELU | HALO_MODE_POST;
    c->in_C = row_elems; c->p.n_groups = 1;
    use_sched(c, SCH_S2D);
    const int kper = rowb / 32, per_ky = kper + kper / 2, real = 3 * per_ky;
    c->p.out_C = co; c->p.stats_c = co;
    pack_b_blocks(c->total_ksteps, co, [&](int ks, int n, int e) -> float {
        if (ks >= real) return 0.f;
        const int ky = ks / per_ky, l = ks % per_ky;
        const int w2off = l >= kper ? 1 : 0, kk = l >= kper ? l - kper : l;
        const int cc = kk * 16 + e, pw = cc / ci, cin = cc % ci;
        const int kx = 2 * w2off + pw;
        if (kx > 2) return 0.f;
        return k[((size_t)(ky * 3 + kx) * ci + cin) * co + n];
    }, packed);
    col_bias->assign(bias, bias + co);
    col_scale->assign(bn_scale, bn_scale + co);
    col_shift->assign(bn_shift, bn_shift + co);
    return true;
}

// Conv2DTranspose 3x3 stride 2 'same' (Keras kernel (3,3,co,ci)): out[2i+a, 2j+b] over the 2x2 input window
// rows {i-1, i} x cols {j-1, j}; phase a uses ky = {d==1: a, d==0: a==0 ? 2 : none}.
static void setup_convt2(HaloConv* c, int ci, int co, const float* k, const float* bias,
                         std::vector<__nv_bfloat16>* packed, std::vector<float>* col_bias) {
    const int row_elems = ci <= 32 ? 32 : 64, rowb = row_elems * 2;
    c->launch.N = 4 * co; c->launch.row_bytes = rowb; c->launch.epi = EPI_CONVT2;
    c->in_C = (ci + row_elems - 1) / row_elems * row_elems;
    c->p.n_groups = c->in_C / row_elems;
    const int kper = row_elems / 16;
    use_sched(c, SCH_T2);
    c->launch.mode = 0;
    c->p.out_C = co; c->p.stats_c = co;
    const int ksteps = c->p.ksteps;
    auto tap_of = [](int phase_bit, int d) -> int { return d == 1 ? phase_bit : (phase_bit == 0 ? 2 : -1); };
    pack_b_blocks(c->total_ksteps, 4 * co, [&](int ks, int n, int e) -> float {
        const int g = ks / ksteps, l = ks % ksteps;
        if (l >= 4 * kper) return 0.f;
        const int tap = l / kper, dy = tap / 2, dx = tap % 2;
        const int cin = g * row_elems + (l % kper) * 16 + e;
        const int phase = n / co, o = n % co;
        const int ky = tap_of(phase >> 1, dy), kx = tap_of(phase & 1, dx);
        if (ky < 0 || kx < 0 || cin >= ci) return 0.f;
        return k[((size_t)(ky * 3 + kx) * co + o) * ci + cin];
    }, packed);
    col_bias->resize(4 * co);
    for (int n = 0; n < 4 * co; ++n) (*col_bias)[n] = bias[n % co];
}

// Conv2DTranspose 9x9 stride 1 'same', 16 -> 3 channels (Keras kernel (9,9,3,16)); a GEMM row is 4 adjacent
// pixels (one 128-byte row unit), columns j*3+c for pixel j of the quad; window pixel kx' covers 4r-4+kx'.
static void setup_head(HaloConv* c, const float* k, const float* bias, std::vector<__nv_bfloat16>* packed,
                       std::vector<float>* col_bias) {
    c->launch.N = 16; c->launch.row_bytes = 128; c->launch.epi = EPI_QUAD3;
    c->in_C = 64; c->p.n_groups = 1;
    use_sched(c, SCH_HEAD);
    c->launch.mode = HALO_MODE_F32;
    c->p.out_C = 3; c->p.stats_c = 3;
    pack_b_blocks(c->total_ksteps, 16, [&](int ks, int n, int e) -> float {
        if (n >= 12) return 0.f;
        const int dy = ks / 12, kxp = ks % 12, j = n / 3, o = n % 3;
        const int ky = 8 - dy, kx = j + 8 - kxp;
        if (kx < 0 || kx > 8) return 0.f;
        return k[((size_t)(ky * 9 + kx) * 3 + o) * 16 + e];
    }, packed);
    col_bias->assign(16, 0.f);
    for (int n = 0; n < 12; ++n) (*col_bias)[n] = bias[n % 3];
}

// The same layer with 8 pixels per GEMM row (SCH_HEAD8, needs W % 8 == 0): block-Toeplitz weight units, see halo_gemm.cuh.
static void setup_head8(HaloConv* c, const float* k, const float* bias, std::vector<__nv_bfloat16>* packed,
                        std::vector<float>* col_bias) {
    c->launch.N = 32; c->launch.row_bytes = 128; c->launch.epi = EPI_OCT3;
    c->in_C = 64; c->p.n_groups = 1;
    use_sched(c, SCH_HEAD8);
    c->launch.mode = HALO_MODE_F32;
    c->p.out_C = 3; c->p.stats_c = 3;
    packed->assign((size_t)kHead8Boxes * 256 * 16, __float2bfloat16(0.f));
    for (int dy = 0; dy < 9; ++dy) {
        const int ky = 8 - dy;
        for (int par = 0; par < 2; ++par)
            for (int m = 0; m < kHead8SeqUnits; ++m) {
                const int kx = m + par - 7;                  // copy `par` holds the sequence shifted by one unit
                if (kx < 0 || kx > 8) continue;
                for (int o = 0; o < 3; ++o)
                    for (int e = 0; e < 16; ++e)
                        (*packed)[((((size_t)dy * 2 + par) * kHead8SeqUnits + m) * 4 + o) * 16 + e] =
                            __float2bfloat16(k[((size_t)(ky * 9 + kx) * 3 + o) * 16 + e]);
            }
    }
    col_bias->assign(32, 0.f);
    for (int n = 0; n < 32; ++n) if (n % 4 < 3) (*col_bias)[n] = bias[n % 4];
}

// ------------------------------------------------------------------------------------------------
struct Bf16State {
    int num_sms = 148;
    bool tc_decoder = false;            // expand layers on tensor cores (standard 2-expand geometry)
    bool fuse1 = false;                 // first norm of every residual block applied by the consuming conv's loader warps
    int l2_hints = 0;                   // RST_L2_HINTS (A/B): bits 0-1 trunk conv (loads evict_first / stores evict_last), bit 2 norm pass skip loads
    StemLayout stem_layout{};
    HaloConv stem, contract[4], trunk[10], e0, e1, head;
    __nv_bfloat16 *s_in = nullptr;      // packed stem input
    __nv_bfloat16* enc[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};   // encoder activations (enc[0] = stem output)
    __nv_bfloat16 *bx = nullptr, *by = nullptr, *bz = nullptr;   // bottleneck tensors
    __nv_bfloat16 *ye0 = nullptr, *ye1 = nullptr;                  // decoder tensors (normalised in place)
    float* ylast = nullptr;
    double* stats = nullptr;            // [13][max_batch][F][2]
    size_t stats_bytes = 0, stats_stride = 0;
    ~Bf16State() {
        for (auto q : enc) if (q) cudaFree(q);
        for (void* q : {(void*)s_in, (void*)bx, (void*)by, (void*)bz, (void*)ye0, (void*)ye1, (void*)ylast, (void*)stats})
            if (q) cudaFree(q);
    }
};

int bf16_create(rst_ctx* c) {
    const rst_config& g = c->cfg;
    const int F = g.bottleneck_num_filters;
    if (F != 64 && F != 128)
        return fail(c, RST_ERR_UNSUPPORTED, "bf16 path: bottleneck_num_filters must be 64 or 128 (use the fp32 path otherwise)");
    std::string err;
    if (!umma_init(&err)) return fail(c, RST_ERR_CUDA, err);
    auto st = std::make_shared<Bf16State>();
    if (!stem_layout(g.in_c, &st->stem_layout))
        return fail(c, RST_ERR_UNSUPPORTED, "bf16 path: more than 19 input channels are not supported by the stem packing");
    cudaDeviceProp prop;
    RST_CUDA(c, cudaGetDeviceProperties(&prop, c->device));
    st->num_sms = prop.multiProcessorCount;
    // L2 eviction priorities (B200, profiles/r02_05_l2_hints.md): every convolution's activation loads and the norm pass's skip
    // loads are evict_first -- those tensors are dead (or not needed for two kernels) once read, and the tensor the next kernel
    // reads, this kernel's output, stays in L2 instead.  RST_L2_HINTS=<bits> for A/B: 1 trunk loads, 2 trunk stores evict_last
    // (no gain, off), 4 skip loads, 8 the other convolutions' loads.
    st->l2_hints = ab_env("RST_L2_HINTS") ? atoi(ab_env("RST_L2_HINTS")) : 13;
    const size_t B = g.max_batch;
    const size_t pin = B * g.in_h * g.in_w, pb = B * c->bott_h * c->bott_w;
    RST_CUDA(c, cudaMalloc(&st->s_in, pin * st->stem_layout.row_elems * 2));
    for (size_t i = 0; i < c->contract.size(); ++i) {
        const LayerDesc& L = c->contract[i];
        if ((L.ho % 2 || L.wo % 2) && i + 1 < c->contract.size())
            return fail(c, RST_ERR_UNSUPPORTED, "bf16 path: odd-sized encoder activations");
        RST_CUDA(c, cudaMalloc(&st->enc[i], B * L.ho * L.wo * L.co * 2));
    }
    RST_CUDA(c, cudaMalloc(&st->bx, pb * F * 2));
    RST_CUDA(c, cudaMalloc(&st->by, pb * F * 2));
    RST_CUDA(c, cudaMalloc(&st->bz, pb * F * 2));
    st->tc_decoder = c->n_expand == 2 && F == 128 && g.out_w % 64 == 0;
    if (st->tc_decoder) {
        const size_t p1 = pb * 4, p2 = pb * 16;
        RST_CUDA(c, cudaMalloc(&st->ye0, p1 * 32 * 2));
        RST_CUDA(c, cudaMalloc(&st->ye1, p2 * 16 * 2));
        RST_CUDA(c, cudaMalloc(&st->ylast, p2 * 3 * 4));
    }
    st->stats_stride = B * F * 2;
    st->stats_bytes = 13 * st->stats_stride * sizeof(double);
    RST_CUDA(c, cudaMalloc(&st->stats, st->stats_bytes));
    c->bf16 = st;
    return RST_OK;
}

int bf16_commit(rst_ctx* c) {
    Bf16State* st = c->bf16.get();
    const rst_config& g = c->cfg;
    const int F = g.bottleneck_num_filters, B = g.max_batch;
    std::string err;
    std::vector<__nv_bfloat16> packed;
    std::vector<float> cb, cs, csh;
    // ---- stem ----
    {
        const LayerDesc& L = c->contract[0];
        const Weight* k = c->find_weight(L.name + "/conv/kernel");
        const Weight* b = c->find_weight(L.name + "/conv/bias");
        std::vector<float> scale(L.co), shift(L.co);
        RST_CUDA(c, cudaMemcpy(scale.data(), c->folded[L.name + "/bn/scale"], L.co * 4, cudaMemcpyDeviceToHost));
        RST_CUDA(c, cudaMemcpy(shift.data(), c->folded[L.name + "/bn/shift"], L.co * 4, cudaMemcpyDeviceToHost));
        const char* env2 = ab_env("RST_STEM_PAIRS");
        const bool pairs = (L.ci == 17 || L.ci == 18) && L.co == 32 && L.wi % 2 == 0 && !(env2 && env2[0] == '0');
        if (pairs) {
            if (L.ci == 18) st->stem_layout.row_elems = 32;             // pair rows: 64 elements per two pixels
            (L.ci == 18 ? setup_stem2b : setup_stem2)(&st->stem, k->host.data(), b->host.data(), scale.data(), shift.data(), &packed, &cb, &cs, &csh);
        } else if (!setup_stem(&st->stem, L.ci, L.co, k->host.data(), b->host.data(), scale.data(), shift.data(), &packed, &cb,
                               &cs, &csh)) {
            return fail(c, RST_ERR_UNSUPPORTED, "bf16 path: stem channel layout");
        }
        RST_CUDA(c, st->stem.upload(packed, cb, &cs, &csh));
        st->stem.p.out_H = L.ho; st->stem.p.out_W = pairs ? L.wo / 2 : L.wo;
        if (!st->stem.bind_input(st->s_in, B, L.hi, pairs ? L.wi / 2 : L.wi, &err)) return fail(c, RST_ERR_CUDA, err);
        st->stem.next_is_conv = true;                            // contract_0 follows
        st->stem.l2_hints = (st->l2_hints >> 3) & 1;
    }
    // ---- strided encoder convs ----
    for (size_t i = 1; i < c->contract.size(); ++i) {
        const LayerDesc& L = c->contract[i];
        std::vector<float> scale(L.co), shift(L.co);
        RST_CUDA(c, cudaMemcpy(scale.data(), c->folded[L.name + "/bn/scale"], L.co * 4, cudaMemcpyDeviceToHost));
        RST_CUDA(c, cudaMemcpy(shift.data(), c->folded[L.name + "/bn/shift"], L.co * 4, cudaMemcpyDeviceToHost));
        HaloConv& hc = st->contract[i - 1];
        if (!setup_conv_s2(&hc, L.ci, L.co, c->find_weight(L.name + "/conv/kernel")->host.data(),
                           c->find_weight(L.name + "/conv/bias")->host.data(), scale.data(), shift.data(), &packed, &cb, &cs, &csh))
            return fail(c, RST_ERR_UNSUPPORTED, "bf16 path: strided conv channel counts");
        RST_CUDA(c, hc.upload(packed, cb, &cs, &csh));
        hc.p.out_H = L.ho; hc.p.out_W = L.wo;
        if (!hc.bind_input(st->enc[i - 1], B, L.ho, L.wo, &err)) return fail(c, RST_ERR_CUDA, err);
        hc.next_is_conv = true;                                   // the next contract layer or residual_block_0/conv0 follows
        hc.l2_hints = (st->l2_hints >> 3) & 1;
    }
    // ---- bottleneck ----
    for (int i = 0; i < 10; ++i) {
        const LayerDesc& L = c->residual[i];
        const Weight* k = c->find_weight(L.name + "/kernel");
        const Weight* b = c->find_weight(L.name + "/bias");
        HaloConv& hc = st->trunk[i];
        hc.l2_hints = st->l2_hints & 3;
        setup_conv3x3(&hc, L.ci, F, k->host.data(), b->host.data(), ACT_RELU, &packed, &cb);
        RST_CUDA(c, hc.upload(packed, cb, nullptr, nullptr));
        hc.p.out_H = L.ho; hc.p.out_W = L.wo;
        // Both norm passes run IN PLACE on the tensor the conv just wrote (59 MB at batch 8: still in the 126 MB L2, and it stays
        // there for the next conv's reads).  Block b: conv0 reads the block input X and writes by; norm 1 on by; conv1 reads by
        // and writes Z; norm 2 on Z adds the skip X; Z is the next block's input.  X / Z alternate between bx and bz.
        const int blk = i / 2;
        const void* in = i == 0 ? (const void*)st->enc[c->contract.size() - 1]
                       : (i % 2 == 1 ? (const void*)st->by : (blk % 2 == 1 ? (const void*)st->bz : (const void*)st->bx));
        if (i == 0 && L.ci > 32) return fail(c, RST_ERR_UNSUPPORTED, "bf16 path: bottleneck input wider than 32 channels");
        if (!hc.bind_input(in, B, L.hi, L.wi, &err)) return fail(c, RST_ERR_CUDA, err);
        hc.tmA_fuse = hc.tmA;          // fused-norm mode reads the same tensor (raw), through the loader warps
    }
    {
        // Fused first norm: OPT-IN (RST_FUSE_NORM=1).  Correct (tests/test_gpu_bf16.py) but slower on B200 than conv + in-place
        // pass: 112 us against 48 + 20 us per conv at batch 8 (profiles/r02_03_fused_norm.md) -- the loader warps' instruction
        // stream, not memory, is the limit.  Single style only (the per-pixel blend of two styles needs the weight mip).
        const char* env = ab_env("RST_FUSE_NORM");
        st->fuse1 = g.num_styles == 1 && env && env[0] == '1' && (long long)B * c->bott_h * c->bott_w * F < (1LL << 31);
        for (int i = 1; i < 10; i += 2) st->fuse1 = st->fuse1 && st->trunk[i].two_cta && st->trunk[i].p.n_groups == 2;
    }
    // ---- decoder ----
    if (st->tc_decoder) {
        const LayerDesc& L0 = c->expand[0];
        const LayerDesc& L1 = c->expand[1];
        const LayerDesc& L2 = c->expand[2];
        setup_convt2(&st->e0, L0.ci, L0.co, c->find_weight(L0.name + "/conv/kernel")->host.data(),
                     c->find_weight(L0.name + "/conv/bias")->host.data(), &packed, &cb);
        RST_CUDA(c, st->e0.upload(packed, cb, nullptr, nullptr));
        st->e0.p.out_H = L0.ho; st->e0.p.out_W = L0.wo;
        if (!st->e0.bind_input(st->bz, B, L0.hi, L0.wi, &err)) return fail(c, RST_ERR_CUDA, err);   // block 4 (even) leaves its output in bz
        st->e0.l2_hints = (st->l2_hints >> 3) & 1;
        setup_convt2(&st->e1, L1.ci, L1.co, c->find_weight(L1.name + "/conv/kernel")->host.data(),
                     c->find_weight(L1.name + "/conv/bias")->host.data(), &packed, &cb);
        RST_CUDA(c, st->e1.upload(packed, cb, nullptr, nullptr));
        st->e1.p.out_H = L1.ho; st->e1.p.out_W = L1.wo;
        if (!st->e1.bind_input(st->ye0, B, L1.hi, L1.wi, &err)) return fail(c, RST_ERR_CUDA, err);
        st->e1.l2_hints = (st->l2_hints >> 3) & 1;
        const char* env8 = ab_env("RST_HEAD8");
        const bool head8 = L2.wi % 8 == 0 && !(env8 && env8[0] == '0');
        (head8 ? setup_head8 : setup_head)(&st->head, c->find_weight(L2.name + "/conv/kernel")->host.data(),
                   c->find_weight(L2.name + "/conv/bias")->host.data(), &packed, &cb);
        RST_CUDA(c, st->head.upload(packed, cb, nullptr, nullptr));
        st->head.p.out_H = L2.ho; st->head.p.out_W = L2.wo;
        if (!st->head.bind_input(st->ye1, B, L2.hi, L2.wi / (head8 ? 8 : 4), &err)) return fail(c, RST_ERR_CUDA, err);
        st->head.l2_hints = (st->l2_hints >> 3) & 1;
    }
    return RST_OK;
}

static int norm_pass(rst_ctx* c, const void* x, bool x_f32, void* y, bool y_f32, const __nv_bfloat16* residual,
                     const double* stats, int batch, int P, int C, int width, const float* d_style_params, int param_off,
                     int act, cudaStream_t s, bool y_u8 = false) {
    CinApplyV a;
    a.x = x; a.x_f32 = x_f32; a.y = y; a.y_f32 = y_f32; a.y_u8 = y_u8 ? 1 : 0; a.residual = residual; a.stats = stats;
    a.params = d_style_params;
    a.param_bstride = (long long)c->cfg.num_styles * c->num_style_params;
    a.param_sstride = c->num_style_params;
    a.scale_off = param_off; a.bias_off = param_off + C;
    a.weights = c->cfg.num_styles == 2 ? mip_for_width(c, width) : nullptr;
    a.B = batch; a.P = P; a.C = C; a.num_styles = c->cfg.num_styles; a.act = act;
    a.l2_hints = c->bf16 ? c->bf16->l2_hints >> 2 : 0;
    if (c->cfg.num_styles == 2 && !a.weights) return fail(c, RST_ERR_STATE, "no style-weight mip for this layer width");
    LaunchScope ls(c, s, "cin_apply_bf16");
    RST_CUDA(c, launch_cin_apply_v(a, s));
    return RST_OK;
}

int bf16_transfer_forward(rst_ctx* c, const void* d_content, int content_dtype, const float* d_style_params,
                          const float* d_style_weights, void* d_out, int out_dtype, int batch, cudaStream_t s) {
    Bf16State* st = c->bf16.get();
    const rst_config& g = c->cfg;
    const int F = g.bottleneck_num_filters;
    const int PB = c->bott_h * c->bott_w;
    const long long px = (long long)batch * PB;
    int rc = build_mips(c, d_style_weights, batch, s);
    if (rc) return rc;
    RST_CUDA(c, cudaMemsetAsync(st->stats, 0, st->stats_bytes, s));

    // ---- encoder ----
    {
        LaunchScope ls(c, s, "pack_input");
        RST_CUDA(c, launch_pack_stem_input(d_content, content_dtype == RST_DTYPE_F16 ? 1 : 0, st->s_in, batch, g.in_h, g.in_w, g.in_c, st->stem_layout.n_real,
                                           st->stem_layout.row_elems, st->stem.launch.sched == SCH_STEM2 || st->stem.launch.sched == SCH_STEM2B ? 1 : 0, s));
    }
    {
        LaunchScope ls(c, s, "stem_umma");
        RST_CUDA(c, st->stem.run(st->enc[0], false, nullptr, batch, st->num_sms, s));
    }
    record_tap(c, c->contract[0].name, st->enc[0], (int64_t)batch * g.in_h * g.in_w * c->contract[0].co, true, s);
    for (size_t i = 1; i < c->contract.size(); ++i) {           // 3x3 stride-2 convs over the space-to-depth view
        const LayerDesc& L = c->contract[i];
        { LaunchScope ls(c, s, "conv_s2_umma"); RST_CUDA(c, st->contract[i - 1].run(st->enc[i], false, nullptr, batch, st->num_sms, s)); }
        record_tap(c, L.name, st->enc[i], (int64_t)batch * L.ho * L.wo * L.co, true, s);
    }

    // ---- residual bottleneck (styleTransfer.py:144-185) ----
    int cursor = 0;
    for (int b = 0; b < 5; ++b) {
        const std::string name = "residual_block_" + std::to_string(b);
        double* st0 = st->stats + (size_t)(2 * b) * st->stats_stride;
        double* st1 = st->stats + (size_t)(2 * b + 1) * st->stats_stride;
        { LaunchScope ls(c, s, "conv3x3_umma"); RST_CUDA(c, st->trunk[2 * b].run(st->by, false, st0, batch, st->num_sms, s)); }
        record_tap(c, name + "/conv0/relu", st->by, px * F, true, s);
        __nv_bfloat16* z = b % 2 == 0 ? st->bz : st->bx;            // conv1 output, then (in place) the block output
        const __nv_bfloat16* skip = b % 2 == 0 ? st->bx : st->bz;   // the block input
        if (st->fuse1 && !c->keep_taps) {
            // conv1 reads the RAW conv0 output; relu(cin(.)) happens in its loader warps (no pass, no normalised tensor)
            HaloGemmParams fin;
            fin.fuse = 1; fin.fin_x = st->by; fin.fin_stats = st0; fin.fin_params = d_style_params;
            fin.fin_param_bstride = (long long)g.num_styles * c->num_style_params;
            fin.fin_scale_off = cursor; fin.fin_bias_off = cursor + F;
            { LaunchScope ls(c, s, "conv3x3_umma"); RST_CUDA(c, st->trunk[2 * b + 1].run(z, false, st1, batch, st->num_sms, s, &fin)); }
        } else {
            rc = norm_pass(c, st->by, false, st->by, false, nullptr, st0, batch, PB, F, c->bott_w, d_style_params, cursor, ACT_RELU, s);
            if (rc) return rc;
            record_tap(c, name + "/conv0/cin", st->by, px * F, true, s);
            { LaunchScope ls(c, s, "conv3x3_umma"); RST_CUDA(c, st->trunk[2 * b + 1].run(z, false, st1, batch, st->num_sms, s)); }
        }
        record_tap(c, name + "/conv1/relu", z, px * F, true, s);
        rc = norm_pass(c, z, false, z, false, b == 0 ? nullptr : skip, st1, batch, PB, F, c->bott_w, d_style_params,
                       cursor + 2 * F, ACT_NONE, s);
        if (rc) return rc;
        record_tap(c, name, z, px * F, true, s);
        cursor += 4 * F;
    }

    // ---- decoder (styleTransfer.py:95-141, :260-276) ----
    if (!st->tc_decoder) {
        float* x = c->act[0];
        float* t1 = c->act[1];
        {
            LaunchScope ls(c, s, "convert");
            RST_CUDA(c, launch_bf16_to_f32_slice(st->bz, x, px, F, F, s));
        }
        if (out_dtype != RST_DTYPE_F32)
            return fail(c, RST_ERR_UNSUPPORTED, "uint8 output needs the tensor-core decoder (2 expand blocks, 128 filters, width % 64 == 0)");
        return fp32_expand_stage(c, x, t1, d_style_params, cursor, (float*)d_out, batch, s);
    }
    const LayerDesc& L0 = c->expand[0];
    const LayerDesc& L1 = c->expand[1];
    const LayerDesc& L2 = c->expand[2];
    double* se0 = st->stats + (size_t)10 * st->stats_stride;
    double* se1 = st->stats + (size_t)11 * st->stats_stride;
    double* se2 = st->stats + (size_t)12 * st->stats_stride;
    { LaunchScope ls(c, s, "convt_umma"); RST_CUDA(c, st->e0.run(st->ye0, false, se0, batch, st->num_sms, s)); }
    record_tap(c, L0.name + "/conv", st->ye0, (int64_t)batch * L0.ho * L0.wo * L0.co, true, s);
    rc = norm_pass(c, st->ye0, false, st->ye0, false, nullptr, se0, batch, L0.ho * L0.wo, L0.co, L0.wo, d_style_params, cursor,
                   ACT_RELU, s);
    if (rc) return rc;
    record_tap(c, L0.name, st->ye0, (int64_t)batch * L0.ho * L0.wo * L0.co, true, s);
    cursor += 2 * L0.co;
    { LaunchScope ls(c, s, "convt_umma"); RST_CUDA(c, st->e1.run(st->ye1, false, se1, batch, st->num_sms, s)); }
    record_tap(c, L1.name + "/conv", st->ye1, (int64_t)batch * L1.ho * L1.wo * L1.co, true, s);
    rc = norm_pass(c, st->ye1, false, st->ye1, false, nullptr, se1, batch, L1.ho * L1.wo, L1.co, L1.wo, d_style_params, cursor,
                   ACT_RELU, s);
    if (rc) return rc;
    record_tap(c, L1.name, st->ye1, (int64_t)batch * L1.ho * L1.wo * L1.co, true, s);
    cursor += 2 * L1.co;
    { LaunchScope ls(c, s, "head_umma"); RST_CUDA(c, st->head.run(st->ylast, true, se2, batch, st->num_sms, s)); }
    record_tap(c, L2.name + "/conv", st->ylast, (int64_t)batch * L2.ho * L2.wo * 3, false, s);
    rc = norm_pass(c, st->ylast, true, d_out, true, nullptr, se2, batch, L2.ho * L2.wo, 3, L2.wo, d_style_params, cursor,
                   ACT_SIGMOID, s, out_dtype == RST_DTYPE_U8);
    if (rc) return rc;
    if (out_dtype == RST_DTYPE_F32) record_tap(c, L2.name, d_out, (int64_t)batch * L2.ho * L2.wo * 3, false, s);
    return RST_OK;
}

// ------------------------------------------------------------------------------------------------
// stand-alone operator (tests): one convolution through the tensor-core kernel with fp32 tensors at the boundary.
// Not a hot path: allocates and frees its scratch.  Supported: 3x3 s1 conv (co 64/128), 9x9 s1 conv (co 32,
// ci <= 19), Conv2DTranspose 3x3 s2 (co 16/32) and Conv2DTranspose 9x9 s1 16 -> 3.
// ------------------------------------------------------------------------------------------------
int op_conv2d_bf16(const float* d_x, const float* d_kernel, const float* d_bias, float* d_y, int batch, int h, int w, int ci,
                   int co, int kh, int kw, int stride, int transposed, int act, cudaStream_t s, std::string* err) {
    enum { K3, STEM, CONVT2, HEAD, S2, NONE } kind = NONE;
    if (!transposed && kh == 3 && kw == 3 && stride == 1 && (co == 64 || co == 128)) kind = K3;
    else if (!transposed && kh == 9 && kw == 9 && stride == 1 && co == 32 && ci <= 19) kind = STEM;
    else if (!transposed && kh == 3 && kw == 3 && stride == 2 && (ci == 16 || ci == 32) && (co == 16 || co == 32) && h % 2 == 0 &&
             w % 2 == 0) kind = S2;
    else if (transposed && kh == 3 && kw == 3 && stride == 2 && (co == 16 || co == 32) && (ci <= 32 || ci % 64 == 0)) kind = CONVT2;
    else if (transposed && kh == 9 && kw == 9 && stride == 1 && co == 3 && ci == 16 && w % 4 == 0) kind = HEAD;
    const bool conv_like = kind == K3 || kind == STEM || kind == S2;
    if (kind == NONE || act != (conv_like ? ACT_RELU : ACT_NONE)) {
        *err = "rst_op_conv2d(bf16): shape/activation not mapped onto the tensor-core kernel (convs: ReLU, transposed convs: none)";
        return RST_ERR_UNSUPPORTED;
    }
    if (!umma_init(err)) return RST_ERR_CUDA;
    const size_t kelems = (size_t)kh * kw * ci * co;
    std::vector<float> hk(kelems), hb(co, 0.f);
    cudaError_t e = cudaMemcpy(hk.data(), d_kernel, kelems * 4, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && d_bias) e = cudaMemcpy(hb.data(), d_bias, co * 4, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { *err = cudaGetErrorString(e); return RST_ERR_CUDA; }
    HaloConv hc;
    std::vector<__nv_bfloat16> packed;
    std::vector<float> cb, cs, csh;
    StemLayout SL{};
    int oh = h, ow = w, wru = w;
    bool y_f32 = false, stem_pairs = false;
    if (kind == K3) {
        setup_conv3x3(&hc, ci, co, hk.data(), hb.data(), act, &packed, &cb);
    } else if (kind == STEM) {
        std::vector<float> one(co, 1.f), zero(co, 0.f);
        stem_layout(ci, &SL);
        if ((ci == 17 || ci == 18) && co == 32 && w % 2 == 0 && !ab_env("RST_STEM_PAIRS_OFF")) {
            if (ci == 18) SL.row_elems = 32;
            (ci == 18 ? setup_stem2b : setup_stem2)(&hc, hk.data(), hb.data(), one.data(), zero.data(), &packed, &cb, &cs, &csh);
            wru = w / 2; stem_pairs = true;
        } else {
            setup_stem(&hc, ci, co, hk.data(), hb.data(), one.data(), zero.data(), &packed, &cb, &cs, &csh);
        }
    } else if (kind == S2) {
        std::vector<float> one(co, 1.f), zero(co, 0.f);
        setup_conv_s2(&hc, ci, co, hk.data(), hb.data(), one.data(), zero.data(), &packed, &cb, &cs, &csh);
        oh = h / 2; ow = w / 2; wru = w / 2;
    } else if (kind == CONVT2) {
        setup_convt2(&hc, ci, co, hk.data(), hb.data(), &packed, &cb);
        oh = 2 * h; ow = 2 * w;
    } else {
        const char* env8 = ab_env("RST_HEAD8");
        const bool head8 = w % 8 == 0 && !(env8 && env8[0] == '0');
        (head8 ? setup_head8 : setup_head)(&hc, hk.data(), hb.data(), &packed, &cb);
        wru = w / (head8 ? 8 : 4); y_f32 = true;
    }
    e = (kind == STEM || kind == S2) ? hc.upload(packed, cb, &cs, &csh) : hc.upload(packed, cb, nullptr, nullptr);
    const long long pin = (long long)batch * h * w, pout = (long long)batch * oh * ow;
    __nv_bfloat16 *xb = nullptr, *yb = nullptr;
    const int in_c_dev = kind == HEAD ? 16 : kind == S2 ? ci : hc.in_C;
    if (e == cudaSuccess) e = cudaMalloc(&xb, pin * in_c_dev * 2);
    if (e == cudaSuccess && !y_f32) e = cudaMalloc(&yb, pout * co * 2);
    int rc = RST_OK;
    if (e == cudaSuccess) {
        if (kind == STEM) e = launch_pack_stem_input(d_x, 0, xb, batch, h, w, ci, SL.n_real, SL.row_elems, stem_pairs ? 1 : 0, s);
        else e = launch_f32_to_bf16_pad(d_x, xb, pin, ci, in_c_dev, s);
    }
    if (e == cudaSuccess) {
        hc.p.out_H = oh; hc.p.out_W = stem_pairs ? ow / 2 : ow;
        if (!hc.bind_input(xb, batch, kind == S2 ? oh : h, wru, err)) rc = RST_ERR_CUDA;
    }
    if (e == cudaSuccess && rc == RST_OK) {
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        e = hc.run(y_f32 ? (void*)d_y : (void*)yb, y_f32, nullptr, batch, sms, s);
        if (e == cudaSuccess && !y_f32) e = launch_bf16_to_f32_slice(yb, d_y, pout, co, co, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    }
    for (void* q : {(void*)xb, (void*)yb}) if (q) cudaFree(q);
    if (e != cudaSuccess) { *err = std::string("rst_op_conv2d(bf16): ") + cudaGetErrorString(e); return RST_ERR_CUDA; }
    return rc;
}

}  // namespace rst
