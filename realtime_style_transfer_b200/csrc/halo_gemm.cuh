// tcgen05 "halo GEMM": the one tensor-core kernel behind every convolution of the transfer network.
//
// A CTA tile is 128 GEMM rows = an 8-row x 16-"row unit" patch of the output grid (row index r = wru*8 + h).
// A row unit (RU) is whatever one shared-memory row of the A operand holds: one pixel with 64 channels
// (128 B, SWIZZLE_128B), one pixel with 32 channels (64 B, SWIZZLE_64B), or four adjacent pixels with 16
// channels (128 B) for the 3-channel output head.
//
// The input patch including its halo is TMA-loaded ONCE per (tile, channel group) with out-of-bounds rows
// zero-filled (= 'same' padding).  Every K-step (one tcgen05.mma, K = 16) reads its A operand through a
// shared-memory descriptor whose start address is the halo base plus a per-K-step byte offset taken from a
// small table: filter taps are shifted views of the same bytes, never copies.  (Verified on B200 with
// tests/cuda/umma_probe.cu: the swizzle XOR is a function of the absolute shared-memory address, so starts
// that are not 1024-byte aligned are legal with base_offset = 0.)
//
// B (weights) is pre-packed on the host as blocks of 4 K-steps: [block][N rows][4 x 16 bf16] (128 B rows,
// SWIZZLE_128B, K-major).  Small layers keep all blocks resident in shared memory for the CTA's lifetime;
// the 128->128 bottleneck convs stream them through a ring.
//
// Warp roles (352 threads): 0 = A TMA producer, 1 = B TMA producer, 2 = MMA issuer (one elected lane),
// 3..10 = epilogue (TMEM -> registers -> global; two warps per TMEM lane quadrant split the column chunks when the
// tile has more than one 32-column chunk -- the epilogue instruction stream was the bottleneck of the 128-column
// kernels), two TMEM accumulator stages.
#pragma once

#include <functional>

#include "rst_internal.cuh"
#include "umma.cuh"

namespace rst {

enum HaloEpi : int {
    EPI_NHWC = 0,     // out[n, gh, gw, col]                       (conv)
    EPI_CONVT2 = 1,   // col = phase*Cq + c -> out[n, 2gh+a, 2gw+b, c] (stride-2 transposed conv as 4 phases)
    EPI_QUAD3 = 2,    // col = j*3 + c (j<4) -> out[n, gh, 4gw+j, c]   (3-channel head, 4 pixels per row unit)
    EPI_OCT3 = 3,     // col = j*4 + c (j<8, c<3) -> out[n, gh, 8gw+j, c]  (3-channel head, 8 pixels per GEMM row)
};

// ---- K-step schedules: compile-time A-operand offsets so that the MMA issue loop is fully unrolled ----------------
// SCH_C3: 3x3 taps x (row_bytes/32) channel slices, halo 10x18.     SCH_T2: 2x2 taps (stride-2 transposed conv), halo 9x17.
// SCH_HEAD: 9 rows x 12 window pixels over 4-pixel row units, halo 16x18.
// SCH_STEM2B: the 18-channel variant of SCH_STEM2.  A pair row is [pixel 0: 16 real | pixel 1: 16 real | window of channel 16 |
//          window of channel 17] (4 x 16 slots = 128 B, both windows x0-4 .. x0+5 shared by the pair): 90 + 9 + 9 K-steps.  Unit
//          sequences of consecutive row taps share their zero unit ([0, W8..W0] x 9, then one 0) so that everything stays resident.
// SCH_HEAD8: the same 9x9 transposed conv with EIGHT pixels per GEMM row = two adjacent 4-pixel row units (the A descriptor's
//          8-row-group stride is doubled): 9 rows x 16 window pixels, N = 8 x (3 + 1 pad) columns, halo 16x35.  A third fewer
//          K-steps per pixel at the same ~41 cycles per MMA.  B is a block-Toeplitz matrix over 4-row weight units (one per
//          column tap); K-step s reads 8 consecutive units starting at 15 - s from the sequence [7 zeros, U0..U8, 7 zeros].
//          A SWIZZLE_32B atom is 8 rows = TWO units, so the sequence is stored twice per row tap, the second copy shifted
//          by one unit, and odd starts read the shifted copy: every descriptor start stays atom-aligned.
// SCH_STEM + 4*REAL + NV: 9x9 taps over the 16 real channels (REAL) plus 9 row taps per windowed channel group (NV).
// SCH_S2D: 3x3 stride-2 conv over the space-to-depth view (h2, row parity, w2, [col parity x channels]) of the input:
//          row taps ky -> (h2 + ky/2, parity ky%2); column taps kx -> (w2 + kx/2, parity kx%2) = channel slice of the row.
// SCH_STEM2: the 17-channel stem with TWO pixels per GEMM row (row = pixel pair = 2 x [16 real | 16 windowed] = 128 B):
//          N = 2 x 32 output columns, 10 window pixels per row tap.  The block-Toeplitz weight matrix is never
//          materialised: B is stored as 32-row "units" [0, W(ky,8), ..., W(ky,0), 0] per row tap and every K-step reads
//          TWO ADJACENT units through a shifted SWIZZLE_32B descriptor (the halo trick applied to the weights).
//          The 17th channel is stored ONCE per pair, as the 10-wide window x0-4 .. x0+5 in the even pixel's spare slots:
//          one K-step per row tap serves both pixels (units [Wa, Wb] = the 9 taps at slot offsets 0 and 1): 90 + 9 K-steps.
enum HaloSched : int { SCH_C3 = 0, SCH_T2 = 1, SCH_HEAD = 2, SCH_S2D = 3, SCH_STEM2 = 4, SCH_HEAD8 = 5, SCH_STEM2B = 6, SCH_STEM = 10 };
__host__ __device__ constexpr bool sched_b_units(int sch) { return sch == SCH_STEM2 || sch == SCH_HEAD8 || sch == SCH_STEM2B; }
constexpr int kStem2Units = 9 * 11 + 9 * 2 + 2;      // 99 real + 18 window units + 2 zero units (padding K-step)
constexpr int kStem2Boxes = (kStem2Units * 32 + 255) / 256;   // TMA boxes of 256 rows (8 KB)
// unit index (1 KB each) where K-step ks of SCH_STEM2 starts reading its 64 B rows
__host__ __device__ constexpr int sched_b_unit(int sch, int ks) {
    return sch != SCH_STEM2 ? 0 : ks < 90 ? (ks / 10) * 11 + (9 - ks % 10) : 99 + (ks - 90) * 2;
}
constexpr int kStem2BUnits = 9 * 10 + 1 + 2 * 9 * 2;              // 91 real + 36 window units (no padding K-step: 108 % 4 == 0)
constexpr int kStem2BBoxes = (kStem2BUnits * 32 + 255) / 256;     // 16
// SCH_HEAD8: 9 row taps x 2 copies x 24 units of 4 rows x 32 B (128 B)
constexpr int kHead8SeqUnits = 24;
constexpr int kHead8Rows = 9 * 2 * kHead8SeqUnits * 4;                 // 1728 rows of 32 B
constexpr int kHead8Boxes = (kHead8Rows + 255) / 256;                  // 7 TMA boxes of 256 rows
// B start of K-step ks in 16-byte units relative to the start of the resident weights
__host__ __device__ constexpr int sched_b_off16(int sch, int ks) {
    if (sch == SCH_STEM2) return sched_b_unit(sch, ks) * 64;
    if (sch == SCH_STEM2B) return (ks < 90 ? (ks / 10) * 10 + (9 - ks % 10) : 91 + (ks - 90) * 2) * 64;
    if (sch == SCH_HEAD8) { const int dy = ks / 16, m0 = 15 - ks % 16, par = m0 & 1; return ((dy * 2 + par) * kHead8SeqUnits + m0 - par) * 8; }
    return 0;
}
__host__ __device__ constexpr int sched_b_boxes(int sch) {
    return sch == SCH_STEM2 ? kStem2Boxes : sch == SCH_STEM2B ? kStem2BBoxes : sch == SCH_HEAD8 ? kHead8Boxes : 0;
}
// GEMM rows advance by this many row units along W (SCH_HEAD8: an 8-pixel row = two 4-pixel units)
__host__ __device__ constexpr int sched_a_unit_stride(int sch) { return sch == SCH_HEAD8 ? 2 : 1; }
__host__ __device__ constexpr int sched_real_ksteps(int sch, int rowb) {
    return sch == SCH_C3 ? 9 * (rowb / 32) : sch == SCH_T2 ? 4 * (rowb / 32) : sch == SCH_HEAD ? 108
           : sch == SCH_S2D ? 3 * (rowb / 32 + rowb / 64) : sch == SCH_STEM2 ? 99 : sch == SCH_STEM2B ? 108 : sch == SCH_HEAD8 ? 144
           : 81 * ((sch - SCH_STEM) / 4) + 9 * ((sch - SCH_STEM) % 4);
}
__host__ __device__ constexpr int sched_ksteps(int sch, int rowb) { return (sched_real_ksteps(sch, rowb) + 3) / 4 * 4; }
__host__ __device__ constexpr int sched_halo_h(int sch) { return sch == SCH_C3 ? 10 : sch == SCH_T2 ? 9 : sch == SCH_S2D ? 18 : 16; }
__host__ __device__ constexpr int sched_halo_w(int sch) { return sch == SCH_C3 ? 18 : sch == SCH_T2 || sch == SCH_S2D ? 17 : sch == SCH_HEAD ? 18 : sch == SCH_STEM2 || sch == SCH_STEM2B ? 20 : sch == SCH_HEAD8 ? 35 : 24; }
__host__ __device__ constexpr int sched_oy(int sch) { return sch == SCH_S2D ? 0 : sch == SCH_C3 || sch == SCH_T2 ? -1 : -4; }
__host__ __device__ constexpr int sched_ox(int sch) { return sch == SCH_S2D ? 0 : sch == SCH_C3 || sch == SCH_T2 ? -1 : sch == SCH_HEAD || sch == SCH_HEAD8 ? -1 : sch == SCH_STEM2 || sch == SCH_STEM2B ? -2 : -4; }
// byte offset of K-step ks into the halo patch
__host__ __device__ constexpr int sched_off(int sch, int rowb, int ks) {
    if (ks >= sched_real_ksteps(sch, rowb)) return 0;
    if (sch == SCH_C3) { const int kper = rowb / 32, tap = ks / kper; return ((tap % 3) * 10 + tap / 3) * rowb + (ks % kper) * 32; }
    if (sch == SCH_T2) { const int kper = rowb / 32, tap = ks / kper; return ((tap % 2) * 9 + tap / 2) * rowb + (ks % kper) * 32; }
    if (sch == SCH_HEAD) { const int dy = ks / 12, kx = ks % 12; return ((kx / 4) * 16 + dy) * 128 + (kx % 4) * 32; }
    if (sch == SCH_HEAD8) { const int dy = ks / 16, kx = ks % 16; return ((kx / 4) * 16 + dy) * 128 + (kx % 4) * 32; }
    if (sch == SCH_STEM2B) {
        if (ks < 90) { const int ky = ks / 10, kxp = ks % 10; return ((kxp / 2) * 16 + ky) * 128 + (kxp % 2) * 32; }
        const int g = (ks - 90) / 9, ky = (ks - 90) % 9;
        return (2 * 16 + ky) * 128 + 64 + g * 32;        // the pair's shared window of channel 16 + g
    }
    if (sch == SCH_STEM2) {
        if (ks < 90) { const int ky = ks / 10, kxp = ks % 10; return ((kxp / 2) * 16 + ky) * 128 + (kxp % 2) * 64; }
        const int ky = ks - 90;
        return (2 * 16 + ky) * 128 + 32;                 // window slots of the pair's even pixel
    }
    if (sch == SCH_S2D) {
        // per row tap: kper slices at w2+0 (column taps 0,1) then kper/2 slices at w2+1 (column tap 2)
        const int kper = rowb / 32, per_ky = kper + kper / 2, ky = ks / per_ky, l = ks % per_ky;
        const int w2off = l >= kper ? 1 : 0, k = l >= kper ? l - kper : l;
        return ((w2off * 2 + ky % 2) * 9 + ky / 2) * rowb + k * 32;
    }
    const int real = (sch - SCH_STEM) / 4;
    if (ks < 81 * real) return ((ks % 9) * 16 + ks / 9) * rowb;
    const int v = (ks - 81 * real) / 9, ky = (ks - 81 * real) % 9;
    return (4 * 16 + ky) * rowb + 32 * real + v * 32;
}

struct HaloGemmParams {
    // output grid in rows / row units, tiling
    int B = 0, H = 0, WRU = 0, tiles_h = 0, tiles_w = 0;
    // A halo
    int oy = 0, ox = 0;              // halo origin relative to the tile origin (rows, RUs)
    int a_w_mul = 1;                 // TMA W coordinate = tile column * a_w_mul + ox (SCH_HEAD8: GEMM rows span two row units)
    int halo_h = 10, halo_w = 18;    // box extent (rows, RUs)
    int n_groups = 1;                // halo loads per tile; group g loads channel coordinate g * row_elems
    int ksteps = 0;                  // K-steps per group, multiple of 4
    int b_resident = 0;
    int n_astages = 3, n_bstages = 6;
    // epilogue
    void* y = nullptr; int y_f32 = 0;
    int out_H = 0, out_W = 0, out_C = 0;
    const float* bias = nullptr;         // [N] (already expanded to GEMM columns)
    const float* post_scale = nullptr;   // [N] optional per-column affine after act1 (inference BatchNorm)
    const float* post_shift = nullptr;
    double* stats = nullptr;             // (B, stats_c, 2) [sum, sumsq] or null
    int stats_c = 0;
    int a_wrap = 0;                      // split tf32: the tensor holds a_wrap channel groups [x_hi | x_lo]; groups >= a_wrap re-read
                                         // group g - a_wrap (the conv runs over [x_hi | x_lo | x_hi] without storing x_hi twice)
    int store_exact = 0;                 // MODE_TF32: store the fp32 accumulator as is (split-tf32 convs) instead of rounding to tf32
    // Fused input transform (halo_gemm2.cu, fuse != 0): the A operand is not the bound tensor as stored but the conditional
    // instance norm of the RAW output of the previous convolution, computed by loader warps on the way into shared memory:
    //   fuse 1:  A = relu(a[n,c] * x + b[n,c])                  (first norm of a residual block, styleTransfer.py:173-182)
    //   fuse 2:  A = skip + a[n,c] * x + b[n,c]  (or without skip), also written to fin_out: the block output (:184)
    // with a = rsqrt(var + eps) * scale, b = bias - mean * a from the statistics the previous conv's epilogue accumulated.
    int fuse = 0;
    const __nv_bfloat16* fin_x = nullptr;      // raw previous conv output, same geometry as the bound tensor
    const __nv_bfloat16* fin_skip = nullptr;
    __nv_bfloat16* fin_out = nullptr;
    const double* fin_stats = nullptr;         // (B, C, 2) [sum, sumsq] over the H*W pixels of fin_x
    const float* fin_params = nullptr;         // style parameters, element (b, j) at fin_params[b * fin_param_bstride + j]
    long long fin_param_bstride = 0;
    int fin_scale_off = 0, fin_bias_off = 0;
    float fin_eps = 1e-5f;
    int l2_hints = 0;                          // trunk kernel: bit 0 = activation loads evict_first, bit 1 = output stores evict_last
    int pdl_trigger = 0;                       // pdl only: let the successor be staged early (set when the successor is another
                                               // convolution: its CTAs cannot park beside this kernel's, they need the shared memory)
    int pdl = 0;                               // launch with the programmatic-dependent-launch attribute (rst_internal.cuh): the weights,
                                               // bias and folded BN constants must not be written by the kernel launched just before
    int fuse_dbg = 0;                          // RST_EXPERIMENTS builds only: 1 = skip the global loads, 2 = skip the smem stores (timing)
};

constexpr int HALO_MODE_RELU = 1, HALO_MODE_POST = 2, HALO_MODE_F32 = 4, HALO_MODE_TF32 = 8;

struct HaloGemmLaunch {
    int N = 128;            // GEMM N (output columns): 16, 32, 64 or 128
    int row_bytes = 128;    // 128 (SWIZZLE_128B) or 64 (SWIZZLE_64B)
    int epi = EPI_NHWC;
    int mode = 0;           // HALO_MODE_* bits (compile-time epilogue variant)
    int sched = SCH_C3;     // HaloSched (compile-time K-step schedule)
    size_t smem_bytes = 0;  // filled by halo_gemm_plan
};

// Chooses stage counts for the shared-memory budget and fills p.n_astages / p.n_bstages / l.smem_bytes.
bool halo_gemm_plan(HaloGemmLaunch* l, HaloGemmParams* p, std::string* err);
cudaError_t launch_halo_gemm(const HaloGemmLaunch& l, const CUtensorMap& tmA, const CUtensorMap& tmB,
                             const HaloGemmParams& p, int num_sms, cudaStream_t s);

// ---- host helpers -------------------------------------------------------------------------------
// Activation tensor (B, H, WRU, row_elems*n_groups) bf16 viewed for the halo box (row_elems, halo_h, halo_w, 1).
bool encode_halo_map(CUtensorMap* out, const void* base, int B, int H, int WRU, int C, int row_elems, int halo_h,
                     int halo_w, std::string* err, bool atom32 = false);   // atom32: 128-byte swizzle in 32-byte chunks
// Packed weights: nblocks*N rows of 64 bf16; box = (64, N).
// Space-to-depth view of an NHWC tensor (B, H, W, C), H and W even: dims (2C, H/2, 2, W/2, B), box (2C, 9, 2, 17, 1).
bool encode_s2d_map(CUtensorMap* out, const void* base, int B, int H, int W, int C, std::string* err);
bool encode_weight_map(CUtensorMap* out, const void* base, int nblocks, int N, std::string* err, int box_rows = 0);
// Weight units (SCH_STEM2): rows of 16 bf16 (32 B), box = (16, 256), SWIZZLE_32B.
bool encode_weight_unit_map(CUtensorMap* out, const void* base, int rows, std::string* err);

// 2-CTA (cta_group::2) weight-resident kernel for the 128-filter 3x3 bottleneck convs (halo_gemm2.cu); tmB_half has a 64-row box.
size_t halo_gemm2_smem_bytes(int n_groups);
cudaError_t launch_halo_gemm2(const CUtensorMap& tmA, const CUtensorMap& tmB_half, const HaloGemmParams& p, int num_sms,
                              cudaStream_t s);
// 2-CTA variant of the pair-pixel stems (halo_stem2cta.cu; sch = SCH_STEM2 or SCH_STEM2B): tmB_units must cover
// sched_b_boxes(sch) * 256 + 32 rows (the peer CTA's copy of the unit array starts one unit later)
size_t halo_stem2cta_smem_bytes(int sch);
cudaError_t launch_halo_stem2cta(int sch, const CUtensorMap& tmA, const CUtensorMap& tmB_units, const HaloGemmParams& p,
                                 int num_sms, cudaStream_t s);
// Packs B: f(kstep, n, e) -> weight of K-step `kstep`, output column n, K element e (0..15).
void pack_b_blocks(int total_ksteps, int N, const std::function<float(int, int, int)>& f, std::vector<__nv_bfloat16>* out);

// ---- tf32 3x3 convolution on fp32 NHWC tensors (conv_tf32.cu): the VGG16 loss model's layers and their input gradients ----
// 3x3, stride 1, 'same'; Cin % 32 == 0, Cout % 64 == 0.  Operands are fp32 bit patterns consumed as tf32 by tcgen05
// (kind::tf32, fp32 accumulation); weights are rounded to tf32 (nearest) once, outputs are stored rounded to tf32.
// x_lo = rna_tf32(x - trunc_tf32(x)), n % 4 == 0 (conv_tf32.cu)
cudaError_t launch_tf32_lo(const float* in, float* out, long long n, cudaStream_t s);
struct Tf32Conv3x3 {
    int ci = 0, co = 0, nb = 128, nblk = 0;
    bool relu = false;
    float* w_packed = nullptr;          // [nblk][n_groups * 9 blocks][nb rows][32 floats]
    float* bias = nullptr;              // [co] or null (owned copy made by setup)
    const float* ext_bias = nullptr;    // or a caller-owned device vector (set by repack)
    int ci_layer = 0, co_layer = 0;
    bool input_gradient = false;
    HaloGemmLaunch launch;
    HaloGemmParams p;
    std::vector<CUtensorMap> tmB;       // one weight map per block of nb output channels
    struct BoundInput { const void* x; int B, H, W; CUtensorMap tm; };
    std::vector<BoundInput> inputs;     // tensor maps of the input tensors seen so far
    Tf32Conv3x3() = default;
    Tf32Conv3x3(const Tf32Conv3x3&) = delete;
    Tf32Conv3x3& operator=(const Tf32Conv3x3&) = delete;
    ~Tf32Conv3x3();
    // k: Keras kernel (3,3,ci_layer,co_layer).  input_gradient = false: y = [relu](conv(x, k) + bias), ci = ci_layer.
    // input_gradient = true: y = d conv / d input applied to x = gradient w.r.t. the layer output (ci = co_layer, co = ci_layer).
    bool setup(int ci_layer, int co_layer, const float* k, const float* bias_host, bool relu, bool input_gradient, std::string* err,
               bool split = false);
    // Same without weights: allocate and plan only; repack() then (re)builds the packed tf32 weights from a DEVICE kernel
    // tensor (3,3,ci_layer,co_layer), e.g. once per training step after the optimizer moved the variables.
    // split = true: error-compensated "3 x tf32" arithmetic with fp32-level accuracy: x = x_hi + x_lo, w = w_hi + w_lo (each part
    // on the tf32 grid), y = x_hi*w_hi + x_lo*w_hi + x_hi*w_lo accumulated in fp32 -- realised as ONE conv over 3*ci channels
    // [x_hi | x_lo | x_hi] with weights [w_hi | w_hi | w_lo]; run() then needs a scratch tensor of 3x the input size.
    bool setup_shape(int ci_layer, int co_layer, bool relu, bool input_gradient, std::string* err, bool split = false);
    bool split = false;
    size_t scratch_floats(int B, int H, int W) const { return split ? (size_t)B * H * W * (ci / 3) * 2 : 0; }   // [x_hi | x_lo]
    // relu_mask (split convs only): the input is masked by (relu_mask > 0) while it is expanded (ReLU backward fused into the pass)
    cudaError_t run_split(const float* x, float* scratch, float* y, int B, int H, int W, int num_sms, cudaStream_t s, std::string* err,
                          const float* relu_mask = nullptr);
    cudaError_t repack(const float* d_kernel, const float* d_bias, cudaStream_t s);
    cudaError_t run(const float* x, float* y, int B, int H, int W, int num_sms, cudaStream_t s, std::string* err);
};

// ---- bf16 elementwise companions ------------------------------------------------------------------
cudaError_t launch_f32_to_bf16_pad(const float* x, __nv_bfloat16* y, long long pixels, int c_in, int c_out, cudaStream_t s);
cudaError_t launch_bf16_to_f32_slice(const __nv_bfloat16* x, float* y, long long pixels, int c_in, int c_out, cudaStream_t s);

// Stem input packing: fp32 NHWC (C channels) -> bf16 rows of `row_elems`: [n_real real channels, zero padded to
// 16 when n_real > 0][one 16-slot group per remaining channel c: x[y, x-4 .. x+4, c] then 7 zeros].
// pair_window (SCH_STEM2, one windowed channel, even W): the group of an EVEN pixel holds x[y, x-4 .. x+5, c] (10 values, shared
// by the pixel pair), the group of an odd pixel is zero.
// pair_window with TWO windowed channels (SCH_STEM2B, C = 18, even W, row_elems = 32 per pixel): a pixel PAIR is stored as
// [even pixel: 16 real | odd pixel: 16 real | window of channel 16 | window of channel 17], windows as above.
// x: fp32, or fp16 when x_f16 != 0.
cudaError_t launch_pack_stem_input(const void* x, int x_f16, __nv_bfloat16* y, int B, int H, int W, int C, int n_real, int row_elems,
                                   int pair_window, cudaStream_t s);

// instance-norm apply: y = act(bias + (x-mean)*inv*scale) [+ residual]; x bf16 or fp32, y bf16 or fp32, C % vec == 0
struct CinApplyV {
    const void* x = nullptr; void* y = nullptr; const __nv_bfloat16* residual = nullptr;
    int x_f32 = 0, y_f32 = 0;
    int y_u8 = 0;                      // 3-channel fp32 head only: store trunc(255*y) as uint8 (y_f32 ignored)
    const double* stats = nullptr; const float* params = nullptr;
    long long param_bstride = 0, param_sstride = 0; int scale_off = 0, bias_off = 0;
    const float* weights = nullptr;    // (B,P,2) or null
    int B = 0, P = 0, C = 0, num_styles = 1, act = ACT_NONE; float eps = 1e-5f;
    int l2_hints = 0;                  // bulk kernel: bit 0 = the skip tensor is read evict_first (dead after this pass)
};
cudaError_t launch_cin_apply_v(const CinApplyV& p, cudaStream_t s);

}  // namespace rst
