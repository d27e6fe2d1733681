// tcgen05 implicit-GEMM 3x3 stride-1 'same' convolution, NHWC bf16 in / bf16 out, fp32 accumulate in TMEM.
//
// GEMM view per CTA tile: M = 128 output pixels (an 8-row x 16-column patch, row index r = w*8 + h),
// N = COUT, K = 9 taps x CIN.  The A operand is NOT materialised per tap: the (8+2)x(16+2) input halo
// patch of one 64-channel half is TMA-loaded ONCE into shared memory (SWIZZLE_128B, one 128-byte row per
// pixel, OOB pixels zero-filled = 'same' padding) and each tap's A tile is the same bytes viewed through a
// shared-memory descriptor whose start address is shifted by (dx*10+dy) rows with stride-byte-offset
// 10*128 B between 8-row groups (verified on B200 by tests/cuda/umma_probe.cu, variant 2).
// Epilogue (4 warps, TMEM -> registers): + bias, ReLU, round to bf16, per-(sample,channel) sum / sum of
// squares for the instance norm that follows (warp butterfly, then fp64 atomics), 16-byte stores.
#pragma once

#include "rst_internal.cuh"
#include "umma.cuh"

namespace rst {

struct ConvUmmaParams {
    __nv_bfloat16* y;        // (B,H,W,COUT) raw conv output after bias+act
    const float* bias;       // (COUT)
    double* stats;           // (B,COUT,2) [sum, sumsq] accumulated with atomics, or null
    int B, H, W;
    int nhalf;               // CIN / 64
    int tiles_h, tiles_w;
    int relu;
};

constexpr int kUmmaTH = 8, kUmmaTW = 16;
constexpr int kHaloH = kUmmaTH + 2, kHaloW = kUmmaTW + 2;
constexpr int kHaloBytes = kHaloH * kHaloW * 128;       // 23040
constexpr int kAStageBytes = 23552;                     // rounded up to a 1024-byte multiple
constexpr int kNumAStages = 3;
constexpr int kUmmaThreads = 224;                       // warps: 0 A-TMA, 1 B-TMA, 2 MMA, 3..6 epilogue

template <int COUT>
struct ConvUmmaSmem {
    static constexpr int kBStageBytes = COUT * 128;
    static constexpr int kNumBStages = COUT >= 128 ? 6 : 8;
    static constexpr int kBytes = kNumAStages * kAStageBytes + kNumBStages * kBStageBytes + 1024 /*align*/ + 1024 /*bars,bias*/;
};

cudaError_t launch_conv3x3_umma(int cout, const CUtensorMap& tmA, const CUtensorMap& tmB, const ConvUmmaParams& p,
                                int num_sms, cudaStream_t s);

// host helpers
bool umma_encode_activation_map(CUtensorMap* out, const void* base, int B, int H, int W, int C, std::string* err);
bool umma_encode_weight_map(CUtensorMap* out, const void* base, int rows, int box_rows, std::string* err);

// bf16 elementwise companions
cudaError_t launch_f32_to_bf16_pad(const float* x, __nv_bfloat16* y, long long pixels, int c_in, int c_out, cudaStream_t s);
cudaError_t launch_bf16_to_f32_slice(const __nv_bfloat16* x, float* y, long long pixels, int c_in, int c_out, cudaStream_t s);

// instance-norm apply on bf16 NHWC with C % 8 == 0: y = act(bias + (x-mean)*inv*scale) [+ residual]
struct CinApplyBf16 {
    const __nv_bfloat16* x; __nv_bfloat16* y; const __nv_bfloat16* residual;
    const double* stats; const float* params; long long param_bstride, param_sstride; int scale_off, bias_off;
    const float* weights;    // (B,P,2) or null
    int B, P, C, num_styles, act; float eps;
};
cudaError_t launch_cin_apply_bf16(const CinApplyBf16& p, cudaStream_t s);

}  // namespace rst
