// Internal declarations shared by the translation units of librst_sm100.so.
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdlib.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/rst_b200.h"

namespace rst {

enum Act : int { ACT_NONE = 0, ACT_RELU = 1, ACT_HSWISH = 2, ACT_HSIGMOID = 3, ACT_SIGMOID = 4 };

// Environment switches.  ab_env(): A/B selection between kernels that compute the SAME result; only ever read at plan / commit
// time, never on a launch path.  exp_env(): bisecting / timing experiments (some give WRONG results); they exist only in builds
// with -DRST_EXPERIMENTS and read as "unset" in the release library.
inline const char* ab_env(const char* name) { return getenv(name); }
inline const char* exp_env(const char* name) {
#ifdef RST_EXPERIMENTS
    return getenv(name);
#else
    (void)name;
    return nullptr;
#endif
}

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// ---- programmatic dependent launch ----------------------------------------------------------------------------------------
// The kernels of the inference forward are launched with cudaLaunchAttributeProgrammaticStreamSerialization.  Every such kernel
// executes pdl_wait() in every thread that reads or writes anything a predecessor touches, before it does so (the wait returns
// when the predecessor grid has completed and its memory is visible; without the launch attribute it is a no-op).  The
// HBM-bound pass kernels (norm apply, input packing) also execute pdl_trigger() right after their wait: the convolution that
// follows is then staged while the pass runs, and its CTAs allocate TMEM and pull their weights into shared memory as soon as
// an SM drains.  A convolution triggers only when its successor is another convolution (HaloGemmParams::pdl_trigger: stem ->
// contract_0 -> contract_1 -> residual_block_0/conv0).  A convolution followed by a pass does NOT: pass CTAs parked in their
// wait while the convolution still runs cost 3 - 6 % of the step on B200, even when they sit on other SMs
// (profiles/r02_04_pdl_and_bulk_norm.md); the pass then starts when the convolution completes (the implicit trigger), which still
// hides the launch latency.  Triggering only after the wait keeps the look-ahead at one kernel, so "predecessor complete" stays
// transitive along the chain.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

inline bool pdl_enabled() {
    static const bool on = !(ab_env("RST_PDL") && atoi(ab_env("RST_PDL")) == 0);
    return on;
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(bool dependent, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = dependent && pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#endif

// TF 'SAME' padding for one spatial dim: returns output size, writes pad_before.
inline int tf_same(int size, int k, int s, int* pad_before) {
    int out = (size + s - 1) / s;
    int total = (out - 1) * s + k - size;
    if (total < 0) total = 0;
    *pad_before = total / 2;
    return out;
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a per-DEVICE setting: remember the largest size configured on each
// device so that a process that drives several GPUs (or switches device) configures every one of them.
struct SmemAttrCache { size_t bytes[64] = {}; };
template <typename F>
inline cudaError_t ensure_dynamic_smem(F kernel, size_t bytes, SmemAttrCache& cache) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    if (cache.bytes[dev] >= bytes && bytes > 0) return cudaSuccess;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess) cache.bytes[dev] = bytes;
    return e;
}

// ---------------------------------------------------------------------------------------------
// Generic fp32 direct convolution (CUDA cores).  Correctness anchor + the RST_PRECISION_FP32 path.
// ---------------------------------------------------------------------------------------------
struct ConvF32 {
    const float* x = nullptr;        // (B,Hi,Wi,Ci)
    const float* w = nullptr;        // indexed tap*w_tap + ci*w_ci + co*w_co
    const float* bias = nullptr;     // (Co) or null
    float* y = nullptr;              // (B,Ho,Wo,Co)
    int B = 0, Hi = 0, Wi = 0, Ci = 0, Ho = 0, Wo = 0, Co = 0;
    int kh = 1, kw = 1, stride = 1, pad_t = 0, pad_l = 0;
    int transposed = 0;              // 1: input-gradient form, iy = (oy + pad_t - ky)/stride when exact
    long long w_tap = 0, w_ci = 0, w_co = 0;
    float in_scale = 1.f, in_shift = 0.f;   // applied to in-bounds inputs only (Keras Rescaling)
    int act1 = ACT_NONE;                    // y = act2(post_scale*act1(acc+bias)+post_shift) + residual
    const float* post_scale = nullptr;
    const float* post_shift = nullptr;
    int act2 = ACT_NONE;
    const float* residual = nullptr;
    int out_tf32 = 0;                       // round the stored value to tf32 (it feeds a tf32 tensor-core conv, which truncates)
    int ph_y = 0, ph_x = 0;                 // set by launch_conv_f32: output parity of one launch of a stride-2 input-gradient conv
};
cudaError_t launch_conv_f32(const ConvF32& p, cudaStream_t s);

struct DepthwiseF32 {
    const float* x = nullptr; const float* w = nullptr;  // w (kh,kw,C)
    float* y = nullptr;
    int B = 0, Hi = 0, Wi = 0, C = 0, Ho = 0, Wo = 0, k = 3, stride = 1, pad_t = 0, pad_l = 0;
    const float* post_scale = nullptr; const float* post_shift = nullptr;
    int act = ACT_NONE;
};
cudaError_t launch_depthwise_f32(const DepthwiseF32& p, cudaStream_t s);

// per-(n,c) moments over H*W: stats (B,C,2) doubles = [sum, sumsq]; must be zeroed first.
cudaError_t launch_moments_f32(const float* x, double* stats, int B, int P, int C, cudaStream_t s);
cudaError_t launch_zero_f64(double* p, long long n, cudaStream_t s);
// out (B,C) = mean over pixels from stats
cudaError_t launch_stats_to_mean(const double* stats, float* mean, int B, int P, int C, cudaStream_t s);

// Conditional instance norm apply (styleTransfer.py:57-71):
//   y = act(bias + (x*inv - mean*inv)*scale) [+ residual];  S==2 blends scale/bias per pixel.
struct CinApply {
    const void* x = nullptr;           // fp32 or bf16 (x_bf16)
    void* y = nullptr;                 // fp32 or bf16 (y_bf16)
    const void* residual = nullptr;    // same dtype as y, or null
    const double* stats = nullptr;     // (B,C,2) [sum,sumsq]
    const float* params = nullptr;     // element (b,s,j) at params[b*param_bstride + s*param_sstride + j]
    long long param_bstride = 0, param_sstride = 0;
    int scale_off = 0, bias_off = 0;   // offsets j of the scale / bias vectors inside a style's params
    const float* weights = nullptr;    // (B,H,W,2) blend weights (S==2) or null
    int B = 0, P = 0, C = 0, num_styles = 1, act = ACT_NONE;
    int x_bf16 = 0, y_bf16 = 0;
    float eps = 1e-5f;
};
cudaError_t launch_cin_apply(const CinApply& p, cudaStream_t s);

// style-weight pyramid (styleTransfer.py:297-303, :335-345)
cudaError_t launch_weights_concat(const float* w_in, float* w_out, long long pixels, int sm1, cudaStream_t s);
// two styles, H % 4 == 0, W % 4 == 0: concat level and the first two pooled levels in one kernel
cudaError_t launch_weights_pyramid3(const float* w_in, float* l0, float* l1, float* l2, int B, int H, int W, cudaStream_t s);
cudaError_t launch_avgpool2_f32(const float* x, float* y, int B, int Hi, int Wi, int C, cudaStream_t s);
cudaError_t launch_maxpool2_f32(const float* x, float* y, int B, int Hi, int Wi, int C, cudaStream_t s);
cudaError_t launch_scale_channels(const float* x, const float* z, float* y, int B, int P, int C, cudaStream_t s);
cudaError_t launch_apply_style_weights(const float* w, const float* params, float* out, int B, long long P, int F,
                                       cudaStream_t s);
cudaError_t launch_gram_f32(const float* x, float* g, int B, int P, int C, cudaStream_t s);

// ---------------------------------------------------------------------------------------------
// bf16 tensor-core path (conv_umma.cu)
// ---------------------------------------------------------------------------------------------
bool umma_init(std::string* err);   // resolves cuTensorMapEncodeTiled through the runtime

}  // namespace rst
