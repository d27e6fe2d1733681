// Kernel launchers used by the training step (train.cu).
#pragma once
#include <string>

#include "rst_internal.cuh"

namespace rst {

// weight gradient of a 'same' Conv2D (HWIO) / Conv2DTranspose (Keras (kh,kw,out,in)).  dw must be zeroed (accumulates).
struct WgradF32 {
    const float* x = nullptr;   // layer input  (B,Hx,Wx,Ci)
    const float* g = nullptr;   // gradient w.r.t. the layer output (B,Hg,Wg,Co)
    float* dw = nullptr;
    int B = 0, Hx = 0, Wx = 0, Ci = 0, Hg = 0, Wg = 0, Co = 0;
    int Hb = 0, Wb = 0;         // base grid: conv -> (Hg,Wg), transposed -> (Hx,Wx)
    int kh = 1, kw = 1, stride = 1, pad_t = 0, pad_l = 0;
    int transposed = 0;
    float in_scale = 1.f, in_shift = 0.f;   // the forward's input rescaling (in-bounds taps only)
};
cudaError_t launch_wgrad_f32(const WgradF32& p, cudaStream_t s);
// tensor-core (kind::tf32, MN-major operands) weight gradient of a 3x3 stride-1 'same' convolution 128 -> 128 (wgrad_tf32.cu);
// split = error-compensated (fp32-level accuracy), needs wgrad_tf32_scratch_floats() floats of scratch
size_t wgrad_tf32_scratch_floats(int B, int H, int W);
// tensor-core Gram matrix (gram_tf32.cu): x (B, P, C) fp32 -> gram (B, C, C) = x^T x / P per sample, C = 64 or a multiple of 128;
// split = error-compensated tf32 (fp32-level accuracy), needs gram_tf32_scratch_floats() floats of scratch
bool gram_tf32_supported(int C);
size_t gram_tf32_scratch_floats(int B, int P, int C);
cudaError_t launch_gram_tf32(const float* x, float* gram, float* lo_scratch, int B, int P, int C, bool split, int num_sms,
                             cudaStream_t s, std::string* err);
cudaError_t launch_wgrad_tf32(const float* x, const float* g, float* dw, float* lo_scratch, int B, int H, int W, bool split,
                              int num_sms, cudaStream_t s, std::string* err);

cudaError_t launch_reduce_over_batch(const double* in, double* out, int B, int C2, cudaStream_t s);
cudaError_t launch_norm_finalize(const double* stats, int G, int C, double count, float eps, const float* scale,
                                 const float* bias, long long gstride, float* mean, float* inv, float* a, float* b,
                                 float* moving_mean, float* moving_var, float momentum, cudaStream_t s);
cudaError_t launch_affine_act(const float* x, float* y, const float* a, const float* b, const float* residual, int B, long long P,
                              int C, int per_sample, int act, cudaStream_t s);
cudaError_t launch_act_bwd(float* g, const float* out, int act, long long n, cudaStream_t s);
cudaError_t launch_norm_bwd_reduce(const float* g, const float* x, const float* mean, const float* inv, double* r, int B, int P,
                                   int C, int per_sample, cudaStream_t s);
cudaError_t launch_norm_bwd_apply(const float* g, const float* x, const float* mean, const float* inv, const float* a,
                                  const double* r, float* gx, int B, long long P, int C, int per_sample, double count,
                                  int accumulate, cudaStream_t s);
cudaError_t launch_affine_act_bwd(float* g, const float* x, const float* a, const float* b, int B, long long P, int C,
                                  int per_sample, int act, cudaStream_t s);
cudaError_t launch_cin_param_grad(const double* r, float* pg, int B, int C, long long ptotal, int off, cudaStream_t s);
cudaError_t launch_gather_f64(const double* in, float* out, int n, int stride, int offset, int accumulate, cudaStream_t s);
cudaError_t launch_add_inplace(float* a, const float* b, long long n, cudaStream_t s);
cudaError_t launch_gap_bwd(const float* g, float* gx, int B, long long P, int C, int accumulate, cudaStream_t s);
cudaError_t launch_scale_channels_bwd(const float* g, const float* z, float* gx, int B, long long P, int C, int accumulate,
                                      cudaStream_t s);
cudaError_t launch_depthwise_dgrad(const DepthwiseF32& p, const float* g, float* gx, int accumulate, cudaStream_t s);
cudaError_t launch_depthwise_wgrad(const DepthwiseF32& p, const float* g, float* dw, cudaStream_t s);
cudaError_t launch_rmsprop(float* w, const float* g, float* rms, float lr, float rho, float eps, long long n, cudaStream_t s);

}  // namespace rst
