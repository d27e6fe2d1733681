// tf32 tensor-core 3x3 convolutions on fp32 NHWC tensors, built on the halo GEMM kernel (halo_gemm.cu): same TMA halo loads,
// same shifted shared-memory descriptors, tcgen05.mma kind::tf32 (K = 8 fp32 per MMA = the same 32 bytes as 16 bf16).
// Used by the VGG16 loss model (loss.cu) for every layer with at least 64 input channels, forward and input gradient.
#include <cstring>

#include "halo_gemm.cuh"

namespace rst {

Tf32Conv3x3::~Tf32Conv3x3() {
    if (w_packed) cudaFree(w_packed);
    if (bias) cudaFree(bias);
}

// packed[j][blk][n][r]: K-step ks = 4*blk + r/8, element e = r%8 -> channel group g = ks/36, tap = (ks%36)/4,
// input channel g*32 + ((ks%36)%4)*8 + e; output channel j*nb + n.  Rounded to tf32 (nearest).
// Split convs: the activation side is [x_hi | x_lo | x_hi], the weight side [w_hi | w_hi | w_lo].
__global__ void tf32_pack_weights_kernel(const float* __restrict__ k, float* __restrict__ out, int ci_layer, int co_layer,
                                         int input_gradient, int nb, int blocks, int ci_real, long long total) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int r = (int)(idx % 32);
    long long t = idx / 32;
    const int n = (int)(t % nb);
    t /= nb;
    const int blk = (int)(t % blocks), j = (int)(t / blocks);
    const int ks = blk * 4 + r / 8, e = r % 8;
    const int g = ks / 36, l = ks % 36, tap = l / 4;
    const int cin_eff = g * 32 + (l % 4) * 8 + e, cout = j * nb + n;
    const int part = cin_eff / ci_real, cin = cin_eff - part * ci_real;       // split convs: parts [w_hi | w_hi | w_lo]
    const float w = !input_gradient ? k[((size_t)tap * ci_layer + cin) * co_layer + cout]
                                    : k[((size_t)(8 - tap) * ci_layer + cout) * co_layer + cin];
    uint32_t rr;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(rr) : "f"(w));
    float v = __uint_as_float(rr);
    if (part == 2) { asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(rr) : "f"(w - v)); v = __uint_as_float(rr); }
    out[idx] = v;
}

// x (P, c) -> [tf32(x) | tf32(x - tf32(x))] (P, 2c), four channels per thread; the conv reads [x_hi | x_lo | x_hi] out of it
// relu_mask (optional): x is a gradient that still has to be masked by the ReLU of its layer, v = mask > 0 ? x : 0
__global__ void tf32_split_expand_kernel(const float* __restrict__ x, float* __restrict__ out, int c4, long long total4,
                                         const float* __restrict__ relu_mask) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total4) return;
    const long long pix = i / c4;
    const int ch4 = (int)(i - pix * c4);
    float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
    if (relu_mask) {
        const float4 m = __ldg(reinterpret_cast<const float4*>(relu_mask) + i);
        v.x = m.x > 0.f ? v.x : 0.f; v.y = m.y > 0.f ? v.y : 0.f; v.z = m.z > 0.f ? v.z : 0.f; v.w = m.w > 0.f ? v.w : 0.f;
    }
    float4 hi = make_float4(umma::round_tf32(v.x), umma::round_tf32(v.y), umma::round_tf32(v.z), umma::round_tf32(v.w));
    float4 lo = make_float4(umma::round_tf32(v.x - hi.x), umma::round_tf32(v.y - hi.y), umma::round_tf32(v.z - hi.z), umma::round_tf32(v.w - hi.w));
    float4* o = reinterpret_cast<float4*>(out) + pix * 2 * c4 + ch4;
    o[0] = hi; o[c4] = lo;
}

// x_lo = rna_tf32(x - trunc_tf32(x)): the compensation term when the MMA reads the RAW tensor as the hi part (kind::tf32 uses
// the upper 19 bits of an fp32 operand).  Used by the tensor-core weight gradient (wgrad_tf32.cu).
__global__ void tf32_lo_kernel(const float* __restrict__ in, float* __restrict__ out, long long n4) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    const float4 v = __ldg(reinterpret_cast<const float4*>(in) + i);
    auto lo = [](float f) { return umma::round_tf32(f - __uint_as_float(__float_as_uint(f) & 0xFFFFE000u)); };
    reinterpret_cast<float4*>(out)[i] = make_float4(lo(v.x), lo(v.y), lo(v.z), lo(v.w));
}
cudaError_t launch_tf32_lo(const float* in, float* out, long long n, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    if (n % 4 != 0) return cudaErrorInvalidValue;
    tf32_lo_kernel<<<(unsigned)((n / 4 + 255) / 256), 256, 0, s>>>(in, out, n / 4);
    return cudaGetLastError();
}

bool Tf32Conv3x3::setup_shape(int ci_layer_, int co_layer_, bool relu_, bool input_gradient_, std::string* err, bool split_) {
    ci_layer = ci_layer_; co_layer = co_layer_; input_gradient = input_gradient_; split = split_;
    ci = (input_gradient ? co_layer : ci_layer) * (split ? 3 : 1);
    co = input_gradient ? ci_layer : co_layer;
    relu = relu_;
    // blocks of 32 output channels exist for the forward conv with ReLU only (the 32-filter bottleneck of the fp32 inference path)
    if (ci % 32 != 0 || (co % 64 != 0 && !(co == 32 && relu && !input_gradient))) {
        if (err) *err = "tf32 conv: needs Cin % 32 == 0 and Cout % 64 == 0 (or Cout == 32 for a forward conv with ReLU)";
        return false;
    }
    nb = co % 128 == 0 ? 128 : co % 64 == 0 ? 64 : 32;
    nblk = co / nb;
    const int n_groups = ci / 32, blocks = n_groups * 9;
    if (w_packed) { cudaFree(w_packed); w_packed = nullptr; }
    if (bias) { cudaFree(bias); bias = nullptr; }
    ext_bias = nullptr;
    cudaError_t e = cudaMalloc(&w_packed, (size_t)nblk * blocks * nb * 32 * sizeof(float));
    if (e != cudaSuccess) {
        if (err) *err = std::string("tf32 conv: ") + cudaGetErrorString(e);
        return false;
    }
    launch = HaloGemmLaunch();
    launch.N = nb; launch.row_bytes = 128; launch.epi = EPI_NHWC; launch.sched = SCH_C3;
    launch.mode = HALO_MODE_F32 | HALO_MODE_TF32 | (relu ? HALO_MODE_RELU : 0);
    p = HaloGemmParams();
    p.n_groups = n_groups;
    p.out_C = co; p.stats_c = co;
    if (!halo_gemm_plan(&launch, &p, err)) return false;
    tmB.resize(nblk);
    for (int j = 0; j < nblk; ++j)
        if (!encode_weight_map(&tmB[j], w_packed + (size_t)j * blocks * nb * 32, blocks, nb, err)) return false;
    inputs.clear();
    return true;
}

cudaError_t Tf32Conv3x3::repack(const float* d_kernel, const float* d_bias, cudaStream_t s) {
    const int blocks = (ci / 32) * 9;
    const long long total = (long long)nblk * blocks * nb * 32;
    tf32_pack_weights_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(d_kernel, w_packed, ci_layer, co_layer,
                                                                            input_gradient ? 1 : 0, nb, blocks, split ? ci / 3 : ci, total);
    ext_bias = d_bias;
    return cudaGetLastError();
}

bool Tf32Conv3x3::setup(int ci_layer_, int co_layer_, const float* k, const float* bias_host, bool relu_, bool input_gradient_,
                        std::string* err, bool split_) {
    if (!setup_shape(ci_layer_, co_layer_, relu_, input_gradient_, err, split_)) return false;
    float* d_k = nullptr;
    const size_t kelems = (size_t)9 * ci_layer_ * co_layer_;
    cudaError_t e = cudaMalloc(&d_k, kelems * sizeof(float));
    if (e == cudaSuccess) e = cudaMemcpy(d_k, k, kelems * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess && bias_host) {
        e = cudaMalloc(&bias, co * sizeof(float));
        if (e == cudaSuccess) e = cudaMemcpy(bias, bias_host, co * sizeof(float), cudaMemcpyHostToDevice);
    }
    if (e == cudaSuccess) e = repack(d_k, bias, nullptr);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (d_k) cudaFree(d_k);
    if (e != cudaSuccess) {
        if (err) *err = std::string("tf32 conv: ") + cudaGetErrorString(e);
        return false;
    }
    return true;
}

// x: (B, H, W, ci) for a plain conv; for a split conv the expanded tensor (B, H, W, 2 ci / 3) = [x_hi | x_lo]
cudaError_t Tf32Conv3x3::run(const float* x, float* y, int B, int H, int W, int num_sms, cudaStream_t s, std::string* err) {
    const CUtensorMap* tmA = nullptr;
    for (auto& b : inputs)
        if (b.x == x && b.B == B && b.H == H && b.W == W) { tmA = &b.tm; break; }
    const int c_stored = split ? ci / 3 * 2 : ci;
    if (!tmA) {
        BoundInput b{x, B, H, W, {}};
        // the activation map addresses bytes: an fp32 pixel with c channels is a row of 2*c bf16-sized elements
        if (!encode_halo_map(&b.tm, x, B, H, W, 2 * c_stored, 64, p.halo_h, p.halo_w, err)) return cudaErrorInvalidValue;
        if (inputs.size() >= 8) inputs.erase(inputs.begin());
        inputs.push_back(b);
        tmA = &inputs.back().tm;
    }
    HaloGemmParams q = p;
    q.B = B; q.H = H; q.WRU = W;
    q.tiles_h = ceil_div(H, 8); q.tiles_w = ceil_div(W, 16);
    q.out_H = H; q.out_W = W; q.out_C = co;
    q.y_f32 = 1; q.stats = nullptr;
    q.store_exact = split ? 1 : 0;
    q.a_wrap = split ? c_stored / 32 : 0;
    for (int j = 0; j < nblk; ++j) {
        q.y = y + (size_t)j * nb;
        q.bias = ext_bias ? ext_bias + (size_t)j * nb : nullptr;
        cudaError_t e = launch_halo_gemm(launch, *tmA, tmB[j], q, num_sms, s);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

cudaError_t Tf32Conv3x3::run_split(const float* x, float* scratch, float* y, int B, int H, int W, int num_sms, cudaStream_t s,
                                   std::string* err, const float* relu_mask) {
    if (!split) return relu_mask ? cudaErrorInvalidValue : run(x, y, B, H, W, num_sms, s, err);   // the mask rides on the expansion
    const int c = ci / 3;
    const long long total4 = (long long)B * H * W * c / 4;
    if (total4 == 0) return cudaSuccess;
    tf32_split_expand_kernel<<<(unsigned)((total4 + 255) / 256), 256, 0, s>>>(x, scratch, c / 4, total4, relu_mask);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    return run(scratch, y, B, H, W, num_sms, s, err);
}

}  // namespace rst
