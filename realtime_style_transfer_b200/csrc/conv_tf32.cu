// tf32 tensor-core 3x3 convolutions on fp32 NHWC tensors, built on the halo GEMM kernel (halo_gemm.cu): same TMA halo loads,
// same shifted shared-memory descriptors, tcgen05.mma kind::tf32 (K = 8 fp32 per MMA = the same 32 bytes as 16 bf16).
// Used by the VGG16 loss model (loss.cu) for every layer with at least 64 input channels, forward and input gradient.
#include <cstring>

#include "halo_gemm.cuh"

namespace rst {

static float to_tf32(float x) {                      // round to nearest, ties away from zero (cvt.rna.tf32.f32)
    uint32_t u;
    std::memcpy(&u, &x, 4);
    if ((u & 0x7f800000u) != 0x7f800000u) u += 0x1000u;
    u &= 0xffffe000u;
    std::memcpy(&x, &u, 4);
    return x;
}

Tf32Conv3x3::~Tf32Conv3x3() {
    if (w_packed) cudaFree(w_packed);
    if (bias) cudaFree(bias);
}

bool Tf32Conv3x3::setup(int ci_layer, int co_layer, const float* k, const float* bias_host, bool relu_, bool input_gradient,
                        std::string* err) {
    ci = input_gradient ? co_layer : ci_layer;
    co = input_gradient ? ci_layer : co_layer;
    relu = relu_;
    if (ci % 32 != 0 || co % 64 != 0) {
        if (err) *err = "tf32 conv: needs Cin % 32 == 0 and Cout % 64 == 0";
        return false;
    }
    nb = co % 128 == 0 ? 128 : 64;
    nblk = co / nb;
    const int n_groups = ci / 32, ksteps = 36, total_ksteps = n_groups * ksteps, blocks = total_ksteps / 4;
    std::vector<float> packed((size_t)nblk * blocks * nb * 32, 0.f);
    for (int j = 0; j < nblk; ++j)
        for (int ks = 0; ks < total_ksteps; ++ks) {
            const int g = ks / ksteps, l = ks % ksteps, tap = l / 4;
            for (int n = 0; n < nb; ++n)
                for (int e = 0; e < 8; ++e) {
                    const int cin = g * 32 + (l % 4) * 8 + e, cout = j * nb + n;
                    const float w = !input_gradient ? k[((size_t)tap * ci_layer + cin) * co_layer + cout]
                                                    : k[((size_t)(8 - tap) * ci_layer + cout) * co_layer + cin];
                    packed[(((size_t)j * blocks + ks / 4) * nb + n) * 32 + (ks % 4) * 8 + e] = to_tf32(w);
                }
        }
    if (w_packed) { cudaFree(w_packed); w_packed = nullptr; }
    if (bias) { cudaFree(bias); bias = nullptr; }
    cudaError_t e = cudaMalloc(&w_packed, packed.size() * sizeof(float));
    if (e == cudaSuccess) e = cudaMemcpy(w_packed, packed.data(), packed.size() * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess && bias_host) {
        e = cudaMalloc(&bias, co * sizeof(float));
        if (e == cudaSuccess) e = cudaMemcpy(bias, bias_host, co * sizeof(float), cudaMemcpyHostToDevice);
    }
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

cudaError_t Tf32Conv3x3::run(const float* x, float* y, int B, int H, int W, int num_sms, cudaStream_t s, std::string* err) {
    const CUtensorMap* tmA = nullptr;
    for (auto& b : inputs)
        if (b.x == x && b.B == B && b.H == H && b.W == W) { tmA = &b.tm; break; }
    if (!tmA) {
        BoundInput b{x, B, H, W, {}};
        // the activation map addresses bytes: an fp32 pixel with ci channels is a row of 2*ci bf16-sized elements
        if (!encode_halo_map(&b.tm, x, B, H, W, 2 * ci, 64, p.halo_h, p.halo_w, err)) return cudaErrorInvalidValue;
        if (inputs.size() >= 8) inputs.erase(inputs.begin());
        inputs.push_back(b);
        tmA = &inputs.back().tm;
    }
    HaloGemmParams q = p;
    q.B = B; q.H = H; q.WRU = W;
    q.tiles_h = ceil_div(H, 8); q.tiles_w = ceil_div(W, 16);
    q.out_H = H; q.out_W = W; q.out_C = co;
    q.y_f32 = 1; q.stats = nullptr;
    for (int j = 0; j < nblk; ++j) {
        q.y = y + (size_t)j * nb;
        q.bias = bias ? bias + (size_t)j * nb : nullptr;
        cudaError_t e = launch_halo_gemm(launch, *tmA, tmB[j], q, num_sms, s);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

}  // namespace rst
