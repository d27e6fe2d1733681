// Microbenchmark of the conditional-instance-norm apply pass (launch_cin_apply_v) on bf16 NHWC tensors of the trunk's shape,
// in place, back to back, so that the tensor stays in L2 when it fits.  Usage: norm_pass_bench [C] [P] [first batch size]
// Build (from the repo root, after __graft_entry__.build()):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -cudart shared -I realtime_style_transfer_b200/csrc \
//        -o tests/cuda/norm_pass_bench tests/cuda/norm_pass_bench.cu -L realtime_style_transfer_b200/csrc -lrst_sm100 \
//        -Xlinker -rpath -Xlinker '$ORIGIN/../../realtime_style_transfer_b200/csrc'
// (the binary is git- and gpurun-ignored; drop the line from .gpurunignore to run it on a GPU box).  Environment switches of the library (RST_NORM_BULK, RST_NORM_PPB, ...) apply.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_bf16.h>
#include "halo_gemm.cuh"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

int main(int argc, char** argv) {
    const int C = argc > 1 ? atoi(argv[1]) : 128, P = argc > 2 ? atoi(argv[2]) : 120 * 240, BMIN = argc > 3 ? atoi(argv[3]) : 1, BMAX = 8;
    const size_t elems = (size_t)BMAX * P * C;
    __nv_bfloat16 *x, *r;
    double* stats; float* params;
    CK(cudaMalloc(&x, elems * 2)); CK(cudaMalloc(&r, elems * 2));
    CK(cudaMalloc(&stats, sizeof(double) * 2 * BMAX * C)); CK(cudaMalloc(&params, sizeof(float) * 2 * C));
    CK(cudaMemset(x, 0, elems * 2)); CK(cudaMemset(r, 0, elems * 2));
    std::vector<double> hs(2 * BMAX * C); for (size_t i = 0; i < hs.size(); i += 2) { hs[i] = 0.0; hs[i + 1] = (double)P; }
    std::vector<float> hp(2 * C, 1.0f);
    CK(cudaMemcpy(stats, hs.data(), hs.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(params, hp.data(), hp.size() * 4, cudaMemcpyHostToDevice));
    cudaStream_t s; CK(cudaStreamCreate(&s));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int res = 0; res < 2; ++res)
        for (int B = BMIN; B <= BMAX; B *= 2) {
            rst::CinApplyV p;
            p.x = x; p.y = x; p.residual = res ? r : nullptr;
            p.stats = stats; p.params = params; p.param_bstride = 0; p.param_sstride = 0; p.scale_off = 0; p.bias_off = C;
            p.B = B; p.P = P; p.C = C; p.num_styles = 1; p.act = rst::ACT_RELU;
            const int reps = 50;
            for (int i = 0; i < 5; ++i) CK(rst::launch_cin_apply_v(p, s));
            cudaGraph_t g; cudaGraphExec_t ge;                   // the product replays a captured graph: time the same thing
            CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
            for (int i = 0; i < reps; ++i) CK(rst::launch_cin_apply_v(p, s));
            CK(cudaStreamEndCapture(s, &g));
            CK(cudaGraphInstantiate(&ge, g, 0));
            CK(cudaGraphLaunch(ge, s));
            CK(cudaEventRecord(e0, s));
            CK(cudaGraphLaunch(ge, s));
            CK(cudaEventRecord(e1, s));
            CK(cudaStreamSynchronize(s));
            CK(cudaGraphExecDestroy(ge)); CK(cudaGraphDestroy(g));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            const double us = ms * 1e3 / reps, mb = (double)B * P * C * 2 / 1e6;
            printf("res=%d B=%d tensor %.1f MB  %.2f us/launch  %.2f TB/s (read+write, %d tensors)\n", res, B, mb, us,
                   mb * (2 + res) / us, 2 + res);
        }
    return 0;
}
