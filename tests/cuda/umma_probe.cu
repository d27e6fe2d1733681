// Hardware probe for the tcgen05 implicit-GEMM convolution building blocks (run on a B200 via gpurun):
//   v1  canonical SW128 K-major A tile, one shifted TMA load per filter tap (TMA zero-fill = SAME padding)
//   v2a halo tile loaded once, per-tap A descriptors with a shifted (non-1024B-aligned) start, base_offset = 0
//   v2b same, base_offset = (start >> 7) & 7
//   v3  halo tile in the no-swizzle "interleaved" K-major layout (5-D TMA box with a 16-byte inner dimension)
// Each variant computes one 8x16-pixel x 128-channel output tile of a 3x3 conv over 64 input channels and is
// compared with a CPU loop.  Usage: umma_probe <variant> [h0 w0]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_bf16.h>

#include "../../realtime_style_transfer_b200/csrc/umma.cuh"

using namespace rst::umma;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2); } } while (0)

constexpr int TH = 8, TW = 16, C = 64, CO = 128;
constexpr int HALO_H = TH + 2, HALO_W = TW + 2;

struct Params { float* D; int n, h0, w0, variant; };

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, Params p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;                 // up to 23040 B (halo) ; 16384 B (canonical)
    uint8_t* sB = smem + 24576;         // 16384 B, 1024-aligned
    __shared__ uint64_t full_bar, mma_bar;
    __shared__ uint32_t tmem_base_s;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        mbar_init(&full_bar, 1);
        mbar_init(&mma_bar, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(&tmem_base_s, 128);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    const uint32_t idesc = make_idesc_bf16(128, 128);

    if (threadIdx.x == 0) {
        uint32_t ph_full = 0, ph_mma = 0;
        if (p.variant != 1) {   // halo tile once
            mbar_expect_tx(&full_bar, HALO_H * HALO_W * 128);
            if (p.variant == 3) tma_load_5d(sA, &tmA, &full_bar, 0, p.h0 - 1, p.w0 - 1, p.n, 0);
            else tma_load_4d(sA, &tmA, &full_bar, 0, p.h0 - 1, p.w0 - 1, p.n);
            mbar_wait(&full_bar, ph_full); ph_full ^= 1;
        }
        for (int tap = 0; tap < 9; ++tap) {
            const int dy = tap / 3, dx = tap % 3;
            uint32_t bytes = CO * 128;
            if (p.variant == 1) bytes += TH * TW * 128;
            mbar_expect_tx(&full_bar, bytes);
            if (p.variant == 1) tma_load_4d(sA, &tmA, &full_bar, 0, p.h0 + dy - 1, p.w0 + dx - 1, p.n);
            tma_load_2d(sB, &tmB, &full_bar, 0, tap * CO);
            mbar_wait(&full_bar, ph_full); ph_full ^= 1;
            tc_fence_after();
            for (int k = 0; k < 4; ++k) {
                uint64_t da, db;
                db = make_smem_desc(smem_u32(sB) + k * 32, 16, 1024, SWIZZLE_128B);
                if (p.variant == 1) {
                    da = make_smem_desc(smem_u32(sA) + k * 32, 16, 1024, SWIZZLE_128B);
                } else if (p.variant == 3) {
                    uint32_t start = smem_u32(sA) + (2 * k) * (HALO_H * HALO_W * 16) + (dx * HALO_H + dy) * 16;
                    da = make_smem_desc(start, HALO_H * HALO_W * 16, HALO_H * 16, SWIZZLE_NONE);
                } else {
                    uint32_t start = smem_u32(sA) + (dx * HALO_H + dy) * 128 + k * 32;
                    uint32_t bo = p.variant == 22 ? ((start >> 7) & 7) : 0;
                    da = make_smem_desc(start, 16, HALO_H * 128, SWIZZLE_128B, bo);
                }
                mma_f16_ss(tmem, da, db, idesc, (tap | k) != 0);
            }
            mma_commit(&mma_bar);
            mbar_wait(&mma_bar, ph_mma); ph_mma ^= 1;
        }
    }
    __syncthreads();
    tc_fence_after();
    // epilogue: warp q owns TMEM lanes 32q..32q+31 = GEMM rows
    const int row = threadIdx.x;
    for (int c0 = 0; c0 < CO; c0 += 32) {
        float v[32];
        tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
        tmem_ld_wait();
        for (int j = 0; j < 32; ++j) p.D[row * CO + c0 + j] = v[j];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 128);
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
    int variant = argc > 1 ? atoi(argv[1]) : 1;
    const int N = 2, H = 24, W = 48;
    int h0 = argc > 3 ? atoi(argv[2]) : 0, w0 = argc > 3 ? atoi(argv[3]) : 0, n = 1;
    EncodeFn encode = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres));
    if (!encode) { printf("no cuTensorMapEncodeTiled\n"); return 2; }

    std::vector<__nv_bfloat16> x((size_t)N * H * W * C), wt((size_t)9 * CO * C);
    std::vector<float> xf(x.size()), wf(wt.size());
    srand(123);
    for (size_t i = 0; i < x.size(); ++i) { float v = (rand() % 2001 - 1000) / 1000.f; x[i] = __float2bfloat16(v); xf[i] = __bfloat162float(x[i]); }
    for (size_t i = 0; i < wt.size(); ++i) { float v = (rand() % 2001 - 1000) / 4000.f; wt[i] = __float2bfloat16(v); wf[i] = __bfloat162float(wt[i]); }
    __nv_bfloat16 *dx, *dw; float* dD;
    CK(cudaMalloc(&dx, x.size() * 2)); CK(cudaMalloc(&dw, wt.size() * 2)); CK(cudaMalloc(&dD, 128 * CO * 4));
    CK(cudaMemcpy(dx, x.data(), x.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dw, wt.data(), wt.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0xFF, 128 * CO * 4));

    CUtensorMap tmA, tmB;
    CUresult r;
    if (variant == 3) {
        cuuint64_t gdim[5] = {8, (cuuint64_t)H, (cuuint64_t)W, (cuuint64_t)N, C / 8};
        cuuint64_t gstr[4] = {(cuuint64_t)W * C * 2, (cuuint64_t)C * 2, (cuuint64_t)H * W * C * 2, 16};
        cuuint32_t box[5] = {8, HALO_H, HALO_W, 1, 8};
        cuuint32_t es[5] = {1, 1, 1, 1, 1};
        r = encode(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, dx, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
        cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)H, (cuuint64_t)W, (cuuint64_t)N};
        cuuint64_t gstr[3] = {(cuuint64_t)W * C * 2, (cuuint64_t)C * 2, (cuuint64_t)H * W * C * 2};
        cuuint32_t box[4] = {64, (cuuint32_t)(variant == 1 ? TH : HALO_H), (cuuint32_t)(variant == 1 ? TW : HALO_W), 1};
        cuuint32_t es[4] = {1, 1, 1, 1};
        r = encode(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, dx, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r != CUDA_SUCCESS) { printf("encode A failed: %d\n", (int)r); return 2; }
    {
        cuuint64_t gdim[2] = {(cuuint64_t)C, (cuuint64_t)9 * CO};
        cuuint64_t gstr[1] = {(cuuint64_t)C * 2};
        cuuint32_t box[2] = {64, CO};
        cuuint32_t es[2] = {1, 1};
        r = encode(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dw, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode B failed: %d\n", (int)r); return 2; }
    }
    Params p{dD, n, h0, w0, variant};
    size_t smem = 24576 + 16384 + 1024;
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    probe_kernel<<<1, 128, smem>>>(tmA, tmB, p);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<float> D(128 * CO);
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0, maxref = 0;
    int bad = 0;
    for (int w = 0; w < TW; ++w) for (int h = 0; h < TH; ++h) for (int co = 0; co < CO; ++co) {
        double acc = 0;
        for (int dy = 0; dy < 3; ++dy) for (int dxx = 0; dxx < 3; ++dxx) {
            int iy = h0 + h + dy - 1, ix = w0 + w + dxx - 1;
            if (iy < 0 || iy >= H || ix < 0 || ix >= W) continue;
            const float* xp = &xf[(((size_t)n * H + iy) * W + ix) * C];
            const float* wp = &wf[((size_t)(dy * 3 + dxx) * CO + co) * C];
            for (int c = 0; c < C; ++c) acc += (double)xp[c] * wp[c];
        }
        double got = D[(w * TH + h) * CO + co];
        double err = fabs(got - acc);
        if (!(err <= 1e-2)) ++bad;
        if (err > maxerr || err != err) maxerr = err;
        if (fabs(acc) > maxref) maxref = fabs(acc);
    }
    printf("variant %d tile(h0=%d,w0=%d): max_err=%.4g max_ref=%.4g bad=%d -> %s\n", variant, h0, w0, maxerr, maxref, bad,
           bad == 0 ? "PASS" : "FAIL");
    return bad == 0 ? 0 : 1;
}
