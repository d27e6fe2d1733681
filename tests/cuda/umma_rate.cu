// Micro-benchmark: sustained tcgen05.mma (cta_group::1, M=128, K=16, bf16) issue-to-completion cost as a function
// of N and of the A-operand descriptor geometry used by the halo GEMM.  One CTA, operands are zeros in smem.
// Usage: umma_rate   (prints a table of cycles per MMA)
#include <cstdio>
#include <cstdlib>
#include <cuda_bf16.h>

#include "../../realtime_style_transfer_b200/csrc/umma.cuh"

using namespace rst::umma;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(2); } } while (0)

struct Cfg { int N; int mode; int nmma; int rounds; int ctas; };
// mode 0: canonical SW128, SBO 1024, A start fixed (+k*32)        mode 1: halo SW128, SBO 1280, start shifted per tap
// mode 2: halo SW64 (64-byte rows), SBO 16*64, shifted per tap    mode 3: canonical but B also re-read from a different block each time
__global__ void __launch_bounds__(128, 1) rate_kernel(Cfg c, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (threadIdx.x < 32) tmem_alloc(&tmem_slot, 256);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = make_idesc_bf16(128, c.N);
        const uint32_t a_base = smem_u32(smem), b_base = smem_u32(smem + 64 * 1024);
        uint32_t ph = 0;
        long long t0 = 0, t1 = 0;
        for (int r = 0; r < c.rounds + 1; ++r) {
            if (r == 1) t0 = clock64();
            for (int i = 0; i < c.nmma; ++i) {
                uint64_t da, db;
                const int tap = (i / 4) % 9, k = i % 4;
                if (c.mode == 0 || c.mode == 3) da = make_smem_desc(a_base + k * 32, 16, 1024, SWIZZLE_128B);
                else if (c.mode == 1) da = make_smem_desc(a_base + ((tap % 3) * 10 + tap / 3) * 128 + k * 32, 16, 1280, SWIZZLE_128B);
                else da = make_smem_desc(a_base + ((tap % 3) * 16 + tap / 3) * 64 + (k & 1) * 32, 16, 16 * 64, SWIZZLE_64B);
                const int blk = c.mode == 3 ? (i / 4) % 4 : 0;
                db = make_smem_desc(b_base + blk * c.N * 128 + k * 32, 16, 1024, SWIZZLE_128B);
                mma_f16_ss(tmem, da, db, idesc, i != 0);
            }
            mma_commit(&bar);
            mbar_wait(&bar, ph); ph ^= 1;
        }
        t1 = clock64();
        if (blockIdx.x == 0) out[0] = (t1 - t0);
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tmem, 256);
}

int main() {
    long long* d; CK(cudaMalloc(&d, 8));
    size_t smem = 162 * 1024;
    CK(cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const char* names[4] = {"canonical SW128 SBO1024", "halo SW128 SBO1280 shifted", "halo SW64 rows, shifted", "canonical, 4 B blocks"};
    for (int ctas : {1, 148})
        for (int mode = 0; mode < 4; ++mode)
            for (int N : {16, 32, 64, 128, 256}) {
                Cfg c{N, mode, 72, 20, ctas};
                rate_kernel<<<ctas, 128, smem>>>(c, d);
                CK(cudaDeviceSynchronize());
                long long cyc; CK(cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost));
                printf("ctas=%3d mode=%d (%-28s) N=%3d : %7.1f cycles/MMA  (ideal %d)\n", ctas, mode, names[mode], N,
                       (double)cyc / (c.nmma * c.rounds), 128 * N / 256 < 1 ? 1 : 128 * N / 256);
            }
    return 0;
}
