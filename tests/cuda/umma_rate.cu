// Micro-benchmark: sustained tcgen05.mma (cta_group::1, M=128, K=16, bf16, operands in shared memory) cost per
// instruction as a function of N and of the A-operand row width / descriptor geometry of the halo GEMM.
// The issue loop is fully unrolled with uniform-datapath descriptor arithmetic (3 instructions per MMA), so the
// numbers reflect the tensor pipe + shared-memory operand fetch, not the issuing thread.
// (The first version of this benchmark built descriptors with runtime integer arithmetic in one thread and measured
//  155-215 cycles per MMA independent of N: that is what exposed the issue-bound MMA loop, see profiles/.)
#include <cstdio>
#include <cstdlib>
#include <cuda_bf16.h>

#include "../../realtime_style_transfer_b200/csrc/umma.cuh"

using namespace rst::umma;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(2); } } while (0)

// MODE 0: canonical SW128 (128-byte rows, SBO 1024)     MODE 1: halo SW128, SBO 1280, start shifted per tap
// MODE 2: halo SW64 (64-byte rows, SBO 16*64), shifted   MODE 3: canonical SW128, 2-CTA not used; B advanced per MMA
template <int N, int MODE>
__global__ void __launch_bounds__(128, 1) rate_kernel(int rounds, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    for (int i = threadIdx.x; i < 150 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (threadIdx.x < 32) tmem_alloc(&tmem_slot, 256);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (threadIdx.x < 32) {
        const uint32_t tmem = __shfl_sync(0xffffffffu, tmem_slot, 0);
        const uint32_t idesc = make_idesc_bf16(128, N);
        const uint32_t a16 = __shfl_sync(0xffffffffu, smem_u32(smem) >> 4, 0);
        const uint32_t b16 = __shfl_sync(0xffffffffu, smem_u32(smem + 48 * 1024) >> 4, 0);
        const uint64_t da_c = MODE == 0 || MODE == 3 ? make_smem_desc(0, 16, 1024, SWIZZLE_128B)
                              : MODE == 1 ? make_smem_desc(0, 16, 1280, SWIZZLE_128B) : make_smem_desc(0, 16, 16 * 64, SWIZZLE_64B);
        const uint64_t db_c = make_smem_desc(0, 16, 1024, SWIZZLE_128B);
        const bool leader = elect_one();
        uint32_t ph = 0;
        long long t0 = 0;
        for (int r = 0; r < rounds + 1; ++r) {
            if (r == 1) t0 = clock64();
#pragma unroll
            for (int i = 0; i < 72; ++i) {
                const int tap = (i / 4) % 9, k = i % 4;
                const uint32_t aoff = MODE == 0 || MODE == 3 ? k * 32 : MODE == 1 ? ((tap % 3) * 10 + tap / 3) * 128 + k * 32
                                                                                   : ((tap % 3) * 16 + tap / 3) * 64 + (k & 1) * 32;
                const uint32_t boff = (MODE == 3 ? ((i / 4) % 4) * N * 128 : 0) + k * 32;
                const uint64_t da = da_c | (uint64_t)(a16 + (aoff >> 4));
                const uint64_t db = db_c | (uint64_t)(b16 + (boff >> 4));
                if (leader) mma_f16_ss(tmem, da, db, idesc, i != 0);
            }
            if (leader) mma_commit(&bar);
            mbar_wait(&bar, ph); ph ^= 1;
        }
        const long long t1 = clock64();
        if (blockIdx.x == 0 && leader) out[0] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tmem_slot, 256);
}

template <int N, int MODE>
void run(int ctas, long long* d, const char* name) {
    size_t smem = 200 * 1024;
    CK(cudaFuncSetAttribute(rate_kernel<N, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int rounds = 20;
    rate_kernel<N, MODE><<<ctas, 128, smem>>>(rounds, d);
    CK(cudaDeviceSynchronize());
    long long cyc; CK(cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost));
    printf("ctas=%3d %-30s N=%3d : %6.1f cycles/MMA (tensor ideal %3d, smem@128B/cyc %3d)\n", ctas, name, N,
           (double)cyc / (72.0 * rounds), 128 * N / 256, (MODE == 2 ? 4096 + N * 32 : 4096 + N * 32) / 128);
}

int main() {
    long long* d; CK(cudaMalloc(&d, 8));
    for (int ctas : {1, 148}) {
        run<16, 0>(ctas, d, "canonical SW128"); run<32, 0>(ctas, d, "canonical SW128"); run<64, 0>(ctas, d, "canonical SW128");
        run<128, 0>(ctas, d, "canonical SW128"); run<256, 0>(ctas, d, "canonical SW128");
        run<32, 1>(ctas, d, "halo SW128 SBO1280 shifted"); run<128, 1>(ctas, d, "halo SW128 SBO1280 shifted");
        run<16, 2>(ctas, d, "halo SW64 rows shifted"); run<32, 2>(ctas, d, "halo SW64 rows shifted"); run<64, 2>(ctas, d, "halo SW64 rows shifted");
        run<64, 3>(ctas, d, "canonical, B block per 4 MMAs"); run<128, 3>(ctas, d, "canonical, B block per 4 MMAs");
    }
    return 0;
}
