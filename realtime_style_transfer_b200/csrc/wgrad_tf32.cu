// Weight gradient of the residual trunk's 3x3 convolutions (128 -> 128, stride 1, 'same') on the tensor cores.
//
//   dW[ky, kx, ci, co] = sum_{n,y,x} X[n, y + ky - 1, x + kx - 1, ci] * G[n, y, x, co]
//
// is a GEMM whose reduction dimension is the PIXEL index, while NHWC tensors are contiguous along the channels: both operands
// are "MN-major".  tcgen05 reads such tiles directly (instruction-descriptor bits 15 / 16): a TMA box of 32 channels x P
// pixels lands as P rows of 128 B, which is the canonical MN-major layout (8 pixel rows per K-step of a kind::tf32 MMA,
// channel groups of 32 one box apart = the descriptor's leading-dimension offset).  For tf32 the only MN-major layout the
// hardware accepts is the 128-byte swizzle in 32-byte chunks over 4 rows: CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B on the TMA
// side, descriptor layout type 1 on the MMA side (the plain 128-byte swizzle silently yields zeros).  The column tap kx is
// a shift of the X descriptor by kx rows (the halo trick on the reduction axis); the row tap ky is a different input row, so
// the three ky run in different CTAs and every CTA keeps 3 accumulators of 128 x 128 fp32 (384 TMEM columns).
//
//   D_kx[co, ci] += G[pixels, co]^T  X[pixels + kx, ci]          A = G (M = co), B = X (N = ci), K = 8 pixels per MMA
//
// fp32-level accuracy (RST_PRECISION_FP32 math): the hardware reads only the upper 19 bits of an operand, so each product is
// issued as three MMAs, hi*hi + hi*lo + lo*hi, with lo = rna_tf32(v - trunc_tf32(v)) written by a small pre-pass; the
// accumulator (which truncates on every add) is flushed to the fp32 gradient buffer with atomics every kWgFlush work items.
#include "halo_gemm.cuh"
#include "train_kernels.cuh"

namespace rst {

using namespace umma;

constexpr int kWgTW = 32;                              // pixels per work item: one image-row segment
constexpr int kWgXBox = 5120;                          // (32 + 2) rows x 128 B, padded to whole 1024-byte swizzle atoms
constexpr int kWgGBox = kWgTW * 128;                   // 4096
constexpr int kWgStage = 2 * 4 * kWgXBox + 2 * 4 * kWgGBox;      // [X hi | X lo | G hi | G lo] x 4 channel groups = 72 KB
constexpr int kWgStages = 3;
constexpr int kWgThreads = 192;                        // TMA warp, MMA warp, 4 epilogue warps
constexpr int kWgFlush = 40;                           // work items between accumulator flushes

__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_tf32_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmXlo,
                  const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmGlo, float* __restrict__ dw,
                  int B, int H, int tiles_w, int nslices, int split) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* tail = smem + kWgStages * kWgStage;
    uint64_t* full = reinterpret_cast<uint64_t*>(tail);
    uint64_t* empty = full + kWgStages;
    uint64_t* acc_full = empty + kWgStages;            // MMA -> epilogue: the accumulators hold a finished chunk
    uint64_t* acc_empty = acc_full + 1;                // epilogue -> MMA: they have been read
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ky = blockIdx.x % 3, slice = blockIdx.x / 3;
    const long long items = (long long)B * H * tiles_w;
    const long long per = (items + nslices - 1) / nslices;
    const long long it0 = slice * per, it1 = it0 + per < items ? it0 + per : items;

    if (threadIdx.x == 0) {
        for (int i = 0; i < kWgStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(acc_full, 1);
        mbar_init(acc_empty, 4);
        fence_barrier_init();
    }
    if (warp == 0 && lane == 0) { prefetch_tmap(&tmX); prefetch_tmap(&tmG); prefetch_tmap(&tmXlo); prefetch_tmap(&tmGlo); }
    __syncthreads();
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            const uint32_t tx = (uint32_t)((split ? 2 : 1) * 4 * ((kWgTW + 2) * 128 + kWgGBox));
            uint32_t stage = 0, phase = 0;
            for (long long it = it0; it < it1; ++it) {
                const int xt = (int)(it % tiles_w);
                const long long r = it / tiles_w;
                const int y = (int)(r % H), n = (int)(r / H);
                mbar_wait(&empty[stage], phase ^ 1);
                mbar_expect_tx(&full[stage], tx);
                uint8_t* st = smem + stage * kWgStage;
#pragma unroll
                for (int cg = 0; cg < 4; ++cg) {
                    // maps are declared over bf16 pairs: 64 elements = 32 fp32 channels
                    tma_load_4d(st + cg * kWgXBox, &tmX, &full[stage], cg * 64, y + ky - 1, xt * kWgTW - 1, n);
                    tma_load_4d(st + 8 * kWgXBox + cg * kWgGBox, &tmG, &full[stage], cg * 64, y, xt * kWgTW, n);
                    if (split) {
                        tma_load_4d(st + (4 + cg) * kWgXBox, &tmXlo, &full[stage], cg * 64, y + ky - 1, xt * kWgTW - 1, n);
                        tma_load_4d(st + 8 * kWgXBox + (4 + cg) * kWgGBox, &tmGlo, &full[stage], cg * 64, y, xt * kWgTW, n);
                    }
                }
                if (++stage == kWgStages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // both operands MN-major (SWIZZLE_128B_BASE32B): leading-dimension offset = one channel-group box, 4-pixel groups 512 B apart
        const uint32_t idesc = make_idesc_tf32(128, 128) | (1u << 15) | (1u << 16);
        const uint64_t dg_const = make_smem_desc(0, kWgGBox, 512, SWIZZLE_128B_BASE32B);
        const uint64_t dx_const = make_smem_desc(0, kWgXBox, 512, SWIZZLE_128B_BASE32B);
        const bool issuer = elect_one();
        const uint32_t s16 = __shfl_sync(0xffffffffu, smem_u32(smem) >> 4, 0);
        uint32_t stage = 0, phase = 0, chunk_phase = 0;
        long long in_chunk = 0;
        for (long long it = it0; it < it1; ++it) {
            if (in_chunk == 0 && it != it0) {              // the previous chunk must have left the accumulators
                mbar_wait(acc_empty, chunk_phase);
                chunk_phase ^= 1;
                tc_fence_after();
            }
            mbar_wait(&full[stage], phase);
            tc_fence_after();
            const uint32_t st16 = s16 + stage * (kWgStage >> 4);
            const uint32_t x_hi = st16, x_lo = st16 + (4 * kWgXBox >> 4);
            const uint32_t g_hi = st16 + (8 * kWgXBox >> 4), g_lo = g_hi + (4 * kWgGBox >> 4);
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
#pragma unroll
                for (int kg = 0; kg < kWgTW / 8; ++kg) {
                    const uint32_t xo = (uint32_t)((kg * 8 + kx) * 128) >> 4, go = (uint32_t)(kg * 1024) >> 4;
                    const uint32_t acc = (in_chunk != 0 || kg != 0) ? 1u : 0u;
                    if (issuer) {
                        mma_tf32_ss(tmem_base + kx * 128, dg_const | (uint64_t)(g_hi + go), dx_const | (uint64_t)(x_hi + xo), idesc, acc);
                        if (split) {
                            mma_tf32_ss(tmem_base + kx * 128, dg_const | (uint64_t)(g_hi + go), dx_const | (uint64_t)(x_lo + xo), idesc, 1u);
                            mma_tf32_ss(tmem_base + kx * 128, dg_const | (uint64_t)(g_lo + go), dx_const | (uint64_t)(x_hi + xo), idesc, 1u);
                        }
                    }
                }
            }
            if (issuer) mma_commit(&empty[stage]);
            if (++stage == kWgStages) { stage = 0; phase ^= 1; }
            if (++in_chunk == kWgFlush || it + 1 == it1) {
                if (issuer) mma_commit(acc_full);
                in_chunk = 0;
            }
        }
    } else {
        // epilogue: TMEM lane = co, column = ci; a warp's 32 lanes add to 32 consecutive floats of dW[ky, kx, ci, :]
        const int q = warp & 3;
        const long long n_items = it1 - it0;
        const long long chunks = n_items > 0 ? (n_items + kWgFlush - 1) / kWgFlush : 0;
        uint32_t ph = 0;
        for (long long c = 0; c < chunks; ++c) {
            mbar_wait(acc_full, ph);
            ph ^= 1;
            tc_fence_after();
#pragma unroll 1
            for (int kx = 0; kx < 3; ++kx) {
#pragma unroll 1
                for (int cc = 0; cc < 4; ++cc) {
                    float v[32];
                    tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + kx * 128 + cc * 32, v);
                    tmem_ld_wait();
                    float* o = dw + ((size_t)((ky * 3 + kx) * 128 + cc * 32)) * 128 + q * 32 + lane;
#pragma unroll
                    for (int j = 0; j < 32; ++j) atomicAdd(o + (size_t)j * 128, v[j]);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

size_t wgrad_tf32_scratch_floats(int B, int H, int W) { return 2 * (size_t)B * H * W * 128; }

// x, g: (B, H, W, 128) fp32; dw (3, 3, 128, 128) is accumulated into; lo_scratch: wgrad_tf32_scratch_floats() floats (split only)
cudaError_t launch_wgrad_tf32(const float* x, const float* g, float* dw, float* lo_scratch, int B, int H, int W, bool split,
                              int num_sms, cudaStream_t s, std::string* err) {
    if ((long long)B * H * W == 0) return cudaSuccess;
    const size_t n = (size_t)B * H * W * 128;
    const float* xlo = x;
    const float* glo = g;
    if (split) {
        float* a = lo_scratch;
        float* b = lo_scratch + n;
        cudaError_t e = launch_tf32_lo(x, a, (long long)n, s);
        if (e == cudaSuccess) e = launch_tf32_lo(g, b, (long long)n, s);
        if (e != cudaSuccess) return e;
        xlo = a; glo = b;
    }
    CUtensorMap tmX, tmXlo, tmG, tmGlo;
    // fp32 channels declared as bf16 pairs (the encoder is shared with the bf16 path): 256 elements per pixel, 64 per box row
    if (!encode_halo_map(&tmX, x, B, H, W, 256, 64, 1, kWgTW + 2, err, true) || !encode_halo_map(&tmXlo, xlo, B, H, W, 256, 64, 1, kWgTW + 2, err, true) ||
        !encode_halo_map(&tmG, g, B, H, W, 256, 64, 1, kWgTW, err, true) || !encode_halo_map(&tmGlo, glo, B, H, W, 256, 64, 1, kWgTW, err, true))
        return cudaErrorInvalidValue;
    const size_t smem = (size_t)kWgStages * kWgStage + 1024 + 256;
    static SmemAttrCache configured;
    if (cudaError_t e = ensure_dynamic_smem(wgrad_tf32_kernel, smem, configured)) return e;
    const int tiles_w = ceil_div(W, kWgTW);
    const long long items = (long long)B * H * tiles_w;
    int nslices = num_sms / 3;
    if (nslices > items) nslices = (int)items;
    if (nslices < 1) nslices = 1;
    wgrad_tf32_kernel<<<3 * nslices, kWgThreads, smem, s>>>(tmX, tmXlo, tmG, tmGlo, dw, B, H, tiles_w, nslices, split ? 1 : 0);
    return cudaGetLastError();
}

}  // namespace rst
