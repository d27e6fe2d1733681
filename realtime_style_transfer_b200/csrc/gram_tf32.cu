// Gram matrices of the style loss (styleLoss.py:11-18) on the tensor cores.
//
//   G[b, c, d] = (1 / P) * sum_p F[b, p, c] * F[b, p, d]            F: (B, P = H*W, C) fp32, C in {64, 128, 256, 512}
//
// is a GEMM whose reduction dimension is the PIXEL index, while NHWC is contiguous along the channels: both operands are
// "MN-major", exactly the situation of the trunk's weight gradient (wgrad_tf32.cu), whose recipe is reused: TMA boxes of
// 32 channels x 32 pixels land as 32 rows of 128 B in the 128-byte swizzle with 32-byte atoms (the only MN-major layout
// tcgen05 kind::tf32 accepts), the instruction descriptor sets the a/b-major bits, a K-step is 8 pixel rows (1 KB further).
//
//   D[c, d] += F[pixels, c-block]^T  F[pixels, d-block]       A = 128 channels (M), B = min(128, C) channels (N), K = 8 pixels
//
// One CTA owns (sample, c-block, d-block, pixel slice).  Diagonal blocks load one operand and point both descriptors at it.
// TMEM lane = c, column = d; the epilogue writes G[b, d, c] (G is symmetric), so that a warp's 32 lanes add to 32 consecutive
// floats.  fp32-level accuracy (RST_PRECISION_FP32 math): the hardware reads only the upper 19 bits of an operand, so each
// product is issued as three MMAs, hi*hi + hi*lo + lo*hi, with lo = rna_tf32(v - trunc_tf32(v)) written by a small pre-pass;
// the accumulator (which truncates on every add) is flushed to the fp32 result with atomics every kGrFlush work items.
// The pass is HBM-bound (64 FLOP per byte of F at C = 64 .. 512 at C = 512 against a tensor/HBM ridge of ~100 for tf32 x3):
// the point of the tensor cores here is that the reduction keeps up with the stream, which the SIMT kernel did not.
#include "halo_gemm.cuh"
#include "train_kernels.cuh"

namespace rst {

using namespace umma;

constexpr int kGrTP = 32;                              // pixels per work item
constexpr int kGrBox = kGrTP * 128;                    // one 32-channel group of one item: 4 KB
constexpr int kGrOperand = 4 * kGrBox;                 // 128 channels: 16 KB
constexpr int kGrStage = 4 * kGrOperand;               // [A hi | A lo | B hi | B lo] = 64 KB
constexpr int kGrStages = 3;
constexpr int kGrThreads = 192;                        // TMA warp, MMA warp, 4 epilogue warps
constexpr int kGrFlush = 40;                           // work items between accumulator flushes (chains of 40 x 12 MMAs)

__global__ void __launch_bounds__(kGrThreads, 1)
gram_tf32_kernel(const __grid_constant__ CUtensorMap tmF, const __grid_constant__ CUtensorMap tmFlo, float* __restrict__ gram,
                 int C, int P, int nblk, int nslices, int split, float inv_p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* tail = smem + kGrStages * kGrStage;
    uint64_t* full = reinterpret_cast<uint64_t*>(tail);
    uint64_t* empty = full + kGrStages;
    uint64_t* acc_full = empty + kGrStages;
    uint64_t* acc_empty = acc_full + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int slice = blockIdx.x, bi = blockIdx.y / nblk, bj = blockIdx.y % nblk, b = blockIdx.z;
    const bool diag = bi == bj;
    const int N = C < 128 ? C : 128;                   // columns of D in use (C = 64: the upper half of A's rows is zero fill)
    const int items = (P + kGrTP - 1) / kGrTP;
    const int per = (items + nslices - 1) / nslices;
    const int it0 = slice * per, it1 = min(items, it0 + per);

    if (threadIdx.x == 0) {
        for (int i = 0; i < kGrStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(acc_full, 1);
        mbar_init(acc_empty, 4);
        fence_barrier_init();
    }
    if (warp == 0 && lane == 0) { prefetch_tmap(&tmF); prefetch_tmap(&tmFlo); }
    __syncthreads();
    if (warp == 1) tmem_alloc(tmem_slot, 128);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            const int b_groups = N / 32;
            const uint32_t tx = (uint32_t)((split ? 2 : 1) * (4 + (diag ? 0 : b_groups)) * kGrBox);
            uint32_t stage = 0, phase = 0;
            for (int it = it0; it < it1; ++it) {
                mbar_wait(&empty[stage], phase ^ 1);
                mbar_expect_tx(&full[stage], tx);
                uint8_t* st = smem + stage * kGrStage;
                const int p0 = it * kGrTP;
                // maps are declared over bf16 pairs: 64 elements = 32 fp32 channels; channel groups past C are zero filled
#pragma unroll
                for (int cg = 0; cg < 4; ++cg) {
                    tma_load_4d(st + cg * kGrBox, &tmF, &full[stage], (bi * 4 + cg) * 64, 0, p0, b);
                    if (split) tma_load_4d(st + kGrOperand + cg * kGrBox, &tmFlo, &full[stage], (bi * 4 + cg) * 64, 0, p0, b);
                }
                if (!diag) {
                    for (int cg = 0; cg < b_groups; ++cg) {
                        tma_load_4d(st + 2 * kGrOperand + cg * kGrBox, &tmF, &full[stage], (bj * 4 + cg) * 64, 0, p0, b);
                        if (split) tma_load_4d(st + 3 * kGrOperand + cg * kGrBox, &tmFlo, &full[stage], (bj * 4 + cg) * 64, 0, p0, b);
                    }
                }
                if (++stage == kGrStages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // both operands MN-major (SWIZZLE_128B_BASE32B): leading-dimension offset = one channel-group box, 4-pixel groups 512 B apart
        const uint32_t idesc = make_idesc_tf32(128, N) | (1u << 15) | (1u << 16);
        const uint64_t d_const = make_smem_desc(0, kGrBox, 512, SWIZZLE_128B_BASE32B);
        const bool issuer = elect_one();
        const uint32_t s16 = __shfl_sync(0xffffffffu, smem_u32(smem) >> 4, 0);
        uint32_t stage = 0, phase = 0, chunk_phase = 0;
        int in_chunk = 0;
        for (int it = it0; it < it1; ++it) {
            if (in_chunk == 0 && it != it0) {              // the previous chunk must have left the accumulator
                mbar_wait(acc_empty, chunk_phase);
                chunk_phase ^= 1;
                tc_fence_after();
            }
            mbar_wait(&full[stage], phase);
            tc_fence_after();
            const uint32_t st16 = s16 + stage * (kGrStage >> 4);
            const uint32_t a_hi = st16, a_lo = st16 + (kGrOperand >> 4);
            const uint32_t b_hi = diag ? a_hi : st16 + (2 * kGrOperand >> 4), b_lo = diag ? a_lo : st16 + (3 * kGrOperand >> 4);
#pragma unroll
            for (int kg = 0; kg < kGrTP / 8; ++kg) {
                const uint32_t o = (uint32_t)(kg * 1024) >> 4;
                const uint32_t acc = (in_chunk != 0 || kg != 0) ? 1u : 0u;
                if (issuer) {
                    mma_tf32_ss(tmem_base, d_const | (uint64_t)(a_hi + o), d_const | (uint64_t)(b_hi + o), idesc, acc);
                    if (split) {
                        mma_tf32_ss(tmem_base, d_const | (uint64_t)(a_hi + o), d_const | (uint64_t)(b_lo + o), idesc, 1u);
                        mma_tf32_ss(tmem_base, d_const | (uint64_t)(a_lo + o), d_const | (uint64_t)(b_hi + o), idesc, 1u);
                    }
                }
            }
            if (issuer) mma_commit(&empty[stage]);
            if (++stage == kGrStages) { stage = 0; phase ^= 1; }
            if (++in_chunk == kGrFlush || it + 1 == it1) {
                if (issuer) mma_commit(acc_full);
                in_chunk = 0;
            }
        }
    } else {
        // epilogue: TMEM lane = c, column = d; written transposed (G is symmetric): 32 lanes -> 32 consecutive floats of G[b, d, :]
        const int q = warp & 3;
        const int n_items = it1 - it0;
        const int chunks = n_items > 0 ? (n_items + kGrFlush - 1) / kGrFlush : 0;
        const int c = bi * 128 + q * 32 + lane;
        uint32_t ph = 0;
        for (int ch = 0; ch < chunks; ++ch) {
            mbar_wait(acc_full, ph);
            ph ^= 1;
            tc_fence_after();
#pragma unroll 1
            for (int cc = 0; cc < N / 32; ++cc) {
                float v[32];
                tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + cc * 32, v);
                tmem_ld_wait();
                if (c < C) {
                    float* o = gram + ((size_t)b * C + (size_t)bj * 128 + cc * 32) * C + c;
#pragma unroll
                    for (int j = 0; j < 32; ++j) atomicAdd(o + (size_t)j * C, v[j] * inv_p);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 128);
}

bool gram_tf32_supported(int C) { return C == 64 || (C % 128 == 0 && C <= 1024); }
size_t gram_tf32_scratch_floats(int B, int P, int C) { return (size_t)B * P * C; }

// x: (B, P, C) fp32; gram: (B, C, C), overwritten.  split: fp32-level accuracy (needs lo_scratch of gram_tf32_scratch_floats()).
cudaError_t launch_gram_tf32(const float* x, float* gram, float* lo_scratch, int B, int P, int C, bool split, int num_sms,
                             cudaStream_t s, std::string* err) {
    if ((long long)B * P == 0) return cudaSuccess;
    if (!gram_tf32_supported(C)) return cudaErrorInvalidValue;
    cudaError_t e = cudaMemsetAsync(gram, 0, (size_t)B * C * C * sizeof(float), s);
    if (e != cudaSuccess) return e;
    const float* lo = x;
    if (split) {
        if (!lo_scratch) return cudaErrorInvalidValue;
        e = launch_tf32_lo(x, lo_scratch, (long long)B * P * C, s);
        if (e != cudaSuccess) return e;
        lo = lo_scratch;
    }
    CUtensorMap tmF, tmFlo;
    // (B, P, C) fp32 as the 4-D activation map (c, h = 1, w = P, n = B) over bf16 pairs; box = 32 channels x kGrTP pixels
    if (!encode_halo_map(&tmF, x, B, 1, P, 2 * C, 64, 1, kGrTP, err, true) ||
        !encode_halo_map(&tmFlo, lo, B, 1, P, 2 * C, 64, 1, kGrTP, err, true))
        return cudaErrorInvalidValue;
    const size_t smem = (size_t)kGrStages * kGrStage + 1024 + 256;
    static SmemAttrCache configured;
    if ((e = ensure_dynamic_smem(gram_tf32_kernel, smem, configured)) != cudaSuccess) return e;
    const int nblk = (C + 127) / 128;
    const int items = (P + kGrTP - 1) / kGrTP;
    int nslices = (3 * num_sms + B * nblk * nblk - 1) / (B * nblk * nblk);      // ~3 CTAs per SM in flight over the launch
    if (nslices > items) nslices = items;
    if (nslices < 1) nslices = 1;
    dim3 grid((unsigned)nslices, (unsigned)(nblk * nblk), (unsigned)B);
    gram_tf32_kernel<<<grid, kGrThreads, smem, s>>>(tmF, tmFlo, gram, C, P, nblk, nslices, split ? 1 : 0, 1.f / (float)P);
    return cudaGetLastError();
}

}  // namespace rst
