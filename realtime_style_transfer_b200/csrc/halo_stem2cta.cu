// 2-CTA (tcgen05 cta_group::2) variant of the pair-pixel 9x9 stem (SCH_STEM2, 17 channels -> 32, two pixels per GEMM row).
//
// Why: the 1-CTA stem is bound by shared-memory operand fetch -- every N = 64 MMA reads 4 KB of A and 2 KB of B for 32 cycles
// of tensor work (ncu: L1/shared 82.6 %, tensor pipe 57.8 %).  With UMMA_M = 256 across a CTA pair each SM still reads its own
// 4 KB of A but only HALF of B: its 32 of the 64 output columns = ONE weight unit.  The block-Toeplitz trick carries over
// unchanged: a K-step reads units U (columns of the even pixel) and U+1 (odd pixel); the leader CTA holds the unit array as is,
// the peer CTA holds it shifted by one unit, and both use the same descriptor.
#include "halo_gemm.cuh"

namespace rst {

using namespace umma;

constexpr int kSN = 64;                                             // output columns: 2 pixels x 32 channels
constexpr int kSHalo = sched_halo_h(SCH_STEM2) * sched_halo_w(SCH_STEM2) * 128;   // 16 x 20 row units x 128 B
constexpr int kSAStage = (kSHalo + 1023) & ~1023;
constexpr int kSAStages = 2;
constexpr int kSTail = 6400;
constexpr int kSThreads = 96 + 128 * 2;                             // 3 control warps + 8 epilogue warps

// SCH = SCH_STEM2 (17 channels, 100 K-steps, 15 weight boxes) or SCH_STEM2B (18 channels, 108 K-steps, 16 weight boxes)
template <int SCH>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kSThreads, 1)
halo_stem2cta_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const HaloGemmParams p) {
    constexpr int N = kSN, CW = 32, NCH = N / CW;
    constexpr uint32_t TMEM_COLS = 2 * N;
    constexpr int kSKS = sched_ksteps(SCH, 128), kBoxes = sched_b_boxes(SCH), kSBBytes = kBoxes * 8192;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;
    uint8_t* sB = sA + kSAStages * kSAStage;
    uint8_t* tail = sB + kSBBytes;
    uint64_t* a_full = reinterpret_cast<uint64_t*>(tail);
    uint64_t* a_empty = a_full + 4;
    uint64_t* b_full = a_empty + 4;
    uint64_t* acc_full = b_full + 2;
    uint64_t* acc_empty = acc_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
    float* bias_s = reinterpret_cast<float*>(tail + 256);
    float* scale_s = bias_s + N;
    float* shift_s = scale_s + N;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader_cta = rank == 0;
    const int tiles_per_img = p.tiles_h * p.tiles_w;
    const int total_tiles = p.B * tiles_per_img;
    const int total_pairs = (total_tiles + 1) / 2;
    const int nclusters = gridDim.x / 2, cid = blockIdx.x / 2;
    const int ppc = (total_pairs + nclusters - 1) / nclusters;
    const int pair_begin = cid * ppc;
    const int pair_end = min(total_pairs, pair_begin + ppc);

    if (threadIdx.x == 0) {
        for (int i = 0; i < 4; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
        mbar_init(&b_full[0], 1);
        for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 2 * 4 * NCH); }
        fence_barrier_init();
    }
    if (warp == 0 && lane == 0) { prefetch_tmap(&tmA); prefetch_tmap(&tmB); }
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        bias_s[i] = p.bias ? p.bias[i] : 0.f;
        scale_s[i] = p.post_scale ? p.post_scale[i] : 1.f;
        shift_s[i] = p.post_shift ? p.post_shift[i] : 0.f;
    }
    __syncthreads();
    cluster_sync();                                           // barriers of both CTAs are initialised
    if (warp == 2) tmem_alloc_2sm(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (warp != 1) { pdl_wait(); if (p.pdl_trigger) pdl_trigger(); }     // the weight producer reads constants only

    auto tile_coords = [&](int t, int& n, int& h0, int& w0) {
        if (t >= total_tiles) { n = p.B; h0 = 0; w0 = 0; return; }     // phantom tile: out of bounds everywhere -> zeros
        n = t / tiles_per_img;
        const int r = t - n * tiles_per_img;
        h0 = (r / p.tiles_w) * 8; w0 = (r % p.tiles_w) * 16;
    };

    if (warp == 0) {
        // ================= A producer (both CTAs): own halo patch, transactions land on the leader's barrier ====
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (int pr = pair_begin; pr < pair_end; ++pr) {
                int n, h0, w0;
                tile_coords(2 * pr + (int)rank, n, h0, w0);
                mbar_wait(&a_empty[stage], phase ^ 1);
                if (leader_cta) mbar_expect_tx(&a_full[stage], 2 * kSHalo);
                if (p.l2_hints & 1) tma_load_4d_2sm_hint(sA + stage * kSAStage, &tmA, &a_full[stage], 0, h0 + p.oy, w0 + p.ox, n, l2_policy_evict_first());
                else tma_load_4d_2sm(sA + stage * kSAStage, &tmA, &a_full[stage], 0, h0 + p.oy, w0 + p.ox, n);
                if (++stage == kSAStages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ================= B producer (both CTAs): the whole unit array once; the peer CTA's copy starts one unit later ===
        if (lane == 0 && pair_begin < pair_end) {
            if (leader_cta) mbar_expect_tx(&b_full[0], 2 * kSBBytes);
            for (int kb = 0; kb < kBoxes; ++kb)
                tma_load_2d_2sm(sB + kb * 8192, &tmB, &b_full[0], 0, kb * 256 + (int)rank * 32);
        }
    } else if (warp == 2) {
        // ================= MMA issuer: leader CTA only ==========================================================
        if (leader_cta) {
            const uint32_t idesc = make_idesc_bf16(256, N);
            const uint64_t da_const = make_smem_desc(0, 16, sched_halo_h(SCH) * 128, SWIZZLE_128B);
            const uint64_t db_const = make_smem_desc(0, 16, 256, SWIZZLE_32B);
            const bool issuer = elect_one();
            uint32_t as = 0, aph = 0, cs = 0, cph = 0;
            if (pair_begin < pair_end) mbar_wait(&b_full[0], 0);
            const uint32_t sB16 = __shfl_sync(0xffffffffu, smem_u32(sB) >> 4, 0);
            const uint32_t sA16 = __shfl_sync(0xffffffffu, smem_u32(sA) >> 4, 0);
            const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
            for (int pr = pair_begin; pr < pair_end; ++pr) {
                mbar_wait(&acc_empty[cs], cph ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_u + cs * N;
                mbar_wait(&a_full[as], aph);
                tc_fence_after();
                const uint32_t a_base16 = sA16 + as * (kSAStage >> 4);
#pragma unroll
                for (int ks = 0; ks < kSKS; ++ks) {
                    const uint64_t da = da_const | (uint64_t)(a_base16 + (uint32_t)(sched_off(SCH, 128, ks) >> 4));
                    const uint64_t db = db_const | (uint64_t)(sB16 + sched_b_off16(SCH, ks));
                    if (issuer) mma_f16_ss_2sm(tmem_d, da, db, idesc, ks != 0 ? 1u : 0u);
                }
                if (issuer) mma_commit_2sm(&a_empty[as], 3);
                if (++as == kSAStages) { as = 0; aph ^= 1; }
                if (issuer) mma_commit_2sm(&acc_full[cs], 3);
                if (++cs == 2) { cs = 0; cph ^= 1; }
            }
        }
    } else {
        // ================= epilogue (both CTAs): ReLU(bias) -> BatchNorm affine -> ReLU -> bf16, one warp per (quadrant, pixel) ==
        const int q = warp & 3, c = (warp - 3) >> 2;
        const int row = q * 32 + lane;
        const int w_l = row >> 3, h_l = row & 7;
        uint32_t cs = 0, cph = 0;
        for (int pr = pair_begin; pr < pair_end; ++pr) {
            int n, h0, w0;
            tile_coords(2 * pr + (int)rank, n, h0, w0);
            const int gh = h0 + h_l, gw = w0 + w_l;
            const bool valid = n < p.B && gh < p.H && gw < p.WRU;
            mbar_wait(&acc_full[cs], cph);
            tc_fence_after();
            float v[32];
            tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + cs * N + c * CW, v);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(&acc_empty[cs]);
            if (valid) {
                uint32_t packed[CW / 2];
#pragma unroll
                for (int j = 0; j < CW; j += 2) {
                    float x0 = fmaxf(v[j] + bias_s[c * CW + j], 0.f), x1 = fmaxf(v[j + 1] + bias_s[c * CW + j + 1], 0.f);
                    x0 = fmaxf(fmaf(x0, scale_s[c * CW + j], shift_s[c * CW + j]), 0.f);
                    x1 = fmaxf(fmaf(x1, scale_s[c * CW + j + 1], shift_s[c * CW + j + 1]), 0.f);
                    __nv_bfloat162 h2 = __floats2bfloat162_rn(x0, x1);
                    packed[j >> 1] = *reinterpret_cast<uint32_t*>(&h2);
                }
                __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.y) +
                                   (((size_t)n * p.out_H + gh) * p.out_W + gw) * p.out_C + c * CW;
                st_global_v8(o, packed);
                st_global_v8(o + 16, packed + 8);
            }
            if (++cs == 2) { cs = 0; cph ^= 1; }
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync();                                           // both CTAs are done with TMEM and with remote barriers
    if (warp == 2) tmem_dealloc_2sm(tmem_base, TMEM_COLS);
}

size_t halo_stem2cta_smem_bytes(int sch) { return (size_t)kSAStages * kSAStage + sched_b_boxes(sch) * 8192 + kSTail + 1024; }

template <int SCH>
static cudaError_t launch_stem2cta(const CUtensorMap& tmA, const CUtensorMap& tmB_units, const HaloGemmParams& p, int num_sms,
                                   cudaStream_t s) {
    static SmemAttrCache configured;
    const size_t smem = halo_stem2cta_smem_bytes(SCH);
    if (cudaError_t e = ensure_dynamic_smem(halo_stem2cta_kernel<SCH>, smem, configured)) return e;
    const int total = p.B * p.tiles_h * p.tiles_w;
    if (total == 0) return cudaSuccess;
    const int pairs = (total + 1) / 2;
    int clusters = num_sms / 2;
    if (pairs < clusters) clusters = pairs;
    if (cudaError_t e = launch_pdl(p.pdl != 0, halo_stem2cta_kernel<SCH>, dim3(2 * clusters), dim3(kSThreads), smem, s, tmA, tmB_units, p)) return e;
    return cudaGetLastError();
}

cudaError_t launch_halo_stem2cta(int sch, const CUtensorMap& tmA, const CUtensorMap& tmB_units, const HaloGemmParams& p,
                                 int num_sms, cudaStream_t s) {
    if (sch == SCH_STEM2B) return launch_stem2cta<SCH_STEM2B>(tmA, tmB_units, p, num_sms, s);
    if (sch == SCH_STEM2) return launch_stem2cta<SCH_STEM2>(tmA, tmB_units, p, num_sms, s);
    return cudaErrorInvalidValue;
}

}  // namespace rst
