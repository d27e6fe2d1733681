// tcgen05 / TMA halo-GEMM kernel (design notes in halo_gemm.cuh) and its bf16 elementwise companions.
#include <mutex>

#include "halo_gemm.cuh"

namespace rst {

using namespace umma;

__device__ __forceinline__ float actf(float v, int act) {
    if (act == ACT_RELU) return fmaxf(v, 0.f);
    if (act == ACT_SIGMOID) return 1.f / (1.f + __expf(-v));
    return v;
}

// Every lane holds v[0..31] (32 columns of its row); afterwards lane j holds the sum over the 32 lanes of
// column j.  31 shuffles instead of 160.
__device__ __forceinline__ float warp_transpose_reduce(float (&v)[32], int lane) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const bool upper = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < off; ++i) {
            float send = upper ? v[i] : v[i + off];
            float keep = upper ? v[i + off] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    return v[0];
}

__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}

constexpr int kTailBytes = 6400;    // barriers (256) + per-column bias/scale/shift (3*N*4) + per-warp stats (4*2*N*4)

// MODE bits: what the epilogue does besides "+ bias" (compile-time so that the epilogue stays small: an epilogue
// that does not fit the instruction cache is fetch-bound, see profiles/r01_01_icache.md)
constexpr int MODE_RELU = 1;      // ReLU after the bias
constexpr int MODE_POST = 2;      // then per-column affine (inference BatchNorm) and a second ReLU
constexpr int MODE_F32 = 4;       // store fp32 instead of bf16
constexpr int MODE_TF32 = 8;      // operands are fp32 (tf32) instead of bf16: K = 8 per MMA, same byte geometry; stores are
                                  // rounded to tf32 so that the next layer's operand fetch (which truncates) sees exact values

// epilogue warps per TMEM lane quadrant: one per 32-column chunk (the epilogue is issue-latency bound, more warps hide it)
__host__ __device__ constexpr int halo_esplit(int n) { return n >= 128 ? 4 : n >= 64 ? 2 : 1; }
__host__ __device__ constexpr int halo_threads(int n) { return 96 + 128 * halo_esplit(n); }

template <int N, int ROWB, int EPI, int MODE, int SCH, bool BRES>
__global__ void __launch_bounds__(halo_threads(N), 1)
halo_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const HaloGemmParams p) {
    constexpr int BBLK = N * 128;                      // bytes of one B block (4 K-steps)
    constexpr uint64_t SWZ = ROWB == 128 ? SWIZZLE_128B : SWIZZLE_64B;
    constexpr int ROW_ELEMS = ROWB / 2;
    constexpr uint32_t TMEM_COLS = 2 * N < 32 ? 32 : 2 * N;
    constexpr int CW = N >= 32 ? 32 : 16;              // columns per epilogue chunk
    constexpr int NCH = N / CW;
    constexpr int ESPLIT = halo_esplit(N);             // epilogue warps per TMEM lane quadrant
    constexpr int KS = sched_ksteps(SCH, ROWB);         // K-steps per channel group
    static_assert(N % 16 == 0 && N >= 16 && N <= 256, "UMMA_M=128 needs 16 <= N <= 256, N % 16 == 0");
    static_assert((TMEM_COLS & (TMEM_COLS - 1)) == 0, "TMEM allocation must be a power of two");

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int halo_bytes = p.halo_h * p.halo_w * ROWB;
    const int a_stage_bytes = (halo_bytes + 1023) & ~1023;
    const int blocks_per_tile = p.n_groups * p.ksteps / 4;
    const int nbslots = BRES ? blocks_per_tile : p.n_bstages;
    uint8_t* sA = smem;
    uint8_t* sB = sA + p.n_astages * a_stage_bytes;
    uint8_t* tail = sB + (sched_b_units(SCH) ? sched_b_boxes(SCH) * 8192 : nbslots * BBLK);
    uint64_t* a_full = reinterpret_cast<uint64_t*>(tail);
    uint64_t* a_empty = a_full + 4;
    uint64_t* b_full = a_empty + 4;
    uint64_t* b_empty = b_full + 8;
    uint64_t* acc_full = b_empty + 8;
    uint64_t* acc_empty = acc_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
    float* bias_s = reinterpret_cast<float*>(tail + 256);
    float* scale_s = bias_s + N;
    float* shift_s = scale_s + N;
    float* stat_s = shift_s + N;                       // [4 warps][2][N]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_per_img = p.tiles_h * p.tiles_w;
    const int total_tiles = p.B * tiles_per_img;
    const int tpc = (total_tiles + gridDim.x - 1) / gridDim.x;
    const int tile_begin = blockIdx.x * tpc;
    const int tile_end = min(total_tiles, tile_begin + tpc);

    if (threadIdx.x == 0) {
        for (int i = 0; i < 4; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < 8; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4 * ESPLIT); }
        fence_barrier_init();
    }
    if (warp == 0 && lane == 0) { prefetch_tmap(&tmA); prefetch_tmap(&tmB); }
    if (warp == 2) tmem_alloc(tmem_slot, TMEM_COLS);
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        bias_s[i] = p.bias ? p.bias[i] : 0.f;
        scale_s[i] = p.post_scale ? p.post_scale[i] : 1.f;
        shift_s[i] = p.post_shift ? p.post_shift[i] : 0.f;
    }
    for (int i = threadIdx.x; i < 8 * N; i += blockDim.x) stat_s[i] = 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // everything but the weights (warp 1 streams constants only) may be something the previous kernel wrote or still reads
    if (warp != 1) { pdl_wait(); if (p.pdl_trigger) pdl_trigger(); }

    if (warp == 0) {
        // ================= A producer: one halo patch per (tile, channel group) =================
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            const uint64_t pol_first = l2_policy_evict_first();
            for (int t = tile_begin; t < tile_end; ++t) {
                const int n = t / tiles_per_img, r = t - n * tiles_per_img;
                const int h0 = (r / p.tiles_w) * 8, w0 = (r % p.tiles_w) * 16;
                for (int g = 0; g < p.n_groups; ++g) {
                    mbar_wait(&a_empty[stage], phase ^ 1);
                    mbar_expect_tx(&a_full[stage], halo_bytes);
                    if (SCH == SCH_S2D) {
                        if (p.l2_hints & 1) tma_load_5d_hint(sA + stage * a_stage_bytes, &tmA, &a_full[stage], 0, h0, 0, w0, n, pol_first);
                        else tma_load_5d(sA + stage * a_stage_bytes, &tmA, &a_full[stage], 0, h0, 0, w0, n);
                    }
                    else if ((MODE & MODE_TF32) && p.a_wrap) {
                        // split tf32: one conv over [x_hi | x_lo | x_hi]; the tensor stores [x_hi | x_lo], the third part re-reads the first
                        tma_load_4d(sA + stage * a_stage_bytes, &tmA, &a_full[stage], (g >= p.a_wrap ? g - p.a_wrap : g) * ROW_ELEMS,
                                    h0 + p.oy, w0 * p.a_w_mul + p.ox, n);
                    } else if (p.l2_hints & 1) {
                        tma_load_4d_hint(sA + stage * a_stage_bytes, &tmA, &a_full[stage], g * ROW_ELEMS, h0 + p.oy, w0 * p.a_w_mul + p.ox, n, pol_first);
                    } else tma_load_4d(sA + stage * a_stage_bytes, &tmA, &a_full[stage], g * ROW_ELEMS, h0 + p.oy, w0 * p.a_w_mul + p.ox, n);
                    if (++stage == (uint32_t)p.n_astages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================= B producer ===========================================================
        if (lane == 0 && tile_begin < tile_end) {
            if (sched_b_units(SCH)) {
                mbar_expect_tx(&b_full[0], sched_b_boxes(SCH) * 8192);
                for (int kb = 0; kb < sched_b_boxes(SCH); ++kb) tma_load_2d(sB + kb * 8192, &tmB, &b_full[0], 0, kb * 256);
            } else if (BRES) {
                mbar_expect_tx(&b_full[0], blocks_per_tile * BBLK);
                for (int kb = 0; kb < blocks_per_tile; ++kb) tma_load_2d(sB + kb * BBLK, &tmB, &b_full[0], 0, kb * N);
            } else {
                uint32_t stage = 0, phase = 0;
                for (int t = tile_begin; t < tile_end; ++t) {
                    for (int kb = 0; kb < blocks_per_tile; ++kb) {
                        mbar_wait(&b_empty[stage], phase ^ 1);
                        mbar_expect_tx(&b_full[stage], BBLK);
                        tma_load_2d(sB + stage * BBLK, &tmB, &b_full[stage], 0, kb * N);
                        if (++stage == (uint32_t)p.n_bstages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 2) {
        // ================= MMA issuer ===========================================================
        // The whole warp runs the (fully unrolled, compile-time scheduled) loop so that descriptor arithmetic stays in
        // the uniform datapath; only the tcgen05.mma / tcgen05.commit instructions themselves are issued by lane 0.
        {
            const uint32_t idesc = (MODE & MODE_TF32) ? make_idesc_tf32(128, N) : make_idesc_bf16(128, N);
            const uint64_t da_const = make_smem_desc(0, 16, sched_a_unit_stride(SCH) * sched_halo_h(SCH) * ROWB, SWZ);
            const uint64_t db_const = sched_b_units(SCH) ? make_smem_desc(0, 16, 256, SWIZZLE_32B) : make_smem_desc(0, 16, 1024, SWIZZLE_128B);
            const bool leader = elect_one();
            uint32_t as = 0, aph = 0, bs = 0, bph = 0, cs = 0, cph = 0;
            if (BRES && tile_begin < tile_end) mbar_wait(&b_full[0], 0);
            // warp-uniform by construction; the shuffles let ptxas prove it and keep descriptors in uniform registers
            const uint32_t sB16 = __shfl_sync(0xffffffffu, smem_u32(sB) >> 4, 0);
            const uint32_t sA16 = __shfl_sync(0xffffffffu, smem_u32(sA) >> 4, 0);
            const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
            const uint32_t a_stage16 = (uint32_t)a_stage_bytes >> 4;
            // CHUNK (tf32 kernels): every channel group accumulates into a FRESH accumulator stage and the epilogue warps add the
            // partial sums in registers.  The tensor core truncates (rounds toward zero) at every accumulate step: measured
            // ~2^-25 relative per MMA, i.e. 4e-5 after the 1728 MMAs of a split-tf32 512-channel layer -- a bias that
            // compounds over VGG16's 13 layers.  Chains of 36 MMAs plus round-to-nearest fp32 adds outside remove it.
            constexpr bool CHUNK = (MODE & MODE_TF32) != 0;
            for (int t = tile_begin; t < tile_end; ++t) {
                if (!CHUNK) {
                    mbar_wait(&acc_empty[cs], cph ^ 1);
                    tc_fence_after();
                }
                uint32_t tmem_d = tmem_u + cs * N;
                for (int g = 0; g < p.n_groups; ++g) {
                    if (CHUNK) {
                        mbar_wait(&acc_empty[cs], cph ^ 1);
                        tc_fence_after();
                        tmem_d = tmem_u + cs * N;
                    }
                    mbar_wait(&a_full[as], aph);
                    tc_fence_after();
                    const uint32_t a_base16 = sA16 + as * a_stage16;
                    uint32_t b_base16 = sB16 + g * (KS / 4) * (BBLK / 16);
#pragma unroll
                    for (int ks = 0; ks < KS; ++ks) {
                        if (!BRES && (ks & 3) == 0) {
                            mbar_wait(&b_full[bs], bph);
                            tc_fence_after();
                            b_base16 = sB16 + bs * (BBLK / 16);
                        }
                        const uint64_t da = da_const | (uint64_t)(a_base16 + (uint32_t)(sched_off(SCH, ROWB, ks) >> 4));
                        const uint64_t db = db_const | (uint64_t)(sched_b_units(SCH) ? sB16 + sched_b_off16(SCH, ks)
                                                                                    : b_base16 + (BRES ? (ks / 4) * (BBLK / 16) : 0) + (ks & 3) * 2);
                        if (leader) {
                            if (MODE & MODE_TF32) mma_tf32_ss(tmem_d, da, db, idesc, ks == 0 ? 0u : 1u);
                            else mma_f16_ss(tmem_d, da, db, idesc, ks == 0 ? (uint32_t)(g != 0) : 1u);
                        }
                        if (!BRES && (ks & 3) == 3) {
                            if (leader) mma_commit(&b_empty[bs]);
                            if (++bs == (uint32_t)p.n_bstages) { bs = 0; bph ^= 1; }
                        }
                    }
                    if (leader) mma_commit(&a_empty[as]);
                    if (++as == (uint32_t)p.n_astages) { as = 0; aph ^= 1; }
                    if (CHUNK) {
                        if (leader) mma_commit(&acc_full[cs]);
                        if (++cs == 2) { cs = 0; cph ^= 1; }
                    }
                }
                if (!CHUNK) {
                    if (leader) mma_commit(&acc_full[cs]);
                    if (++cs == 2) { cs = 0; cph ^= 1; }
                }
            }
        }
    } else if (warp < 3 + 4 * ESPLIT) {
        // ================= epilogue ==============================================================
        const int q = warp & 3;                  // TMEM lane quadrant this warp may access
        const int c_begin = ((warp - 3) >> 2) * (NCH / ESPLIT), c_end = c_begin + NCH / ESPLIT;   // this warp's column chunks
        const int row = q * 32 + lane;
        const int w_l = row >> 3, h_l = row & 7;
        float* my_sum = stat_s + q * 2 * N;      // per-warp column accumulators (lane j owns column c*CW+j)
        float* my_sq = my_sum + N;
        const bool do_stats = p.stats != nullptr;
        int cur_n = -1;
        // Narrow kernels (N <= 64: few epilogue warps, registers to spare) keep per-thread column sums across all the tiles of
        // a sample and pay for the 2 x 31-shuffle warp reduction once per sample instead of once per tile.
        constexpr bool ACCUM = N <= 64 && NCH / ESPLIT == 1 && !(MODE & MODE_POST);
        float acc_s[ACCUM ? 32 : 1], acc_q[ACCUM ? 32 : 1];
#pragma unroll
        for (int j = 0; j < (ACCUM ? 32 : 1); ++j) { acc_s[j] = 0.f; acc_q[j] = 0.f; }
        auto flush = [&]() {
            if (do_stats && cur_n >= 0) {
                if constexpr (ACCUM) {
                    const float s1 = warp_transpose_reduce(acc_s, lane);
                    const float s2 = warp_transpose_reduce(acc_q, lane);
#pragma unroll
                    for (int j = 0; j < 32; ++j) { acc_s[j] = 0.f; acc_q[j] = 0.f; }
                    const int col = c_begin * CW + lane;
                    int ch = col;
                    bool on = lane < CW;
                    if (EPI == EPI_CONVT2) ch = col % p.stats_c;
                    if (EPI == EPI_QUAD3) { ch = col % 3; on = on && col < 12; }
                    if (EPI == EPI_OCT3) { ch = col % 4; on = on && ch < 3; }
                    if (on) {
                        double* dst = p.stats + ((size_t)cur_n * p.stats_c + ch) * 2;
                        atomicAdd(dst, (double)s1);
                        atomicAdd(dst + 1, (double)s2);
                    }
                } else {
#pragma unroll 1
                for (int c = c_begin; c < c_end; ++c) {
                    const int col = c * CW + lane;
                    int ch = col;
                    bool on = lane < CW;
                    if (EPI == EPI_CONVT2) ch = col % p.stats_c;
                    if (EPI == EPI_QUAD3) { ch = col % 3; on = on && col < 12; }
                    if (lane < CW) {
                        if (on) {
                            double* dst = p.stats + ((size_t)cur_n * p.stats_c + ch) * 2;
                            atomicAdd(dst, (double)my_sum[col]);
                            atomicAdd(dst + 1, (double)my_sq[col]);
                        }
                        my_sum[col] = 0.f; my_sq[col] = 0.f;
                    }
                }
                }
            }
        };
        uint32_t cs = 0, cph = 0;
        // tile coordinates advance incrementally: the two integer divisions per tile were on this warp's per-tile latency chain
        int n = tile_begin / tiles_per_img, tile_h, tile_w;
        {
            const int r0 = tile_begin - n * tiles_per_img;
            tile_h = r0 / p.tiles_w; tile_w = r0 - tile_h * p.tiles_w;
        }
        auto next_tile = [&]() {
            if (++tile_w == p.tiles_w) {
                tile_w = 0;
                if (++tile_h == p.tiles_h) { tile_h = 0; ++n; }
            }
        };
        for (int t = tile_begin; t < tile_end; ++t) {
            const int gh = tile_h * 8 + h_l, gw = tile_w * 16 + w_l;
            const bool valid = gh < p.H && gw < p.WRU;
            if (n != cur_n) { flush(); cur_n = n; }
            mbar_wait(&acc_full[cs], cph);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + cs * N;
            // A warp owns one or two column chunks: all its TMEM loads are issued up front, the accumulator stage is handed
            // back to the MMA warp as soon as the values are in registers, then the math / stores / statistics run.
            auto process = [&](float (&v)[32], int c) {
                const float* bs_ = bias_s + c * CW;
#pragma unroll
                for (int j = 0; j < CW; ++j) {
                    float x = v[j] + bs_[j];
                    if (MODE & MODE_RELU) x = fmaxf(x, 0.f);
                    if (MODE & MODE_POST) x = fmaxf(fmaf(x, scale_s[c * CW + j], shift_s[c * CW + j]), 0.f);
                    v[j] = x;
                }
                if (MODE & MODE_F32) {
                    if ((MODE & MODE_TF32) && !p.store_exact) {
#pragma unroll
                        for (int j = 0; j < CW; ++j) v[j] = round_tf32(v[j]);
                    }
                    if (valid && EPI == EPI_OCT3) {
                        // 8 pixels x 3 channels = 24 contiguous floats; column j*4 + 3 is padding
                        float4* o4 = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.y) + (((size_t)n * p.out_H + gh) * p.out_W + 8 * gw) * 3);
                        float w24[24];
#pragma unroll
                        for (int j = 0; j < 8; ++j) { w24[3 * j] = v[4 * j]; w24[3 * j + 1] = v[4 * j + 1]; w24[3 * j + 2] = v[4 * j + 2]; }
#pragma unroll
                        for (int j = 0; j < 6; ++j) o4[j] = make_float4(w24[4 * j], w24[4 * j + 1], w24[4 * j + 2], w24[4 * j + 3]);
                    } else if (valid) {
                        float4* o4;
                        if (EPI == EPI_QUAD3)
                            o4 = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.y) + (((size_t)n * p.out_H + gh) * p.out_W + 4 * gw) * 3);
                        else
                            o4 = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.y) + (((size_t)n * p.out_H + gh) * p.out_W + gw) * p.out_C + c * CW);
                        constexpr int NV = EPI == EPI_QUAD3 ? 3 : CW / 4;
#pragma unroll
                        for (int j = 0; j < NV; ++j) o4[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                    } else {
#pragma unroll
                        for (int j = 0; j < CW; ++j) v[j] = 0.f;
                    }
                } else {
                    // bf16 stores; statistics are taken on the values as stored
                    uint32_t packed[CW / 2];
#pragma unroll
                    for (int j = 0; j < CW; j += 2) {
                        __nv_bfloat162 h2 = __floats2bfloat162_rn(v[j], v[j + 1]);
                        packed[j >> 1] = *reinterpret_cast<uint32_t*>(&h2);
                        v[j] = valid ? __low2float(h2) : 0.f;
                        v[j + 1] = valid ? __high2float(h2) : 0.f;
                    }
                    if (valid) {
                        __nv_bfloat16* yb = reinterpret_cast<__nv_bfloat16*>(p.y);
                        if (EPI == EPI_NHWC) {
                            __nv_bfloat16* o = yb + (((size_t)n * p.out_H + gh) * p.out_W + gw) * p.out_C + c * CW;
                            if (CW % 16 == 0) {
#pragma unroll
                                for (int j = 0; j < CW / 16; ++j) st_global_v8(o + 16 * j, packed + 8 * j);   // full sectors
                            } else {
                                uint4* o4 = reinterpret_cast<uint4*>(o);
#pragma unroll
                                for (int j = 0; j < CW / 8; ++j)
                                    o4[j] = make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
                            }
                        } else if (EPI == EPI_CONVT2) {
                            // columns = phase*CQ + ch with CQ = N/4; a 32-column chunk holds 32/CQ phases
                            constexpr int CQ = N / 4 >= 8 ? N / 4 : 8;
                            constexpr int PER_CHUNK = CW / CQ > 0 ? CW / CQ : 1;
#pragma unroll
                            for (int sph = 0; sph < PER_CHUNK; ++sph) {
                                const int phase = c * PER_CHUNK + sph;
                                const int oy = 2 * gh + (phase >> 1), ox = 2 * gw + (phase & 1);
                                __nv_bfloat16* o = yb + (((size_t)n * p.out_H + oy) * p.out_W + ox) * CQ;
                                if (CQ % 16 == 0) {
#pragma unroll
                                    for (int j = 0; j < CQ / 16; ++j) st_global_v8(o + 16 * j, packed + sph * (CQ / 2) + 8 * j);
                                } else {
                                    uint4* o4 = reinterpret_cast<uint4*>(o);
#pragma unroll
                                    for (int j = 0; j < CQ / 8; ++j)
                                        o4[j] = make_uint4(packed[sph * (CQ / 2) + 4 * j], packed[sph * (CQ / 2) + 4 * j + 1],
                                                           packed[sph * (CQ / 2) + 4 * j + 2], packed[sph * (CQ / 2) + 4 * j + 3]);
                                }
                            }
                        }
                    }
                }
                if constexpr (ACCUM) {
                    if (do_stats) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) { acc_s[j] += v[j]; acc_q[j] = fmaf(v[j], v[j], acc_q[j]); }
                    }
                } else if (do_stats) {
                    float sq[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) sq[j] = v[j] * v[j];
                    const float s1 = warp_transpose_reduce(v, lane);
                    const float s2 = warp_transpose_reduce(sq, lane);
                    if (lane < CW) {
                        my_sum[c * CW + lane] += s1;
                        my_sq[c * CW + lane] += s2;
                    }
                }
            };
            constexpr int CPW = NCH / ESPLIT;        // chunks per warp: 1 or 2
            static_assert(CPW == 1 || CPW == 2, "epilogue warp handles one or two chunks");
            float va[32], vb[32];
            if constexpr ((MODE & MODE_TF32) != 0) {
                // partial sums of all channel groups but the last: read, release the stage, add (fp32, round to nearest)
                float part[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) part[j] = 0.f;
                for (int g = 0; g + 1 < p.n_groups; ++g) {
                    tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + cs * N + c_begin * 32, va);
                    tmem_ld_wait();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&acc_empty[cs]);
#pragma unroll
                    for (int j = 0; j < 32; ++j) part[j] += va[j];
                    if (++cs == 2) { cs = 0; cph ^= 1; }
                    mbar_wait(&acc_full[cs], cph);
                    tc_fence_after();
                }
                tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + cs * N + c_begin * 32, va);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[cs]);
#pragma unroll
                for (int j = 0; j < 32; ++j) va[j] += part[j];
                process(va, c_begin);
                if (++cs == 2) { cs = 0; cph ^= 1; }
                next_tile();
                continue;
            }
            if (CW == 32) tmem_ld_32x32(taddr + c_begin * 32, va);
            else {
                tmem_ld_32x16(taddr + c_begin * 16, va);
#pragma unroll
                for (int j = 16; j < 32; ++j) va[j] = 0.f;
            }
            if (CPW == 2) tmem_ld_32x32(taddr + (c_begin + 1) * 32, vb);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[cs]);
            process(va, c_begin);
            if (CPW == 2) process(vb, c_begin + 1);
            if (++cs == 2) { cs = 0; cph ^= 1; }
            next_tile();
        }
        flush();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
constexpr int kSmemBudget = 227 * 1024;

static bool sched_resident(const HaloGemmLaunch& l) {
    return !(l.sched == SCH_C3 && l.row_bytes == 128 && l.N == 128) && !(l.mode & MODE_TF32);
}

bool halo_gemm_plan(HaloGemmLaunch* l, HaloGemmParams* p, std::string* err) {
    p->ksteps = sched_ksteps(l->sched, l->row_bytes);
    p->halo_h = sched_halo_h(l->sched); p->halo_w = sched_halo_w(l->sched);
    p->oy = sched_oy(l->sched); p->ox = sched_ox(l->sched);
    const int a_stage = (p->halo_h * p->halo_w * l->row_bytes + 1023) & ~1023;
    const int bblk = l->N * 128;
    const int blocks = p->n_groups * p->ksteps / 4;
    const int fixed = 1024 + kTailBytes;
    p->b_resident = sched_resident(*l) ? 1 : 0;
    int b_bytes;
    if (sched_b_units(l->sched)) {
        b_bytes = sched_b_boxes(l->sched) * 8192;
    } else if (p->b_resident) {
        b_bytes = blocks * bblk;
    } else {
        p->n_bstages = 6;
        b_bytes = p->n_bstages * bblk;
    }
    p->n_astages = 4;
    while (p->n_astages > 1 && p->n_astages * a_stage + b_bytes + fixed > kSmemBudget) --p->n_astages;
    if (!p->b_resident && p->n_astages > 3) p->n_astages = 3;
    l->smem_bytes = (size_t)p->n_astages * a_stage + b_bytes + fixed;
    if (l->smem_bytes > (size_t)kSmemBudget || (p->b_resident && p->n_astages < 2)) {
        if (err) *err = "halo_gemm: tile does not fit in shared memory";
        return false;
    }
    return true;
}

template <int N, int ROWB, int EPI, int MODE, int SCH, bool BRES>
static cudaError_t launch_t(const HaloGemmLaunch& l, const CUtensorMap& tmA, const CUtensorMap& tmB, const HaloGemmParams& p,
                            int num_sms, cudaStream_t s) {
    static SmemAttrCache configured;
    if (cudaError_t e = ensure_dynamic_smem(halo_gemm_kernel<N, ROWB, EPI, MODE, SCH, BRES>, l.smem_bytes, configured)) return e;
    const int total = p.B * p.tiles_h * p.tiles_w;
    if (total == 0) return cudaSuccess;
    const int grid = total < num_sms ? total : num_sms;
    if (cudaError_t e = launch_pdl(p.pdl != 0, halo_gemm_kernel<N, ROWB, EPI, MODE, SCH, BRES>, dim3(grid), dim3(halo_threads(N)), l.smem_bytes, s, tmA, tmB, p)) return e;
    return cudaGetLastError();
}

cudaError_t launch_halo_gemm(const HaloGemmLaunch& l, const CUtensorMap& tmA, const CUtensorMap& tmB,
                             const HaloGemmParams& p, int num_sms, cudaStream_t s) {
    constexpr int STEM_MODE = MODE_RELU | MODE_POST;
#define RST_HALO_CASE(NN, RB, EP, MD, SC, RES) \
    if (l.N == NN && l.row_bytes == RB && l.epi == EP && l.mode == (MD) && l.sched == (SC)) \
        return launch_t<NN, RB, EP, (MD), (SC), RES>(l, tmA, tmB, p, num_sms, s);
    RST_HALO_CASE(128, 128, EPI_NHWC, MODE_RELU, SCH_C3, false)             // bottleneck 3x3 convs, weights streamed
    RST_HALO_CASE(128, 64, EPI_NHWC, MODE_RELU, SCH_C3, true)               // residual_block_0/conv0 (32 input channels)
    RST_HALO_CASE(64, 128, EPI_NHWC, MODE_RELU, SCH_C3, true)
    RST_HALO_CASE(64, 64, EPI_NHWC, MODE_RELU, SCH_C3, true)
    RST_HALO_CASE(64, 128, EPI_NHWC, STEM_MODE, SCH_STEM2, true)            // 9x9 stem, 17 channels, two pixels per GEMM row
    RST_HALO_CASE(64, 128, EPI_NHWC, STEM_MODE, SCH_STEM2B, true)           // 9x9 stem, 18 channels, two pixels per GEMM row
    RST_HALO_CASE(32, 64, EPI_NHWC, STEM_MODE, SCH_STEM + 4 + 1, true)      // 9x9 stem, 17 channels: 16 real + 1 windowed
    RST_HALO_CASE(32, 64, EPI_NHWC, STEM_MODE, SCH_STEM + 4 + 0, true)      // 5..16 channels
    RST_HALO_CASE(32, 128, EPI_NHWC, STEM_MODE, SCH_STEM + 4 + 2, true)     // 18 channels
    RST_HALO_CASE(32, 128, EPI_NHWC, STEM_MODE, SCH_STEM + 0 + 3, true)     // RGB: three windowed channels
    RST_HALO_CASE(16, 128, EPI_NHWC, STEM_MODE, SCH_S2D, true)              // contract_0: 3x3 stride 2, 32 -> 16
    RST_HALO_CASE(32, 64, EPI_NHWC, STEM_MODE, SCH_S2D, true)               // contract_1: 16 -> 32
    RST_HALO_CASE(32, 128, EPI_NHWC, STEM_MODE, SCH_S2D, true)              // deeper contract blocks: 32 -> 32
    RST_HALO_CASE(128, 128, EPI_CONVT2, 0, SCH_T2, true)                    // expand_0: 4 phases x 32 channels
    RST_HALO_CASE(64, 64, EPI_CONVT2, 0, SCH_T2, true)                      // expand_1: 4 phases x 16 channels
    RST_HALO_CASE(16, 128, EPI_QUAD3, MODE_F32, SCH_HEAD, true)             // expand_last: 4 pixels x 3 channels
    RST_HALO_CASE(32, 128, EPI_OCT3, MODE_F32, SCH_HEAD8, true)             // expand_last: 8 pixels x 3 (+1) channels
    RST_HALO_CASE(128, 128, EPI_NHWC, MODE_RELU | MODE_F32 | MODE_TF32, SCH_C3, false)   // tf32 3x3 convs (VGG16 loss model)
    RST_HALO_CASE(64, 128, EPI_NHWC, MODE_RELU | MODE_F32 | MODE_TF32, SCH_C3, false)
    RST_HALO_CASE(32, 128, EPI_NHWC, MODE_RELU | MODE_F32 | MODE_TF32, SCH_C3, false)    // 32-filter bottleneck (rst-960-120-32-3), fp32 inference
    RST_HALO_CASE(128, 128, EPI_NHWC, MODE_F32 | MODE_TF32, SCH_C3, false)               // ... and their input gradients
    RST_HALO_CASE(64, 128, EPI_NHWC, MODE_F32 | MODE_TF32, SCH_C3, false)
#undef RST_HALO_CASE
    return cudaErrorInvalidValue;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;
static std::once_flag g_encode_once;

bool umma_init(std::string* err) {
    std::call_once(g_encode_once, [] {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            g_encode = reinterpret_cast<EncodeTiledFn>(fn);
    });
    if (!g_encode && err) *err = "cuTensorMapEncodeTiled is not available from this driver";
    return g_encode != nullptr;
}

// Dims ordered (C, H, WRU, B) so that a box enumerates rows (h) fastest among row units.
bool encode_halo_map(CUtensorMap* out, const void* base, int B, int H, int WRU, int C, int row_elems, int halo_h,
                     int halo_w, std::string* err, bool atom32) {
    if (!umma_init(err)) return false;
    cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)H, (cuuint64_t)WRU, (cuuint64_t)B};
    cuuint64_t gstr[3] = {(cuuint64_t)WRU * C * 2, (cuuint64_t)C * 2, (cuuint64_t)H * WRU * C * 2};
    cuuint32_t box[4] = {(cuuint32_t)row_elems, (cuuint32_t)halo_h, (cuuint32_t)halo_w, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUtensorMapSwizzle sw = row_elems == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    if (atom32 && row_elems == 64) sw = CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;    // MN-major tf32 operands (wgrad_tf32.cu)
    if (row_elems != 64 && row_elems != 32) {
        if (err) *err = "encode_halo_map: row must be 32 or 64 bf16 elements";
        return false;
    }
    CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gdim, gstr, box, es,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        if (err) *err = "cuTensorMapEncodeTiled(activation) failed with code " + std::to_string((int)r);
        return false;
    }
    return true;
}

bool encode_s2d_map(CUtensorMap* out, const void* base, int B, int H, int W, int C, std::string* err) {
    if (!umma_init(err)) return false;
    if ((H & 1) || (W & 1) || (2 * C != 64 && 2 * C != 32)) {
        if (err) *err = "encode_s2d_map: needs even H, W and 16 or 32 channels";
        return false;
    }
    cuuint64_t gdim[5] = {(cuuint64_t)2 * C, (cuuint64_t)H / 2, 2, (cuuint64_t)W / 2, (cuuint64_t)B};
    cuuint64_t gstr[4] = {(cuuint64_t)2 * W * C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)2 * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[5] = {(cuuint32_t)2 * C, 9, 2, 17, 1};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUtensorMapSwizzle sw = 2 * C == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), gdim, gstr, box, es,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        if (err) *err = "cuTensorMapEncodeTiled(space-to-depth) failed with code " + std::to_string((int)r);
        return false;
    }
    return true;
}

bool encode_weight_map(CUtensorMap* out, const void* base, int nblocks, int N, std::string* err, int box_rows) {
    if (!umma_init(err)) return false;
    cuuint64_t gdim[2] = {64, (cuuint64_t)nblocks * N};
    cuuint64_t gstr[1] = {128};
    cuuint32_t box[2] = {64, (cuuint32_t)(box_rows > 0 ? box_rows : N)};
    cuuint32_t es[2] = {1, 1};
    CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, es,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        if (err) *err = "cuTensorMapEncodeTiled(weights) failed with code " + std::to_string((int)r);
        return false;
    }
    return true;
}

bool encode_weight_unit_map(CUtensorMap* out, const void* base, int rows, std::string* err) {
    if (!umma_init(err)) return false;
    cuuint64_t gdim[2] = {16, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {32};
    cuuint32_t box[2] = {16, 256};
    cuuint32_t es[2] = {1, 1};
    CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, es,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        if (err) *err = "cuTensorMapEncodeTiled(weight units) failed with code " + std::to_string((int)r);
        return false;
    }
    return true;
}

void pack_b_blocks(int total_ksteps, int N, const std::function<float(int, int, int)>& f, std::vector<__nv_bfloat16>* out) {
    const int nblocks = total_ksteps / 4;
    out->assign((size_t)nblocks * N * 64, __float2bfloat16(0.f));
    for (int ks = 0; ks < total_ksteps; ++ks)
        for (int n = 0; n < N; ++n)
            for (int e = 0; e < 16; ++e) {
                float w = f(ks, n, e);
                if (w != 0.f) (*out)[((size_t)(ks / 4) * N + n) * 64 + (ks % 4) * 16 + e] = __float2bfloat16(w);
            }
}

// ------------------------------------------------------------------------------------------------
// bf16 elementwise companions
// ------------------------------------------------------------------------------------------------
__global__ void f32_to_bf16_pad_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, long long total, int c_in,
                                       int c_out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int c = (int)(i % c_out);
    long long pix = i / c_out;
    y[i] = __float2bfloat16(c < c_in ? x[pix * c_in + c] : 0.f);
}
cudaError_t launch_f32_to_bf16_pad(const float* x, __nv_bfloat16* y, long long pixels, int c_in, int c_out, cudaStream_t s) {
    long long total = pixels * c_out;
    if (total == 0) return cudaSuccess;
    f32_to_bf16_pad_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(x, y, total, c_in, c_out);
    return cudaGetLastError();
}

__global__ void bf16_to_f32_slice_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ y, long long total, int c_in,
                                         int c_out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int c = (int)(i % c_out);
    long long pix = i / c_out;
    y[i] = __bfloat162float(x[pix * c_in + c]);
}
cudaError_t launch_bf16_to_f32_slice(const __nv_bfloat16* x, float* y, long long pixels, int c_in, int c_out, cudaStream_t s) {
    long long total = pixels * c_out;
    if (total == 0) return cudaSuccess;
    bf16_to_f32_slice_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(x, y, total, c_in, c_out);
    return cudaGetLastError();
}

// The G-buffer arrives as fp32 (the reference's numpy / TF tensors) or as fp16 (what the EXR planes of the Unreal capture are,
// dataloaders/hdrScreenshots.py:14-29): the pack kernels are instantiated for both.  ld_quad(base, i) returns elements
// 4i .. 4i+3 as floats: one 16-byte load of fp32, one 8-byte load of fp16 (the 64-pixel segments are 16- resp. 8-byte aligned).
__device__ __forceinline__ float in_to_f32(float v) { return v; }
__device__ __forceinline__ float in_to_f32(__half v) { return __half2float(v); }
__device__ __forceinline__ float4 ld_quad(const float* base, int i) { return __ldg(reinterpret_cast<const float4*>(base) + i); }
__device__ __forceinline__ float4 ld_quad(const __half* base, int i) {
    const uint2 r = __ldg(reinterpret_cast<const uint2*>(base) + i);
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&r.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&r.y));
    return make_float4(a.x, a.y, b.x, b.y);
}

// One thread per (pixel, 8-element slot group of the packed row).
template <typename TIn>
__global__ void pack_stem_input_kernel(const TIn* __restrict__ x, __nv_bfloat16* __restrict__ y, int H, int W, int C,
                                       int n_real, int row_elems, int pair_window, long long total_slots) {
    pdl_wait();
    pdl_trigger();
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total_slots) return;
    const int slots = row_elems / 8;
    const int slot = (int)(i % slots);
    const long long pix = i / slots;
    const int px = (int)(pix % W);
    const TIn* xp = x + pix * C;
    float v[8];
    const int e0 = slot * 8;
    const int real_elems = n_real > 0 ? 16 : 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int e = e0 + j;
        float val = 0.f;
        if (e < real_elems) {
            if (e < n_real) val = in_to_f32(xp[e]);
        } else {
            const int g = (e - real_elems) / 16, t = (e - real_elems) % 16;   // virtual group g, horizontal tap t
            const int ch = n_real + g;
            const int taps = pair_window ? ((px & 1) ? 0 : 10) : 9;
            if (ch < C && t < taps) {
                const int sx = px + t - 4;
                if (sx >= 0 && sx < W) val = in_to_f32(xp[(long long)(t - 4) * C + ch]);
            }
        }
        v[j] = val;
    }
    __nv_bfloat162 h0 = __floats2bfloat162_rn(v[0], v[1]), h1 = __floats2bfloat162_rn(v[2], v[3]);
    __nv_bfloat162 h2 = __floats2bfloat162_rn(v[4], v[5]), h3 = __floats2bfloat162_rn(v[6], v[7]);
    uint4 o = make_uint4(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1),
                         *reinterpret_cast<uint32_t*>(&h2), *reinterpret_cast<uint32_t*>(&h3));
    reinterpret_cast<uint4*>(y)[i] = o;
}
// 18-channel pair layout (SCH_STEM2B), any even W: one thread per (pixel pair, 8-element group of the 64-element pair row).
template <typename TIn>
__global__ void pack_stem_pairs18_kernel(const TIn* __restrict__ x, __nv_bfloat16* __restrict__ y, int W, long long total_groups) {
    pdl_wait();
    pdl_trigger();
    constexpr int C = 18;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total_groups) return;
    const int grp = (int)(i & 7);
    const long long pair = i >> 3;
    const int pw = W / 2;
    const int x0 = (int)(pair % pw) * 2;
    const TIn* xp = x + ((pair / pw) * W + x0) * C;                   // the even pixel
    float v[8];
    if (grp < 4) {                                                     // real channels of pixel grp / 2
        const TIn* q = xp + (grp >> 1) * C + (grp & 1) * 8;
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = in_to_f32(q[j]);
    } else {                                                           // window slots t = 8 * (grp & 1) + j of channel 16 + (grp - 4) / 2
        const int ch = 16 + ((grp - 4) >> 1);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int t = (grp & 1) * 8 + j, sx = x0 + t - 4;
            v[j] = (t < 10 && sx >= 0 && sx < W) ? in_to_f32(xp[(long long)(t - 4) * C + ch]) : 0.f;
        }
    }
    __nv_bfloat162 h0 = __floats2bfloat162_rn(v[0], v[1]), h1 = __floats2bfloat162_rn(v[2], v[3]);
    __nv_bfloat162 h2 = __floats2bfloat162_rn(v[4], v[5]), h3 = __floats2bfloat162_rn(v[6], v[7]);
    reinterpret_cast<uint4*>(y)[i] = make_uint4(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1),
                                                *reinterpret_cast<uint32_t*>(&h2), *reinterpret_cast<uint32_t*>(&h3));
}
// Fast variant (W % 64 == 0): one warp per 64-pixel row segment staged in shared memory as in pack_stem_rows_kernel below;
// each lane assembles ONE pixel pair (128 B), the warp writes its 4 KB of packed pairs as contiguous, coalesced runs.
template <typename TIn>
__global__ void __launch_bounds__(256) pack_stem_pair_rows18_kernel(const TIn* __restrict__ x, __nv_bfloat16* __restrict__ y, int W,
                                                                    long long total_segments) {
    pdl_wait();
    pdl_trigger();
    constexpr int C = 18, NF4 = 18 * C;                                // 72 pixels x 18 floats
    extern __shared__ float4 pack_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long seg = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
    if (seg >= total_segments) return;
    const int segs_per_row = W / 64;
    const int sx = (int)(seg % segs_per_row);
    const long long pix0 = (seg / segs_per_row) * W + sx * 64;
    float4* s4 = pack_smem + warp * NF4;
    const TIn* g4 = x + (pix0 - 4) * C;
    const bool left_oob = sx == 0, right_oob = sx == segs_per_row - 1;
#pragma unroll
    for (int i = lane; i < NF4; i += 32) {
        const bool oob = (left_oob && i < C) || (right_oob && i >= 17 * C);
        s4[i] = oob ? make_float4(0.f, 0.f, 0.f, 0.f) : ld_quad(g4, i);
    }
    __syncwarp();
    const float* sf = reinterpret_cast<const float*>(s4);
    const float* px = sf + (4 + 2 * lane) * C;                         // the lane's even pixel; its window starts 4 pixels earlier
    uint32_t w[32];
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            __nv_bfloat162 b = __floats2bfloat162_rn(px[h * C + 2 * j], px[h * C + 2 * j + 1]);
            w[h * 8 + j] = *reinterpret_cast<uint32_t*>(&b);
        }
#pragma unroll
    for (int g = 0; g < 2; ++g) {
#pragma unroll
        for (int j = 0; j < 5; ++j) {
            __nv_bfloat162 b = __floats2bfloat162_rn(px[(2 * j - 4) * C + 16 + g], px[(2 * j - 3) * C + 16 + g]);
            w[16 + 8 * g + j] = *reinterpret_cast<uint32_t*>(&b);
        }
#pragma unroll
        for (int j = 5; j < 8; ++j) w[16 + 8 * g + j] = 0u;
    }
    uint4* so = reinterpret_cast<uint4*>(s4);                          // 32 pairs x 8 vectors = 256 of the slice's 324 float4
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 8; ++j) so[lane * 8 + (j ^ (lane & 7))] = make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
    __syncwarp();
    uint4* out = reinterpret_cast<uint4*>(y + pix0 * 32);
#pragma unroll
    for (int k = lane; k < 256; k += 32) out[k] = so[(k & ~7) + ((k & 7) ^ ((k >> 3) & 7))];
}
// Fast path for the G-buffer layouts (16 real channels + NV windowed ones, W % 64 == 0): one warp per 64-pixel row segment.
// The segment plus a 4-pixel halo on each side is 72*C floats = 18*C aligned float4 (coalesced loads) staged in the warp's
// own shared-memory slice; each lane then assembles two packed pixels (stride-C reads are bank-conflict free for C = 17).
template <int NV, bool PAIRW, typename TIn>
__global__ void __launch_bounds__(256) pack_stem_rows_kernel(const TIn* __restrict__ x, __nv_bfloat16* __restrict__ y, int W,
                                                             long long total_segments) {
    pdl_wait();
    pdl_trigger();
    constexpr int C = 16 + NV, ROW = NV <= 1 ? 32 : 64, NF4 = 18 * C;
    extern __shared__ float4 pack_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long seg = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
    if (seg >= total_segments) return;
    const int segs_per_row = W / 64;
    const int sx = (int)(seg % segs_per_row);
    const long long row = seg / segs_per_row;                      // n*H + y
    const long long pix0 = row * W + sx * 64;                      // first pixel of the segment
    float4* s4 = pack_smem + warp * NF4;
    const TIn* g4 = x + (pix0 - 4) * C;
    const bool left_oob = sx == 0, right_oob = sx == segs_per_row - 1;
#pragma unroll
    for (int i = lane; i < NF4; i += 32) {
        const bool oob = (left_oob && i < C) || (right_oob && i >= 17 * C);     // 4 pixels = C float4 on each side
        s4[i] = oob ? make_float4(0.f, 0.f, 0.f, 0.f) : ld_quad(g4, i);
    }
    __syncwarp();
    const float* sf = reinterpret_cast<const float*>(s4);
    constexpr int V = ROW / 8;                                     // 16-byte vectors per packed pixel
    uint32_t w[2][ROW / 2];
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int pl = lane + 32 * half;                           // pixel within the segment
        const float* px = sf + (4 + pl) * C;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            __nv_bfloat162 h = __floats2bfloat162_rn(px[2 * j], px[2 * j + 1]);
            w[half][j] = *reinterpret_cast<uint32_t*>(&h);
        }
#pragma unroll
        for (int g = 0; g < NV; ++g) {
            float v[10];
#pragma unroll
            for (int t = 0; t < 9; ++t) v[t] = sf[(pl + t) * C + 16 + g];        // channel 16+g at x + t - 4
            v[9] = 0.f;
            if (PAIRW) {                              // window shared by the pixel pair: 10 values in the even pixel, none in the odd
                const float keep = (pl & 1) ? 0.f : 1.f;                         // branch-free: lanes alternate even / odd pixels
                v[9] = sf[min(pl + 9, 71) * C + 16 + g];                         // even pl <= 62 stays inside the 72-pixel slice
#pragma unroll
                for (int t = 0; t < 10; ++t) v[t] *= keep;
            }
#pragma unroll
            for (int j = 0; j < 5; ++j) {
                __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
                w[half][8 + 8 * g + j] = *reinterpret_cast<uint32_t*>(&h);
            }
#pragma unroll
            for (int j = 5; j < 8; ++j) w[half][8 + 8 * g + j] = 0u;
        }
#pragma unroll
        for (int j = 8 + 8 * NV; j < ROW / 2; ++j) w[half][j] = 0u;
    }
    // transpose through the (now consumed) staging slice so that the warp writes its packed pixels as contiguous, fully
    // coalesced runs: all 64 pixels at once when they fit in the slice (17 channels: 4 KB of 4.8 KB), else 32 at a time
    // (18 channels: 2 x 4 KB of 5.1 KB); the vector index is XOR-swizzled by the pixel to keep the stores conflict-free
    uint4* so = reinterpret_cast<uint4*>(s4);
    constexpr int PASSES = 64 * V <= NF4 ? 1 : 2, PPP = 64 / PASSES;       // pixels per pass
    auto slot = [](int pl, int j) { return pl * V + (j ^ ((pl >> (V == 4 ? 1 : 0)) & (V - 1))); };
#pragma unroll
    for (int pass = 0; pass < PASSES; ++pass) {
        __syncwarp();
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            if (PASSES == 2 && half != pass) continue;
            const int pl = PASSES == 2 ? lane : lane + 32 * half;
#pragma unroll
            for (int j = 0; j < V; ++j)
                so[slot(pl, j)] = make_uint4(w[half][4 * j], w[half][4 * j + 1], w[half][4 * j + 2], w[half][4 * j + 3]);
        }
        __syncwarp();
        uint4* out = reinterpret_cast<uint4*>(y + (pix0 + PPP * pass) * ROW);
#pragma unroll
        for (int k = lane; k < PPP * V; k += 32) out[k] = so[slot(k / V, k % V)];
    }
}

template <typename TIn>
static cudaError_t launch_pack_stem_input_t(const TIn* x, __nv_bfloat16* y, int B, int H, int W, int C, int n_real, int row_elems,
                                            int pair_window, cudaStream_t s) {
    long long total = (long long)B * H * W * (row_elems / 8);
    if (total == 0) return cudaSuccess;
    const int nv = C - 16;
    // a 64-pixel segment minus its 4-pixel halo starts at a multiple of 4 elements: 16-byte (fp32) / 8-byte (fp16) vectors
    const bool aligned = (reinterpret_cast<uintptr_t>(x) & (4 * sizeof(TIn) - 1)) == 0;
    if (pair_window && ((nv != 1 && nv != 2) || n_real != 16 || (W & 1))) return cudaErrorInvalidValue;
    if (pair_window && nv == 2) {
        if (row_elems != 32) return cudaErrorInvalidValue;
        if (W % 64 == 0 && aligned) {
            const long long segments = (long long)B * H * (W / 64);
            (void)launch_pdl(true, pack_stem_pair_rows18_kernel<TIn>, dim3((unsigned)((segments + 7) / 8)), dim3(256), (size_t)8 * 18 * 18 * sizeof(float4), s, x, y, W, segments);
        } else {
            (void)launch_pdl(true, pack_stem_pairs18_kernel<TIn>, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, s, x, y, W, total);
        }
        return cudaGetLastError();
    }
    if (n_real == 16 && (nv == 1 || nv == 2) && W % 64 == 0 && row_elems == (nv <= 1 ? 32 : 64) && aligned) {
        const long long segments = (long long)B * H * (W / 64);
        const unsigned blocks = (unsigned)((segments + 7) / 8);
        const size_t smem = (size_t)8 * 18 * C * sizeof(float4);
        if (nv == 1 && pair_window) (void)launch_pdl(true, pack_stem_rows_kernel<1, true, TIn>, dim3(blocks), dim3(256), smem, s, x, y, W, segments);
        else if (nv == 1) (void)launch_pdl(true, pack_stem_rows_kernel<1, false, TIn>, dim3(blocks), dim3(256), smem, s, x, y, W, segments);
        else (void)launch_pdl(true, pack_stem_rows_kernel<2, false, TIn>, dim3(blocks), dim3(256), smem, s, x, y, W, segments);
        return cudaGetLastError();
    }
    (void)launch_pdl(true, pack_stem_input_kernel<TIn>, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, s, x, y, H, W, C, n_real, row_elems, pair_window, total);
    return cudaGetLastError();
}

cudaError_t launch_pack_stem_input(const void* x, int x_f16, __nv_bfloat16* y, int B, int H, int W, int C, int n_real, int row_elems,
                                   int pair_window, cudaStream_t s) {
    if (x_f16) return launch_pack_stem_input_t(reinterpret_cast<const __half*>(x), y, B, H, W, C, n_real, row_elems, pair_window, s);
    return launch_pack_stem_input_t(reinterpret_cast<const float*>(x), y, B, H, W, C, n_real, row_elems, pair_window, s);
}

// Instance-norm apply, VEC channels per thread-iteration (8 for bf16 input, 4 or fewer for fp32 input).
template <bool XF32, bool YF32>
__global__ void __launch_bounds__(256) cin_apply_v_kernel(const CinApplyV p, int pix_per_block) {
    pdl_wait();
    pdl_trigger();
    extern __shared__ float smf[];
    const int C = p.C;
    float* s_inv = smf;
    float* s_nmi = smf + C;
    float* s_scale = smf + 2 * C;                    // [S][C]
    float* s_bias = s_scale + p.num_styles * C;      // [S][C]
    const int n = blockIdx.y;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        double sum = p.stats[((long long)n * C + c) * 2 + 0];
        double sq = p.stats[((long long)n * C + c) * 2 + 1];
        double mean = sum / (double)p.P;
        double var = sq / (double)p.P - mean * mean;
        if (var < 0.0) var = 0.0;
        float inv = rsqrtf((float)var + p.eps);
        s_inv[c] = inv;
        s_nmi[c] = -(float)mean * inv;
        for (int st = 0; st < p.num_styles; ++st) {
            const float* ps = p.params + n * p.param_bstride + st * p.param_sstride;
            s_scale[st * C + c] = ps[p.scale_off + c];
            s_bias[st * C + c] = ps[p.bias_off + c];
        }
    }
    __syncthreads();
    const bool blend = p.num_styles == 2 && p.weights != nullptr;
    if (XF32) {
        // fp32 input (3-channel head): scalar path, one element per thread-iteration
        const long long base = (long long)n * p.P * C;
        const long long e0 = (long long)blockIdx.x * pix_per_block * C;
        const long long e1 = min((long long)p.P * C, e0 + (long long)pix_per_block * C);
        for (long long e = e0 + threadIdx.x; e < e1; e += blockDim.x) {
            const int c = (int)(e % C);
            float scale = s_scale[c], bias = s_bias[c];
            if (blend) {
                const float2 w = *reinterpret_cast<const float2*>(p.weights + ((long long)n * p.P + e / C) * 2);
                scale = s_scale[c] * w.x + s_scale[C + c] * w.y;
                bias = s_bias[c] * w.x + s_bias[C + c] * w.y;
            }
            float xh = reinterpret_cast<const float*>(p.x)[base + e] * s_inv[c] + s_nmi[c];
            float o = actf(bias + xh * scale, p.act);
            if (YF32) reinterpret_cast<float*>(p.y)[base + e] = o;
            else reinterpret_cast<__nv_bfloat16*>(p.y)[base + e] = __float2bfloat16(o);
        }
        return;
    }
    const int vec_per_pix = C >> 3;
    const long long base = (long long)n * p.P * vec_per_pix;
    const long long v0 = (long long)blockIdx.x * pix_per_block * vec_per_pix;
    const long long v1 = min((long long)p.P * vec_per_pix, v0 + (long long)pix_per_block * vec_per_pix);
    const uint4* x4 = reinterpret_cast<const uint4*>(p.x) + base;
    const uint4* r4 = p.residual ? reinterpret_cast<const uint4*>(p.residual) + base : nullptr;
    for (long long v = v0 + threadIdx.x; v < v1; v += blockDim.x) {
        const int c0 = (int)(v % vec_per_pix) * 8;
        uint4 xin = x4[v];
        uint4 rin = r4 ? r4[v] : make_uint4(0, 0, 0, 0);
        float w0 = 1.f, w1 = 0.f;
        if (blend) {
            const long long pix = v / vec_per_pix;
            const float2 w = *reinterpret_cast<const float2*>(p.weights + ((long long)n * p.P + pix) * 2);
            w0 = w.x; w1 = w.y;
        }
        const __nv_bfloat162* xb = reinterpret_cast<const __nv_bfloat162*>(&xin);
        const __nv_bfloat162* rb = reinterpret_cast<const __nv_bfloat162*>(&rin);
        float o[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float2 xv = __bfloat1622float2(xb[j]);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int c = c0 + 2 * j + h;
                float scale, bias;
                if (blend) {
                    scale = s_scale[c] * w0 + s_scale[C + c] * w1;
                    bias = s_bias[c] * w0 + s_bias[C + c] * w1;
                } else {
                    scale = s_scale[c];
                    bias = s_bias[c];
                }
                float xh = (h ? xv.y : xv.x) * s_inv[c] + s_nmi[c];
                o[2 * j + h] = actf(bias + xh * scale, p.act);
            }
            if (r4) {
                float2 rv = __bfloat1622float2(rb[j]);
                o[2 * j] += rv.x; o[2 * j + 1] += rv.y;
            }
        }
        if (YF32) {
            float4* y4 = reinterpret_cast<float4*>(p.y) + (base + v) * 2;
            y4[0] = make_float4(o[0], o[1], o[2], o[3]);
            y4[1] = make_float4(o[4], o[5], o[6], o[7]);
        } else {
            __nv_bfloat162 h0 = __floats2bfloat162_rn(o[0], o[1]), h1 = __floats2bfloat162_rn(o[2], o[3]);
            __nv_bfloat162 h2 = __floats2bfloat162_rn(o[4], o[5]), h3 = __floats2bfloat162_rn(o[6], o[7]);
            reinterpret_cast<uint4*>(p.y)[base + v] = make_uint4(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1),
                                                                 *reinterpret_cast<uint32_t*>(&h2), *reinterpret_cast<uint32_t*>(&h3));
        }
    }
}

// ---- fast paths ------------------------------------------------------------------------------------
// bf16 -> bf16, C in {16,32,64,128}: every thread owns a fixed group of 8 channels (blockDim % (C/8) == 0), keeps the
// fused coefficients y = x*a + b in registers and streams 4 independent 16-byte vectors per iteration.
template <bool BLEND, bool RES>
__global__ void __launch_bounds__(256) cin_apply_fast_kernel(const CinApplyV p, int pix_per_block) {
    pdl_wait();
    pdl_trigger();
    const int C = p.C, n = blockIdx.y;
    const int vec_per_pix = C >> 3;
    const int c0 = (threadIdx.x % vec_per_pix) * 8;
    const double inv_p = 1.0 / (double)p.P;
    float a0[8], b0[8], a1[8], b1[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = c0 + j;
        const double2 st = *reinterpret_cast<const double2*>(p.stats + ((long long)n * C + c) * 2);
        const double mean = st.x * inv_p;
        double var = fma(st.y, inv_p, -mean * mean);
        if (var < 0.0) var = 0.0;
        const float inv = rsqrtf((float)var + p.eps), nmi = -(float)mean * inv;
        const float* ps = p.params + n * p.param_bstride;
        a0[j] = inv * ps[p.scale_off + c];
        b0[j] = ps[p.bias_off + c] + nmi * ps[p.scale_off + c];
        if (BLEND) {
            // the weight map of this path is always (1 - w, w) (styleTransfer.py:297-302 and its average-pooled mips), so
            // w0*p0 + w1*p1 = p0 + w1*(p1 - p0): the differences are kept, one fma per coefficient and pixel instead of two ops
            const float* p1 = ps + p.param_sstride;
            a1[j] = inv * p1[p.scale_off + c] - a0[j];
            b1[j] = (p1[p.bias_off + c] + nmi * p1[p.scale_off + c]) - b0[j];
        }
    }
    const int vpp_shift = 31 - __clz(vec_per_pix);            // C in {16, 32, 64, 128}: a power of two
    const long long base = (long long)n * p.P * vec_per_pix;
    const long long v0 = (long long)blockIdx.x * pix_per_block * vec_per_pix;
    const long long v1 = min((long long)p.P * vec_per_pix, v0 + (long long)pix_per_block * vec_per_pix);
    const uint4* x4 = reinterpret_cast<const uint4*>(p.x) + base;
    const uint4* r4 = RES ? reinterpret_cast<const uint4*>(p.residual) + base : nullptr;
    uint4* y4 = reinterpret_cast<uint4*>(p.y) + base;
    const float2* w2 = BLEND ? reinterpret_cast<const float2*>(p.weights) + (long long)n * p.P : nullptr;
    const int act = p.act;
    auto one = [&](uint4 xin, uint4 rin, float2 w) -> uint4 {
        const __nv_bfloat162* xb = reinterpret_cast<const __nv_bfloat162*>(&xin);
        const __nv_bfloat162* rb = reinterpret_cast<const __nv_bfloat162*>(&rin);
        uint4 outv;
        __nv_bfloat162* ob = reinterpret_cast<__nv_bfloat162*>(&outv);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 xv = __bfloat1622float2(xb[j]);
            float aa0 = a0[2 * j], bb0 = b0[2 * j], aa1 = a0[2 * j + 1], bb1 = b0[2 * j + 1];
            if (BLEND) {
                aa0 = fmaf(w.y, a1[2 * j], aa0);           bb0 = fmaf(w.y, b1[2 * j], bb0);
                aa1 = fmaf(w.y, a1[2 * j + 1], aa1);       bb1 = fmaf(w.y, b1[2 * j + 1], bb1);
            }
            float o0 = fmaf(xv.x, aa0, bb0), o1 = fmaf(xv.y, aa1, bb1);
            if (act == ACT_RELU) { o0 = fmaxf(o0, 0.f); o1 = fmaxf(o1, 0.f); }
            if (RES) { const float2 rv = __bfloat1622float2(rb[j]); o0 += rv.x; o1 += rv.y; }
            ob[j] = __floats2bfloat162_rn(o0, o1);
        }
        return outv;
    };
    const long long stride = blockDim.x;
    long long v = v0 + threadIdx.x;
    const uint4 z = make_uint4(0, 0, 0, 0);
    for (; v + 3 * stride < v1; v += 4 * stride) {
        uint4 xi[4], ri[4];
        float2 wi[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            xi[u] = x4[v + u * stride];
            ri[u] = RES ? r4[v + u * stride] : z;
            wi[u] = BLEND ? w2[(v + u * stride) >> vpp_shift] : make_float2(1.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) y4[v + u * stride] = one(xi[u], ri[u], wi[u]);
    }
    for (; v < v1; v += stride)
        y4[v] = one(x4[v], RES ? r4[v] : z, BLEND ? w2[v >> vpp_shift] : make_float2(1.f, 0.f));
}

// Same arithmetic, operands staged by the bulk-copy engine: a producer warp streams 16-byte-vector chunks of x (and of the skip
// tensor) into a ring of shared-memory stages with cp.async.bulk + mbarrier transaction counts, eight consumer warps read their
// vectors from shared memory and store the result straight to global memory.  The loads in flight no longer live in registers
// (stages x chunk bytes per CTA instead of 4 vectors per thread), which is what the register version is short of when the
// tensors sit in L2 (profiles/r02_00_summary.md: 35 % warps active, issue 30 %, ~6 TB/s of a ~12 TB/s L2).
// In place (y == x) is safe: a chunk is stored only after its own bulk load completed, other chunks are other addresses.
constexpr int kBulkMaxStages = 8;
constexpr int kBulkConsumers = 256;

__device__ __forceinline__ void bulk_load_1d_hint(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                     umma::smem_u32(dst_smem)),
                 "l"((uint64_t)__cvta_generic_to_global(src)), "r"(bytes), "r"(umma::smem_u32(bar)), "l"(policy)
                 : "memory");
}
__device__ __forceinline__ void bulk_load_1d(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     umma::smem_u32(dst_smem)),
                 "l"((uint64_t)__cvta_generic_to_global(src)), "r"(bytes), "r"(umma::smem_u32(bar))
                 : "memory");
}

template <bool BLEND, bool RES>
__global__ void __launch_bounds__(kBulkConsumers + 32) cin_apply_bulk_kernel(const CinApplyV p, int pix_per_block, int stages,
                                                                             int chunk_vecs) {
    extern __shared__ __align__(128) unsigned char bulk_smem[];
    __shared__ uint64_t full[kBulkMaxStages], empty[kBulkMaxStages];
    const int C = p.C, n = blockIdx.y, tid = threadIdx.x;
    const int vec_per_pix = C >> 3;
    const long long base = (long long)n * p.P * vec_per_pix;
    const long long v0 = (long long)blockIdx.x * pix_per_block * vec_per_pix;
    const long long v1 = min((long long)p.P * vec_per_pix, v0 + (long long)pix_per_block * vec_per_pix);
    const int nchunks = (int)((v1 - v0 + chunk_vecs - 1) / chunk_vecs);
    uint4* xs = reinterpret_cast<uint4*>(bulk_smem);
    uint4* rs = xs + (size_t)stages * chunk_vecs;
    if (tid == 0) {
        for (int i = 0; i < stages; ++i) { umma::mbar_init(&full[i], 1); umma::mbar_init(&empty[i], kBulkConsumers / 32); }
        umma::fence_barrier_init();
    }
    // the style parameters are inputs of the forward (written long before this kernel's predecessor): fetched ahead of the wait
    const int c0 = (tid % vec_per_pix) * 8;
    float sc0[8], bi0[8], sc1[BLEND ? 8 : 1], bi1[BLEND ? 8 : 1];
    {
        const float* ps = p.params + n * p.param_bstride;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            sc0[j] = ps[p.scale_off + c0 + j]; bi0[j] = ps[p.bias_off + c0 + j];
            if (BLEND) { sc1[j] = ps[p.param_sstride + p.scale_off + c0 + j]; bi1[j] = ps[p.param_sstride + p.bias_off + c0 + j]; }
        }
    }
    pdl_wait();
    pdl_trigger();
    __syncthreads();
    if (tid >= kBulkConsumers) {                                   // ---- producer warp
        if (tid == kBulkConsumers) {
            const uint4* x4 = reinterpret_cast<const uint4*>(p.x) + base;
            const uint4* r4 = RES ? reinterpret_cast<const uint4*>(p.residual) + base : nullptr;
            const uint64_t pol_first = umma::l2_policy_evict_first();
            for (int k = 0; k < nchunks; ++k) {
                const int st = k % stages;
                if (k >= stages) umma::mbar_wait(&empty[st], ((k / stages) - 1) & 1);
                const long long v = v0 + (long long)k * chunk_vecs;
                const uint32_t bytes = (uint32_t)min((long long)chunk_vecs, v1 - v) * 16u;
                umma::mbar_expect_tx(&full[st], RES ? 2 * bytes : bytes);
                bulk_load_1d(xs + (size_t)st * chunk_vecs, x4 + v, bytes, &full[st]);
                if (RES) {
                    if (p.l2_hints & 1) bulk_load_1d_hint(rs + (size_t)st * chunk_vecs, r4 + v, bytes, &full[st], pol_first);
                    else bulk_load_1d(rs + (size_t)st * chunk_vecs, r4 + v, bytes, &full[st]);
                }
            }
        }
        return;
    }
    // ---- consumers: fixed group of 8 channels per thread (kBulkConsumers % vec_per_pix == 0)
    const double inv_p = 1.0 / (double)p.P;
    float a0[8], b0[8], a1[8], b1[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = c0 + j;
        const double2 st = *reinterpret_cast<const double2*>(p.stats + ((long long)n * C + c) * 2);
        const double mean = st.x * inv_p;
        double var = fma(st.y, inv_p, -mean * mean);
        if (var < 0.0) var = 0.0;
        const float inv = rsqrtf((float)var + p.eps), nmi = -(float)mean * inv;
        a0[j] = inv * sc0[j];
        b0[j] = bi0[j] + nmi * sc0[j];
        if (BLEND) {
            a1[j] = inv * sc1[j] - a0[j];
            b1[j] = (bi1[j] + nmi * sc1[j]) - b0[j];
        }
    }
    const int vpp_shift = 31 - __clz(vec_per_pix);
    uint4* y4 = reinterpret_cast<uint4*>(p.y) + base;
    const float2* w2 = BLEND ? reinterpret_cast<const float2*>(p.weights) + (long long)n * p.P : nullptr;
    const int act = p.act;
    const int vpt = chunk_vecs / kBulkConsumers;                   // vectors per thread and chunk (launcher: 1..8)
    for (int k = 0; k < nchunks; ++k) {
        const int st = k % stages;
        const long long v = v0 + (long long)k * chunk_vecs;
        const int nv = (int)min((long long)chunk_vecs, v1 - v);
        float wy[8];
        if (BLEND) {                                               // weight map: plain loads issued before the wait
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (u < vpt) { const int i = tid + u * kBulkConsumers; wy[u] = i < nv ? w2[(v + i) >> vpp_shift].y : 0.f; }
        }
        umma::mbar_wait(&full[st], (k / stages) & 1);
        const uint4* xc = xs + (size_t)st * chunk_vecs;
        const uint4* rc = rs + (size_t)st * chunk_vecs;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (u >= vpt) break;
            const int i = tid + u * kBulkConsumers;
            if (i < nv) {
                const uint4 xin = xc[i];
                uint4 rin = make_uint4(0, 0, 0, 0);
                if (RES) rin = rc[i];
                const __nv_bfloat162* xb = reinterpret_cast<const __nv_bfloat162*>(&xin);
                const __nv_bfloat162* rb = reinterpret_cast<const __nv_bfloat162*>(&rin);
                uint4 outv;
                __nv_bfloat162* ob = reinterpret_cast<__nv_bfloat162*>(&outv);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 xv = __bfloat1622float2(xb[j]);
                    float aa0 = a0[2 * j], bb0 = b0[2 * j], aa1 = a0[2 * j + 1], bb1 = b0[2 * j + 1];
                    if (BLEND) {
                        aa0 = fmaf(wy[u], a1[2 * j], aa0);           bb0 = fmaf(wy[u], b1[2 * j], bb0);
                        aa1 = fmaf(wy[u], a1[2 * j + 1], aa1);       bb1 = fmaf(wy[u], b1[2 * j + 1], bb1);
                    }
                    float o0 = fmaf(xv.x, aa0, bb0), o1 = fmaf(xv.y, aa1, bb1);
                    if (act == ACT_RELU) { o0 = fmaxf(o0, 0.f); o1 = fmaxf(o1, 0.f); }
                    if (RES) { const float2 rv = __bfloat1622float2(rb[j]); o0 += rv.x; o1 += rv.y; }
                    ob[j] = __floats2bfloat162_rn(o0, o1);
                }
                y4[v + i] = outv;
            }
        }
        __syncwarp();
        if ((tid & 31) == 0) umma::mbar_arrive(&empty[st]);        // this warp no longer reads the stage
    }
}

template <bool BLEND, bool RES>
static cudaError_t launch_cin_apply_bulk(const CinApplyV& p, int pix_per_block, int stages, int chunk_vecs, cudaStream_t s) {
    const size_t smem = (size_t)stages * chunk_vecs * 16 * (RES ? 2 : 1);
    static SmemAttrCache configured;                               // per device: a process may drive several GPUs
    if (cudaError_t e = ensure_dynamic_smem(cin_apply_bulk_kernel<BLEND, RES>, 200 * 1024, configured)) return e;
    dim3 grid((unsigned)ceil_div(p.P, pix_per_block), (unsigned)p.B);
    (void)launch_pdl(true, cin_apply_bulk_kernel<BLEND, RES>, dim3(grid), dim3(kBulkConsumers + 32), smem, s, p, pix_per_block, stages, chunk_vecs);
    return cudaGetLastError();
}

// fp32 -> fp32 with 3 channels (the image head): 4 consecutive floats per thread, sigmoid.
// YU8: the image leaves as uint8 = trunc(255 * y), the quantisation the reference's callers apply to the prediction
// (predict_using_checkpoint.py:99 `np.uint8(... * 255)`, predict_video_using_checkpoint.py:98 `(... * 255).astype(int)`).
template <bool BLEND, bool YU8>
__global__ void __launch_bounds__(256) cin_apply_c3_kernel(const CinApplyV p) {
    pdl_wait();
    pdl_trigger();
    __shared__ float sa[2][3], sb[2][3];
    const int n = blockIdx.y;
    if (threadIdx.x < 3) {
        const int c = threadIdx.x;
        const double inv_p = 1.0 / (double)p.P;
        const double2 st = *reinterpret_cast<const double2*>(p.stats + ((long long)n * 3 + c) * 2);
        const double mean = st.x * inv_p;
        double var = fma(st.y, inv_p, -mean * mean);
        if (var < 0.0) var = 0.0;
        const float inv = rsqrtf((float)var + p.eps), nmi = -(float)mean * inv;
        for (int st_ = 0; st_ < p.num_styles; ++st_) {
            const float* ps = p.params + n * p.param_bstride + st_ * p.param_sstride;
            sa[st_][c] = inv * ps[p.scale_off + c];
            sb[st_][c] = ps[p.bias_off + c] + nmi * ps[p.scale_off + c];
        }
    }
    __syncthreads();
    // 4 pixels = 12 floats = 3 float4 per thread-iteration: the channel of every lane of every vector is a compile-time constant
    const long long groups = (long long)p.P / 4;                     // P % 4 == 0 (checked by the launcher)
    const float4* x4 = reinterpret_cast<const float4*>(p.x) + (long long)n * groups * 3;
    float4* y4 = reinterpret_cast<float4*>(p.y) + (long long)n * groups * 3;
    uint32_t* y8 = reinterpret_cast<uint32_t*>(p.y) + (long long)n * groups * 3;      // YU8: 12 bytes per 4-pixel group
    const float2* w2 = BLEND ? reinterpret_cast<const float2*>(p.weights) + (long long)n * p.P : nullptr;
    float a[3], b[3], a1[3], b1[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) { a[c] = sa[0][c]; b[c] = sb[0][c]; a1[c] = BLEND ? sa[1][c] : 0.f; b1[c] = BLEND ? sb[1][c] : 0.f; }
    const int act = p.act;
    auto one = [&](const float4 (&in)[3], long long g, float4 (&out)[3]) {
        const float* xi = reinterpret_cast<const float*>(in);
        float* xo = reinterpret_cast<float*>(out);
#pragma unroll
        for (int e = 0; e < 12; ++e) {
            const int c = e % 3;
            float aa = a[c], bb = b[c];
            if (BLEND) {
                const float2 w = w2[g * 4 + e / 3];
                aa = aa * w.x + a1[c] * w.y;
                bb = bb * w.x + b1[c] * w.y;
            }
            const float t = fmaf(xi[e], aa, bb);
            xo[e] = act == ACT_SIGMOID ? 1.f / (1.f + __expf(-t)) : (act == ACT_RELU ? fmaxf(t, 0.f) : t);
        }
    };
    auto store = [&](long long grp, const float4 (&out)[3]) {
        if (!YU8) {
#pragma unroll
            for (int k = 0; k < 3; ++k) y4[grp * 3 + k] = out[k];
            return;
        }
        const float* xo = reinterpret_cast<const float*>(out);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            uint32_t word = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) word |= min(__float2uint_rz(fmaxf(xo[4 * k + j], 0.f) * 255.f), 255u) << (8 * j);
            y8[grp * 3 + k] = word;
        }
    };
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; g + stride < groups; g += 2 * stride) {                   // two groups (6 independent 16-byte loads) in flight
        float4 i0[3], i1[3], o0[3], o1[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) { i0[k] = __ldg(x4 + g * 3 + k); i1[k] = __ldg(x4 + (g + stride) * 3 + k); }
        one(i0, g, o0);
        one(i1, g + stride, o1);
        store(g, o0);
        store(g + stride, o1);
    }
    for (; g < groups; g += stride) {
        float4 i0[3], o0[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) i0[k] = __ldg(x4 + g * 3 + k);
        one(i0, g, o0);
        store(g, o0);
    }
}

cudaError_t launch_cin_apply_v(const CinApplyV& p, cudaStream_t s) {
    if (p.B == 0 || p.P == 0) return cudaSuccess;
    if (!p.x_f32 && p.C % 8 != 0) return cudaErrorInvalidValue;
    if (p.x_f32 && p.residual) return cudaErrorInvalidValue;
    const bool blend = p.num_styles == 2 && p.weights != nullptr;
    if (!p.x_f32 && !p.y_f32 && p.act != ACT_SIGMOID && (p.C == 16 || p.C == 32 || p.C == 64 || p.C == 128)) {
        // Tensors of the batch-8 trunk / decoder size go through the bulk-copy kernel, one wave of ~2 CTAs per SM, each streaming
        // a whole number of chunks (B200, profiles/r02_04_pdl_and_bulk_norm.md); small ones keep the register kernel, whose short CTAs
        // start faster.  RST_NORM_BULK=0/1 forces one of them, RST_NORM_PPB the elements per CTA (A/B runs).
        static const int bulk_mode = ab_env("RST_NORM_BULK") ? atoi(ab_env("RST_NORM_BULK")) : -1;
        static const int ppb_env = ab_env("RST_NORM_PPB") ? atoi(ab_env("RST_NORM_PPB")) : 0;
        const long long tensor_bytes = (long long)p.B * p.P * p.C * 2;
        const bool bulk = bulk_mode < 0 ? tensor_bytes >= (24ll << 20) : bulk_mode != 0;
        if (bulk) {
            static const int stages_x = ab_env("RST_NORM_STAGES") ? atoi(ab_env("RST_NORM_STAGES")) : 4;
            static const int stages_r = ab_env("RST_NORM_STAGES_RES") ? atoi(ab_env("RST_NORM_STAGES_RES")) : 3;
            static const int chunk_vecs = ab_env("RST_NORM_CHUNK") ? atoi(ab_env("RST_NORM_CHUNK")) : 1024;
            const int stages = p.residual ? stages_r : stages_x;
            if (stages < 1 || stages > kBulkMaxStages || chunk_vecs % kBulkConsumers || chunk_vecs < kBulkConsumers ||
                chunk_vecs > 8 * kBulkConsumers || (size_t)stages * chunk_vecs * 32 > 200 * 1024)
                return cudaErrorInvalidValue;
            static int sms_of_device[64] = {};
            int dev = 0, num_sms = 148;
            if (cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < 64) {
                if (!sms_of_device[dev] &&
                    cudaDeviceGetAttribute(&sms_of_device[dev], cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
                    sms_of_device[dev] = 148;
                num_sms = sms_of_device[dev];
            }
            const int chunk_pix = chunk_vecs / (p.C >> 3);
            int pix_per_block = ppb_env ? max(chunk_pix, ppb_env / p.C)
                                        : ceil_div(ceil_div(p.P, max(1, 2 * num_sms / p.B)), chunk_pix) * chunk_pix;
            if (blend && p.residual) return launch_cin_apply_bulk<true, true>(p, pix_per_block, stages, chunk_vecs, s);
            if (blend) return launch_cin_apply_bulk<true, false>(p, pix_per_block, stages, chunk_vecs, s);
            if (p.residual) return launch_cin_apply_bulk<false, true>(p, pix_per_block, stages, chunk_vecs, s);
            return launch_cin_apply_bulk<false, false>(p, pix_per_block, stages, chunk_vecs, s);
        }
        const int pix_per_block = max(64, (ppb_env ? ppb_env : 65536) / p.C);       // swept on B200 (profiles/r01_03_experiments.md): ~3 CTAs per SM is the optimum
        dim3 grid((unsigned)ceil_div(p.P, pix_per_block), (unsigned)p.B);
        if (blend && p.residual) (void)launch_pdl(true, cin_apply_fast_kernel<true, true>, dim3(grid), dim3(256), 0, s, p, pix_per_block);
        else if (blend) (void)launch_pdl(true, cin_apply_fast_kernel<true, false>, dim3(grid), dim3(256), 0, s, p, pix_per_block);
        else if (p.residual) (void)launch_pdl(true, cin_apply_fast_kernel<false, true>, dim3(grid), dim3(256), 0, s, p, pix_per_block);
        else (void)launch_pdl(true, cin_apply_fast_kernel<false, false>, dim3(grid), dim3(256), 0, s, p, pix_per_block);
        return cudaGetLastError();
    }
    if (p.y_u8 && !(p.x_f32 && p.C == 3 && p.P % 4 == 0)) return cudaErrorInvalidValue;   // uint8 only for the 3-channel image head
    if (p.x_f32 && (p.y_f32 || p.y_u8) && p.C == 3 && p.P % 4 == 0) {
        const long long want = ((long long)p.P / 4 + 255) / 256;
        const long long per_sample = max(1, 148 * 4 / p.B);          // ~4 CTAs per SM in total, each streaming a long run
        dim3 grid((unsigned)(want < per_sample ? want : per_sample), (unsigned)p.B);
        if (blend && p.y_u8) (void)launch_pdl(true, cin_apply_c3_kernel<true, true>, dim3(grid), dim3(256), 0, s, p);
        else if (blend) (void)launch_pdl(true, cin_apply_c3_kernel<true, false>, dim3(grid), dim3(256), 0, s, p);
        else if (p.y_u8) (void)launch_pdl(true, cin_apply_c3_kernel<false, true>, dim3(grid), dim3(256), 0, s, p);
        else (void)launch_pdl(true, cin_apply_c3_kernel<false, false>, dim3(grid), dim3(256), 0, s, p);
        return cudaGetLastError();
    }
    int pix_per_block = max(1, 32768 / p.C);
    dim3 grid((unsigned)ceil_div(p.P, pix_per_block), (unsigned)p.B);
    size_t smem = (size_t)(2 + 2 * p.num_styles) * p.C * sizeof(float);
    if (p.x_f32 && p.y_f32) (void)launch_pdl(true, cin_apply_v_kernel<true, true>, dim3(grid), dim3(256), smem, s, p, pix_per_block);
    else if (p.x_f32) (void)launch_pdl(true, cin_apply_v_kernel<true, false>, dim3(grid), dim3(256), smem, s, p, pix_per_block);
    else if (p.y_f32) (void)launch_pdl(true, cin_apply_v_kernel<false, true>, dim3(grid), dim3(256), smem, s, p, pix_per_block);
    else (void)launch_pdl(true, cin_apply_v_kernel<false, false>, dim3(grid), dim3(256), smem, s, p, pix_per_block);
    return cudaGetLastError();
}

}  // namespace rst
