// 2-CTA (tcgen05 cta_group::2) variant of the halo GEMM for the 128->128 3x3 bottleneck convolutions.
//
// Why: with one CTA per tile the 288 KB weight matrix does not fit in shared memory and is re-streamed for every
// 128-pixel tile; the B ring is latency bound and tensor-pipe activity stalls near 50 % (profiles/r01_02).  Here a
// cluster of two CTAs works on two tiles at once with UMMA_M = 256: each CTA supplies its own 128-row A halo patch and
// HALF of B (64 of the 128 output channels: 18 blocks x 8 KB = 144 KB), which therefore stays RESIDENT for the whole
// kernel.  Per MMA each SM reads 4 KB of A + 2 KB of B from shared memory for 64 cycles of tensor work.
//
// Protocol (leader = even CTA of the pair):
//   * both CTAs TMA their operands with .cta_group::2, completing transactions on the LEADER's full barriers;
//   * the leader's elected thread issues tcgen05.mma.cta_group::2 and multicasts tcgen05.commit to the empty /
//     accumulator-full barriers of BOTH CTAs;
//   * both CTAs' epilogue warps drain their own TMEM half and arrive on the leader's accumulator-empty barrier.
#include "halo_gemm.cuh"

namespace rst {

using namespace umma;

__device__ __forceinline__ float warp_transpose_reduce2(float (&v)[32], int lane) {
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

constexpr int kN2 = 128;                 // output channels
constexpr int kRowB2 = 128;              // bytes per A row (64 channels)
constexpr int kKS2 = sched_ksteps(SCH_C3, kRowB2);          // 36 K-steps per 64-channel group
constexpr int kBHalf = (kN2 / 2) * 128;  // bytes of one B block half (4 K-steps x 64 rows)
constexpr int kHalo2 = 10 * 18 * kRowB2;
constexpr int kAStage2 = (kHalo2 + 1023) & ~1023;
constexpr int kAStages2 = 3;
constexpr int kTail2 = 6400;
constexpr int kMaxBBlocks2 = 18;          // 2 channel groups x 9 taps
constexpr int kPrefetchPairs = 3;         // L2 prefetch distance of the A producer, in tile pairs

constexpr int kThreads2 = 96 + 128 * 4;  // 3 control warps + 16 epilogue warps (one per lane quadrant and 32-column chunk)

// FUSE != 0 (HaloGemmParams::fuse): the TMA A producer is replaced by loader warps that read the RAW output of the previous
// convolution from global memory (L2), apply its conditional instance norm in registers and write the swizzled halo stage
// themselves -- the bytes TMA would have written, so the shared-memory port sees no extra traffic (a transform that rewrites
// a TMA-landed stage in place was measured in round 1: it read and wrote every stage once more and stalled the tensor pipe).
// Warps: 0 = B producer + L2 prefetch of later halos, 1 = MMA issuer, 2 .. 2+2*NL-1 = loaders (NL warps per 64-channel
// group; a thread owns 8 fixed channels, so its 16 coefficients live in registers), then the 16 epilogue warps.
constexpr int threads2f(int nl) { return 64 + 2 * nl * 32 + 128 * 4; }
#ifndef RST_TRUNK_REG_BOUND
#define RST_TRUNK_REG_BOUND kThreads2
#endif
// The launch bound the compiler sizes the register budget of the unfused kernel for (65536 / bound, rounded down to 8): the
// kernel is launched with kThreads2 threads whatever this is.  768 -> 80 registers, which leaves 16 K registers of the SM free.
constexpr int kTrunkRegBound = RST_TRUNK_REG_BOUND;

__device__ __forceinline__ void mbar_arrive_leader_release(uint64_t* bar) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & kPeerBitMask) : "memory");
}
__device__ __forceinline__ void mbar_wait_acquire_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    for (;; __nanosleep(40)) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (ok) break;
    }
}
// Waiting warps that poll hot steal issue slots from the loader warps (FUSE): back off between polls.
template <int NS>
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) __nanosleep(NS);
}
__device__ __forceinline__ uint4 ld_global_nc_16(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_shared_16(uint32_t addr, uint4 v) {
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

template <int FUSE, int NL>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(FUSE ? threads2f(NL) : kTrunkRegBound, 1)
halo_gemm2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const HaloGemmParams p) {
    constexpr int W_APROD = FUSE ? -1 : 0, W_BPROD = FUSE ? 0 : 1, W_MMA = FUSE ? 1 : 2;
    (void)W_APROD;
    constexpr int W_LOAD0 = 2, W_EPI0 = FUSE ? 2 + 2 * NL : 3;
    constexpr int N = kN2, CW = 32, NCH = N / CW, ESPLIT = 4;
    constexpr uint32_t TMEM_COLS = 2 * N;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int nblocks = p.n_groups * kKS2 / 4;               // 18 for 128 input channels
    uint8_t* sA = smem;
    uint8_t* sB = sA + kAStages2 * kAStage2;
    uint8_t* tail = sB + nblocks * kBHalf;
    uint64_t* a_full = reinterpret_cast<uint64_t*>(tail);
    uint64_t* a_empty = a_full + 4;
    uint64_t* b_full = a_empty + 4;                          // one per weight block (tap x 64 input channels): kMaxBBlocks2
    uint64_t* acc_full = b_full + kMaxBBlocks2;
    uint64_t* acc_empty = acc_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
    float* bias_s = reinterpret_cast<float*>(tail + 256);
    float* stat_s = bias_s + N;                              // [4 warps][2][N]

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
        // FUSE: every loader warp of both CTAs that fills a stage arrives once on the leader's full barrier
        for (int i = 0; i < 4; ++i) { mbar_init(&a_full[i], FUSE ? 2 * NL : 1); mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < kMaxBBlocks2; ++i) mbar_init(&b_full[i], 1);
        for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 2 * 4 * ESPLIT); }
        fence_barrier_init();
    }
    if (warp == 0 && lane == 0) { prefetch_tmap(&tmA); prefetch_tmap(&tmB); }
    for (int i = threadIdx.x; i < N; i += blockDim.x) bias_s[i] = p.bias ? p.bias[i] : 0.f;
    for (int i = threadIdx.x; i < 8 * N; i += blockDim.x) stat_s[i] = 0.f;
    __syncthreads();
    cluster_sync();                                           // barriers of both CTAs are initialised
    if (warp == W_MMA) tmem_alloc_2sm(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // the weight producer only reads constants: it runs ahead of the previous kernel's completion; everyone else waits here
    if (warp != W_BPROD) pdl_wait();

    auto tile_coords = [&](int t, int& n, int& h0, int& w0) {
        if (t >= total_tiles) { n = p.B; h0 = 0; w0 = 0; return; }     // phantom tile: out of bounds everywhere -> zeros
        n = t / tiles_per_img;
        const int r = t - n * tiles_per_img;
        h0 = (r / p.tiles_w) * 8; w0 = (r % p.tiles_w) * 16;
    };

    if (warp == W_APROD) {
        // ================= A producer (both CTAs): own halo patch, transactions land on the leader's barrier ====
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            const uint64_t pol_first = l2_policy_evict_first();
            for (int pr = pair_begin; pr < pair_end; ++pr) {
                int n, h0, w0;
                if (pr + kPrefetchPairs < pair_end) {            // pull a later tile's halo into L2 while this one is consumed
                    tile_coords(2 * (pr + kPrefetchPairs) + (int)rank, n, h0, w0);
                    for (int g = 0; g < p.n_groups; ++g) tma_prefetch_4d(&tmA, g * 64, h0 - 1, w0 - 1, n);
                }
                tile_coords(2 * pr + (int)rank, n, h0, w0);
                for (int g = 0; g < p.n_groups; ++g) {
                    mbar_wait(&a_empty[stage], phase ^ 1);
                    if (leader_cta) mbar_expect_tx(&a_full[stage], 2 * kHalo2);
                    if (p.l2_hints & 1) tma_load_4d_2sm_hint(sA + stage * kAStage2, &tmA, &a_full[stage], g * 64, h0 - 1, w0 - 1, n, pol_first);
                    else tma_load_4d_2sm(sA + stage * kAStage2, &tmA, &a_full[stage], g * 64, h0 - 1, w0 - 1, n);
                    if (++stage == kAStages2) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == W_BPROD) {
        // ================= B producer (both CTAs): this CTA's 64 output channels of every block, once ==========
        if (lane == 0 && pair_begin < pair_end) {
            const int per_group = kKS2 / 4;
            for (int g = 0; g < p.n_groups; ++g) {
                // one barrier per block (8 KB per CTA), in the order the MMAs consume them: the first tile pair's MMAs start
                // as soon as the first tap has landed and the other 136 KB stream in behind them
                for (int kb = g * per_group; kb < (g + 1) * per_group; ++kb) {
                    if (leader_cta) mbar_expect_tx(&b_full[kb], 2 * kBHalf);
                    tma_load_2d_2sm(sB + kb * kBHalf, &tmB, &b_full[kb], 0, kb * N + (int)rank * (N / 2));
                }
            }
        }
    } else if (FUSE && warp >= W_LOAD0 && warp < W_EPI0) {
        // ================= fused loaders (both CTAs): global -> registers -> norm affine -> swizzled shared memory ===
        const int lw = warp - W_LOAD0;
        const int g = lw / NL;                                    // the 64-channel group this warp serves
        const int lt = (lw % NL) * 32 + lane;                     // index among the group's loader threads
        const int v = lt & 7, slot = lt >> 3;                     // 16-byte chunk of the 128-byte row; pixel slot per iteration
        constexpr int SLOTS = NL * 4, ITERS = (180 + SLOTS - 1) / SLOTS;
        constexpr int BATCH = 3, NBATCH = (ITERS + BATCH - 1) / BATCH;      // register double buffering: 2 x BATCH vectors in flight
        static_assert(SLOTS == 8 || SLOTS == 4, "4 or 8 pixel slots per iteration");
        const int C = p.n_groups * 64, c0 = g * 64 + v * 8;
        const int rowC = p.WRU * C;                                // elements per image row
        // halo pixel hp = slot + 8 i sits at shared-memory row hp; its 16-byte chunk v is swizzled by (row & 7) = slot
        const uint32_t st_thread = smem_u32(sA) + slot * kRowB2;
        float ca[8], cb[8];
        int cur_n = -1;
        uint32_t seq = (uint32_t)g;                               // stages are consumed in (pair, group) order
        if (g < p.n_groups) {
            for (int pr = pair_begin; pr < pair_end; ++pr, seq += p.n_groups) {
                int n, h0, w0;
                if (lt == 0 && pr + kPrefetchPairs < pair_end) {   // keep the global reads L2 hits: TMA prefetch of a later halo
                    tile_coords(2 * (pr + kPrefetchPairs) + (int)rank, n, h0, w0);
                    tma_prefetch_4d(&tmA, g * 64, h0 - 1, w0 - 1, n);
                }
                tile_coords(2 * pr + (int)rank, n, h0, w0);
                if (n != cur_n && n < p.B) {
                    // the arithmetic of cin_apply_fast_kernel (halo_gemm.cu), so that fused and unfused results are identical
                    const double inv_p = 1.0 / ((double)p.H * (double)p.WRU);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const double2 st = *reinterpret_cast<const double2*>(p.fin_stats + ((long long)n * C + c0 + j) * 2);
                        const double mean = st.x * inv_p;
                        double var = fma(st.y, inv_p, -mean * mean);
                        if (var < 0.0) var = 0.0;
                        const float inv = rsqrtf((float)var + p.fin_eps), nmi = -(float)mean * inv;
                        const float* ps = p.fin_params + n * p.fin_param_bstride;
                        ca[j] = inv * ps[p.fin_scale_off + c0 + j];
                        cb[j] = ps[p.fin_bias_off + c0 + j] + nmi * ps[p.fin_scale_off + c0 + j];
                    }
                    cur_n = n;
                }
                const uint32_t stage = seq % kAStages2, phase = (seq / kAStages2) & 1;
                // A ROLLED, register double-buffered loop (the fully unrolled form was 6000 instructions and thrashed the
                // instruction cache).  Issue side walks the halo: iteration i handles row hp = slot + SLOTS*i = column ww, row hh.
                const int hy = h0 - 1, wx = w0 - 1;
                int hp_i = slot, hh = slot, ww = 0;                // slot < 8 < 10
                int off = (n * p.H + (hy + hh)) * rowC + wx * C + c0;                 // element offset (tensor < 2^31 elements)
                const bool n_ok = n < p.B;
                int hp_c = slot;
                uint32_t st_addr = st_thread + stage * kAStage2;
                auto issue = [&](uint4 (&x)[BATCH], uint4 (&sk)[FUSE == 2 ? BATCH : 1], uint32_t& okmask) {
                    okmask = 0;
#pragma unroll
                    for (int u = 0; u < BATCH; ++u) {
                        const bool ok = n_ok && hp_i < 180 && (unsigned)(hy + hh) < (unsigned)p.H && (unsigned)(wx + ww) < (unsigned)p.WRU;
                        okmask |= (ok ? 1u : 0u) << u;
                        x[u] = make_uint4(0, 0, 0, 0);
#ifdef RST_EXPERIMENTS
                        if (!(p.fuse_dbg & 1))
#endif
                        if (ok) x[u] = ld_global_nc_16(p.fin_x + off);
                        if (FUSE == 2) {
                            sk[u] = make_uint4(0, 0, 0, 0);
                            if (ok && p.fin_skip) sk[u] = ld_global_nc_16(p.fin_skip + off);
                        }
                        hp_i += SLOTS; hh += SLOTS; off += SLOTS * rowC;               // next iteration: SLOTS rows down the halo column ...
                        if (hh >= 10) { hh -= 10; ww += 1; off += C - 10 * rowC; }     // ... wrapping into the next column
                    }
                };
                auto consume = [&](const uint4 (&x)[BATCH], const uint4 (&sk)[FUSE == 2 ? BATCH : 1], uint32_t okmask) {
#pragma unroll
                    for (int u = 0; u < BATCH; ++u) {
                        if (hp_c < 180) {                                                       // rows 180.. do not exist
                            uint4 o = make_uint4(0, 0, 0, 0);                                   // 'same' padding stays zero
                            if ((okmask >> u) & 1u) {
                                const uint32_t* xw = reinterpret_cast<const uint32_t*>(&x[u]);
                                const __nv_bfloat162* sb = reinterpret_cast<const __nv_bfloat162*>(&sk[FUSE == 2 ? u : 0]);
                                uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    // bf16 -> fp32 is a 16-bit shift: low half << 16, high half masked
                                    const float x0 = __uint_as_float(xw[j] << 16), x1 = __uint_as_float(xw[j] & 0xFFFF0000u);
                                    const float o0 = fmaf(x0, ca[2 * j], cb[2 * j]), o1 = fmaf(x1, ca[2 * j + 1], cb[2 * j + 1]);
                                    if (FUSE == 1) {
                                        asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(ow[j]) : "f"(o1), "f"(o0));    // relu + pack in one
                                    } else {
                                        const float2 rv = __bfloat1622float2(sb[j]);
                                        __nv_bfloat162 h2 = __floats2bfloat162_rn(o0 + rv.x, o1 + rv.y);
                                        ow[j] = *reinterpret_cast<uint32_t*>(&h2);
                                    }
                                }
                                if (FUSE == 2) {
                                    const int w2 = hp_c / 10, h2 = hp_c - w2 * 10;
                                    if (h2 >= 1 && h2 <= 8 && w2 >= 1 && w2 <= 16) {            // centre pixels: this tile owns them
                                        const int o_off = (n * p.H + (hy + h2)) * rowC + (wx + w2) * C + c0;
                                        *reinterpret_cast<uint4*>(p.fin_out + o_off) = o;
                                    }
                                }
                            }
#ifdef RST_EXPERIMENTS
                            if (!(p.fuse_dbg & 2))
#endif
                            st_shared_16(st_addr + ((v ^ (hp_c & 7)) << 4), o);
                        }
                        hp_c += SLOTS; st_addr += SLOTS * kRowB2;
                    }
                };
                uint4 xa[BATCH], xb_[BATCH], ska[FUSE == 2 ? BATCH : 1], skb[FUSE == 2 ? BATCH : 1];
                uint32_t oka, okb;
                issue(xa, ska, oka);
                mbar_wait(&a_empty[stage], phase ^ 1);            // the first loads are in flight while we wait for the slot
#pragma unroll 1
                for (int k = 0; k < NBATCH; k += 2) {              // batches past the halo's end are predicated off
                    issue(xb_, skb, okb);
                    consume(xa, ska, oka);
                    issue(xa, ska, oka);
                    consume(xb_, skb, okb);
                }
                fence_proxy_async();                              // generic-proxy stores -> visible to the tensor core's async proxy
                __syncwarp();
                if (lane == 0) mbar_arrive_leader_release(&a_full[stage]);
            }
        }
    } else if (warp == W_MMA) {
        // ================= MMA issuer: leader CTA only ==========================================================
        if (leader_cta) {
            const uint32_t idesc = make_idesc_bf16(256, N);
            const uint64_t da_const = make_smem_desc(0, 16, 10 * kRowB2, SWIZZLE_128B);
            const uint64_t db_const = make_smem_desc(0, 16, 1024, SWIZZLE_128B);
            const bool issuer = elect_one();
            uint32_t as = 0, aph = 0, cs = 0, cph = 0;
            const uint32_t sB16 = __shfl_sync(0xffffffffu, smem_u32(sB) >> 4, 0);
            const uint32_t sA16 = __shfl_sync(0xffffffffu, smem_u32(sA) >> 4, 0);
            const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
            for (int pr = pair_begin; pr < pair_end; ++pr) {
                mbar_wait(&acc_empty[cs], cph ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_u + cs * N;
                for (int g = 0; g < p.n_groups; ++g) {
                    if (FUSE) mbar_wait_acquire_cluster(&a_full[as], aph);
                    else mbar_wait(&a_full[as], aph);
                    const bool first_pair = pr == pair_begin;          // weights are resident after the first tile pair
                    tc_fence_after();
                    const uint32_t a_base16 = sA16 + as * (kAStage2 >> 4);
                    const uint32_t b_base16 = sB16 + g * (kKS2 / 4) * (kBHalf >> 4);
#pragma unroll
                    for (int ks = 0; ks < kKS2; ++ks) {
                        if (first_pair && (ks & 3) == 0) { mbar_wait(&b_full[g * (kKS2 / 4) + ks / 4], 0); tc_fence_after(); }
                        const uint64_t da = da_const | (uint64_t)(a_base16 + (uint32_t)(sched_off(SCH_C3, kRowB2, ks) >> 4));
                        const uint64_t db = db_const | (uint64_t)(b_base16 + (ks / 4) * (kBHalf >> 4) + (ks & 3) * 2);
                        if (issuer) mma_f16_ss_2sm(tmem_d, da, db, idesc, ks == 0 ? (uint32_t)(g != 0) : 1u);
                    }
                    if (issuer) mma_commit_2sm(&a_empty[as], 3);
                    if (++as == kAStages2) { as = 0; aph ^= 1; }
                }
                if (issuer) mma_commit_2sm(&acc_full[cs], 3);
                if (++cs == 2) { cs = 0; cph ^= 1; }
            }
        }
    } else if (warp >= W_EPI0) {
        // ================= epilogue (both CTAs): own 128 rows; two warps per lane quadrant split the columns ======
        const int q = warp & 3;
        const int c_begin = ((warp - W_EPI0) >> 2) * (NCH / ESPLIT), c_end = c_begin + NCH / ESPLIT;
        const int row = q * 32 + lane;
        const int w_l = row >> 3, h_l = row & 7;
        float* my_sum = stat_s + q * 2 * N;
        float* my_sq = my_sum + N;
        const bool do_stats = p.stats != nullptr;
        const uint64_t pol_last = l2_policy_evict_last();
        int cur_n = -1;
        auto flush = [&]() {
            if (do_stats && cur_n >= 0 && cur_n < p.B) {
#pragma unroll 1
                for (int c = c_begin; c < c_end; ++c) {
                    const int col = c * CW + lane;
                    double* dst = p.stats + ((size_t)cur_n * p.stats_c + col) * 2;
                    atomicAdd(dst, (double)my_sum[col]);
                    atomicAdd(dst + 1, (double)my_sq[col]);
                    my_sum[col] = 0.f; my_sq[col] = 0.f;
                }
            }
        };
        uint32_t cs = 0, cph = 0;
        for (int pr = pair_begin; pr < pair_end; ++pr) {
            int n, h0, w0;
            tile_coords(2 * pr + (int)rank, n, h0, w0);
            const int gh = h0 + h_l, gw = w0 + w_l;
            const bool valid = n < p.B && gh < p.H && gw < p.WRU;
            if (n != cur_n) { flush(); cur_n = n; }
            if (FUSE) mbar_wait_backoff<200>(&acc_full[cs], cph);
            else mbar_wait(&acc_full[cs], cph);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + cs * N;
            // Each warp owns one 32-column chunk of one lane quadrant: the accumulator stage is handed back to the MMA warp
            // as soon as the values are in registers, then the math / stores / statistics run.
            auto process = [&](float (&v)[32], int c) {
                const float* bs_ = bias_s + c * CW;
                uint32_t packed[CW / 2];
#pragma unroll
                for (int j = 0; j < CW; j += 2) {
                    const float x0 = fmaxf(v[j] + bs_[j], 0.f), x1 = fmaxf(v[j + 1] + bs_[j + 1], 0.f);
                    __nv_bfloat162 h2 = __floats2bfloat162_rn(x0, x1);
                    packed[j >> 1] = *reinterpret_cast<uint32_t*>(&h2);
                    v[j] = valid ? __low2float(h2) : 0.f;
                    v[j + 1] = valid ? __high2float(h2) : 0.f;
                }
                if (valid) {
                    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.y) +
                                       (((size_t)n * p.out_H + gh) * p.out_W + gw) * p.out_C + c * CW;
#pragma unroll
                    for (int j = 0; j < CW / 16; ++j) {                                              // full sectors per lane
                        if (p.l2_hints & 2) st_global_v8_hint(o + 16 * j, packed + 8 * j, pol_last);
                        else st_global_v8(o + 16 * j, packed + 8 * j);
                    }
                }
                if (do_stats) {
                    float sq[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) sq[j] = v[j] * v[j];
                    const float s1 = warp_transpose_reduce2(v, lane);
                    const float s2 = warp_transpose_reduce2(sq, lane);
                    my_sum[c * CW + lane] += s1;
                    my_sq[c * CW + lane] += s2;
                }
            };
            float va[32];
            tmem_ld_32x32(taddr + c_begin * 32, va);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(&acc_empty[cs]);
            process(va, c_begin);
            if (++cs == 2) { cs = 0; cph ^= 1; }
        }
        flush();
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync();                                           // both CTAs are done with TMEM and with remote barriers
    if (warp == W_MMA) tmem_dealloc_2sm(tmem_base, TMEM_COLS);
}

size_t halo_gemm2_smem_bytes(int n_groups) {
    return (size_t)kAStages2 * kAStage2 + (size_t)(n_groups * kKS2 / 4) * kBHalf + kTail2 + 1024;
}

// Weights for the 2-CTA kernel use the same packed blocks; the tensor map's box is 64 rows (one CTA's half of N).
cudaError_t launch_halo_gemm2(const CUtensorMap& tmA, const CUtensorMap& tmB_half, const HaloGemmParams& p, int num_sms,
                              cudaStream_t s) {
    static SmemAttrCache configured[2];
    const size_t smem = halo_gemm2_smem_bytes(p.n_groups);
    const int total = p.B * p.tiles_h * p.tiles_w;
    if (total == 0) return cudaSuccess;
    const int pairs = (total + 1) / 2;
    int clusters = num_sms / 2;
    if (pairs < clusters) clusters = pairs;
    if (p.fuse == 0) {
        if (cudaError_t e = ensure_dynamic_smem(halo_gemm2_kernel<0, 1>, smem, configured[0])) return e;
        if (cudaError_t e = launch_pdl(p.pdl != 0, halo_gemm2_kernel<0, 1>, dim3(2 * clusters), dim3(kThreads2), smem, s, tmA, tmB_half, p)) return e;
        return cudaGetLastError();
    }
    // fuse 1 only is instantiated: fuse 2 (skip add + write-back in the loader) compiles but was never profitable to finish
    if (p.n_groups != 2 || !p.fin_x || !p.fin_stats || !p.fin_params || p.fuse != 1) return cudaErrorInvalidValue;
    if (cudaError_t e = ensure_dynamic_smem(halo_gemm2_kernel<1, 2>, smem, configured[1])) return e;
    if (cudaError_t e = launch_pdl(p.pdl != 0, halo_gemm2_kernel<1, 2>, dim3(2 * clusters), dim3(threads2f(2)), smem, s, tmA, tmB_half, p)) return e;
    return cudaGetLastError();
}

}  // namespace rst
