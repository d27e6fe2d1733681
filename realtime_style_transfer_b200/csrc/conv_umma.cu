// tcgen05 / TMA implicit-GEMM convolution kernels (see conv_umma.cuh for the design) and their bf16
// elementwise companions.
#include <mutex>

#include "conv_umma.cuh"

namespace rst {

using namespace umma;

// ------------------------------------------------------------------------------------------------
// warp butterfly: every lane holds v[0..31] (32 columns of its row); afterwards lane j holds the sum over
// the 32 lanes of column j in v[0].  31 shuffles instead of 160.
// ------------------------------------------------------------------------------------------------
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

template <int COUT>
__global__ void __launch_bounds__(kUmmaThreads, 1)
conv3x3_umma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const ConvUmmaParams p) {
    using S = ConvUmmaSmem<COUT>;
    constexpr int NB = S::kNumBStages;
    constexpr int CH = COUT / 32;
    constexpr uint32_t TMEM_COLS = 2 * COUT;
    static_assert(COUT % 32 == 0 && COUT >= 32 && COUT <= 256, "COUT must be a multiple of 32 in [32,256]");
    static_assert((TMEM_COLS & (TMEM_COLS - 1)) == 0 && TMEM_COLS <= 512, "TMEM columns must be a power of two");

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;
    uint8_t* sB = smem + kNumAStages * kAStageBytes;
    uint8_t* tail = sB + NB * S::kBStageBytes;
    uint64_t* a_full = reinterpret_cast<uint64_t*>(tail);
    uint64_t* a_empty = a_full + kNumAStages;
    uint64_t* b_full = a_empty + kNumAStages;
    uint64_t* b_empty = b_full + NB;
    uint64_t* acc_full = b_empty + NB;
    uint64_t* acc_empty = acc_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
    float* bias_s = reinterpret_cast<float*>(tail + 256);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_per_img = p.tiles_h * p.tiles_w;
    const int total_tiles = p.B * tiles_per_img;
    const int tpc = (total_tiles + gridDim.x - 1) / gridDim.x;
    const int tile_begin = blockIdx.x * tpc;
    const int tile_end = min(total_tiles, tile_begin + tpc);

    if (threadIdx.x == 0) {
        for (int i = 0; i < kNumAStages; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < NB; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4); }
        fence_barrier_init();
    }
    if (warp == 0 && lane == 0) { prefetch_tmap(&tmA); prefetch_tmap(&tmB); }
    if (warp == 2) tmem_alloc(tmem_slot, TMEM_COLS);
    for (int i = threadIdx.x; i < COUT; i += blockDim.x) bias_s[i] = p.bias ? p.bias[i] : 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================= A producer: one halo patch per (tile, channel half) =================
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (int t = tile_begin; t < tile_end; ++t) {
                const int n = t / tiles_per_img, r = t - n * tiles_per_img;
                const int h0 = (r / p.tiles_w) * kUmmaTH, w0 = (r % p.tiles_w) * kUmmaTW;
                for (int half = 0; half < p.nhalf; ++half) {
                    mbar_wait(&a_empty[stage], phase ^ 1);
                    mbar_expect_tx(&a_full[stage], kHaloBytes);
                    tma_load_4d(sA + stage * kAStageBytes, &tmA, &a_full[stage], half * 64, h0 - 1, w0 - 1, n);
                    if (++stage == kNumAStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================= B producer: weight k-blocks (half, tap), COUT x 64 each ==============
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (int t = tile_begin; t < tile_end; ++t) {
                for (int kb = 0; kb < p.nhalf * 9; ++kb) {
                    mbar_wait(&b_empty[stage], phase ^ 1);
                    mbar_expect_tx(&b_full[stage], S::kBStageBytes);
                    tma_load_2d(sB + stage * S::kBStageBytes, &tmB, &b_full[stage], 0, kb * COUT);
                    if (++stage == NB) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 2) {
        // ================= MMA issuer (single thread) ===========================================
        if (lane == 0) {
            const uint32_t idesc = make_idesc_bf16(128, COUT);
            uint32_t as = 0, aph = 0, bs = 0, bph = 0, cs = 0, cph = 0;
            for (int t = tile_begin; t < tile_end; ++t) {
                mbar_wait(&acc_empty[cs], cph ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + cs * COUT;
                for (int half = 0; half < p.nhalf; ++half) {
                    mbar_wait(&a_full[as], aph);
                    tc_fence_after();
                    const uint32_t a_base = smem_u32(sA + as * kAStageBytes);
#pragma unroll 1
                    for (int tap = 0; tap < 9; ++tap) {
                        const int dy = tap / 3, dx = tap - dy * 3;
                        mbar_wait(&b_full[bs], bph);
                        tc_fence_after();
                        const uint32_t a_tap = a_base + (dx * kHaloH + dy) * 128;
                        const uint32_t b_base = smem_u32(sB + bs * S::kBStageBytes);
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint64_t da = make_smem_desc(a_tap + k * 32, 16, kHaloH * 128, SWIZZLE_128B);
                            const uint64_t db = make_smem_desc(b_base + k * 32, 16, 1024, SWIZZLE_128B);
                            mma_f16_ss(tmem_d, da, db, idesc, (uint32_t)((half | tap | k) != 0));
                        }
                        mma_commit(&b_empty[bs]);
                        if (++bs == NB) { bs = 0; bph ^= 1; }
                    }
                    mma_commit(&a_empty[as]);
                    if (++as == kNumAStages) { as = 0; aph ^= 1; }
                }
                mma_commit(&acc_full[cs]);
                if (++cs == 2) { cs = 0; cph ^= 1; }
            }
        }
    } else {
        // ================= epilogue: TMEM -> bias/ReLU/bf16/stats -> global ======================
        const int q = warp & 3;                  // TMEM lane quadrant this warp may access
        const int row = q * 32 + lane;
        const int w_l = row >> 3, h_l = row & 7;
        float acc_sum[CH], acc_sq[CH];
#pragma unroll
        for (int c = 0; c < CH; ++c) { acc_sum[c] = 0.f; acc_sq[c] = 0.f; }
        int cur_n = -1;
        auto flush = [&]() {
            if (p.stats != nullptr && cur_n >= 0) {
#pragma unroll
                for (int c = 0; c < CH; ++c) {
                    double* dst = p.stats + ((size_t)cur_n * COUT + c * 32 + lane) * 2;
                    atomicAdd(dst, (double)acc_sum[c]);
                    atomicAdd(dst + 1, (double)acc_sq[c]);
                    acc_sum[c] = 0.f; acc_sq[c] = 0.f;
                }
            }
        };
        uint32_t cs = 0, cph = 0;
        for (int t = tile_begin; t < tile_end; ++t) {
            const int n = t / tiles_per_img, r = t - n * tiles_per_img;
            const int gh = (r / p.tiles_w) * kUmmaTH + h_l, gw = (r % p.tiles_w) * kUmmaTW + w_l;
            const bool valid = gh < p.H && gw < p.W;
            if (n != cur_n) { flush(); cur_n = n; }
            __nv_bfloat16* out = p.y + (((size_t)n * p.H + gh) * p.W + gw) * COUT;
            mbar_wait(&acc_full[cs], cph);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + cs * COUT;
#pragma unroll
            for (int c = 0; c < CH; ++c) {
                float v[32];
                tmem_ld_32x32(taddr + c * 32, v);
                tmem_ld_wait();
                uint32_t packed[16];
#pragma unroll
                for (int j = 0; j < 32; j += 2) {
                    float a = v[j] + bias_s[c * 32 + j], b = v[j + 1] + bias_s[c * 32 + j + 1];
                    if (p.relu) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); }
                    __nv_bfloat162 h2 = __floats2bfloat162_rn(a, b);
                    packed[j >> 1] = *reinterpret_cast<uint32_t*>(&h2);
                    // statistics of the values as stored (bf16-rounded); rows outside the image count as 0
                    v[j] = valid ? __low2float(h2) : 0.f;
                    v[j + 1] = valid ? __high2float(h2) : 0.f;
                }
                if (valid) {
                    uint4* o4 = reinterpret_cast<uint4*>(out + c * 32);
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        o4[j] = make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
                }
                if (p.stats != nullptr) {
                    float sq[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) sq[j] = v[j] * v[j];
                    acc_sum[c] += warp_transpose_reduce(v, lane);
                    acc_sq[c] += warp_transpose_reduce(sq, lane);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[cs]);
            if (++cs == 2) { cs = 0; cph ^= 1; }
        }
        flush();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, TMEM_COLS);
}

template <int COUT>
static cudaError_t launch_t(const CUtensorMap& tmA, const CUtensorMap& tmB, const ConvUmmaParams& p, int num_sms,
                            cudaStream_t s) {
    static bool configured = false;
    constexpr int smem = ConvUmmaSmem<COUT>::kBytes;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(conv3x3_umma_kernel<COUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    const int total = p.B * p.tiles_h * p.tiles_w;
    if (total == 0) return cudaSuccess;
    const int grid = total < num_sms ? total : num_sms;
    conv3x3_umma_kernel<COUT><<<grid, kUmmaThreads, smem, s>>>(tmA, tmB, p);
    return cudaGetLastError();
}

cudaError_t launch_conv3x3_umma(int cout, const CUtensorMap& tmA, const CUtensorMap& tmB, const ConvUmmaParams& p,
                                int num_sms, cudaStream_t s) {
    switch (cout) {
        case 128: return launch_t<128>(tmA, tmB, p, num_sms, s);
        case 64: return launch_t<64>(tmA, tmB, p, num_sms, s);
        case 32: return launch_t<32>(tmA, tmB, p, num_sms, s);
        default: return cudaErrorInvalidValue;
    }
}

// ------------------------------------------------------------------------------------------------
// tensor maps
// ------------------------------------------------------------------------------------------------
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

// NHWC bf16 activation, dims ordered (C, H, W, N) so that a box enumerates h fastest among pixels:
// box = (64 channels, 10 rows, 18 columns, 1 sample), SWIZZLE_128B.
bool umma_encode_activation_map(CUtensorMap* out, const void* base, int B, int H, int W, int C, std::string* err) {
    if (!umma_init(err)) return false;
    cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)H, (cuuint64_t)W, (cuuint64_t)B};
    cuuint64_t gstr[3] = {(cuuint64_t)W * C * 2, (cuuint64_t)C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)kHaloH, (cuuint32_t)kHaloW, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gdim, gstr, box, es,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        if (err) *err = "cuTensorMapEncodeTiled(activation) failed with code " + std::to_string((int)r);
        return false;
    }
    return true;
}

// packed weights: rows x 64 bf16 (K-major), box = (64, box_rows), SWIZZLE_128B.
bool umma_encode_weight_map(CUtensorMap* out, const void* base, int rows, int box_rows, std::string* err) {
    if (!umma_init(err)) return false;
    cuuint64_t gdim[2] = {64, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {128};
    cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
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

__device__ __forceinline__ float act_f(float v, int act) {
    if (act == ACT_RELU) return fmaxf(v, 0.f);
    if (act == ACT_SIGMOID) return 1.f / (1.f + __expf(-v));
    return v;
}

// 8 channels (16 bytes) per thread-iteration; coefficients for the block's sample staged in shared memory.
__global__ void __launch_bounds__(256) cin_apply_bf16_kernel(const CinApplyBf16 p, int pix_per_block) {
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
    const int vec_per_pix = C >> 3;
    const long long base = (long long)n * p.P * vec_per_pix;
    const long long v0 = (long long)blockIdx.x * pix_per_block * vec_per_pix;
    const long long v1 = min((long long)p.P * vec_per_pix, v0 + (long long)pix_per_block * vec_per_pix);
    const bool blend = p.num_styles == 2 && p.weights != nullptr;
    const uint4* x4 = reinterpret_cast<const uint4*>(p.x) + base;
    const uint4* r4 = p.residual ? reinterpret_cast<const uint4*>(p.residual) + base : nullptr;
    uint4* y4 = reinterpret_cast<uint4*>(p.y) + base;
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
        uint4 outv;
        __nv_bfloat162* ob = reinterpret_cast<__nv_bfloat162*>(&outv);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float2 xv = __bfloat1622float2(xb[j]);
            float o[2];
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
                o[h] = act_f(bias + xh * scale, p.act);
            }
            if (r4) {
                float2 rv = __bfloat1622float2(rb[j]);
                o[0] += rv.x; o[1] += rv.y;
            }
            ob[j] = __floats2bfloat162_rn(o[0], o[1]);
        }
        y4[v] = outv;
    }
}

cudaError_t launch_cin_apply_bf16(const CinApplyBf16& p, cudaStream_t s) {
    if (p.B == 0 || p.P == 0) return cudaSuccess;
    if (p.C % 8 != 0) return cudaErrorInvalidValue;
    int pix_per_block = max(1, 32768 / p.C);
    dim3 grid((unsigned)ceil_div(p.P, pix_per_block), (unsigned)p.B);
    size_t smem = (size_t)(2 + 2 * p.num_styles) * p.C * sizeof(float);
    cin_apply_bf16_kernel<<<grid, 256, smem, s>>>(p, pix_per_block);
    return cudaGetLastError();
}

}  // namespace rst
