// Generic fp32 kernels (CUDA cores): direct convolution as an implicit GEMM, depthwise convolution,
// instance-norm moments / apply, pooling, style-weight blend, Gram matrix.
// These serve the RST_PRECISION_FP32 path (bar: 1e-4 abs vs the oracle), the style predictor and
// every layer the tensor-core path does not cover yet.  NHWC throughout.
#include "rst_internal.cuh"

#include <cstdlib>

namespace rst {

__device__ __forceinline__ float apply_act(float v, int act) {
    switch (act) {
        case ACT_RELU: return fmaxf(v, 0.f);
        case ACT_HSWISH: return v * fminf(fmaxf(v + 3.f, 0.f), 6.f) * (1.f / 6.f);
        case ACT_HSIGMOID: return fminf(fmaxf(v + 3.f, 0.f), 6.f) * (1.f / 6.f);
        case ACT_SIGMOID: return 1.f / (1.f + expf(-v));
        default: return v;
    }
}

// ---------------------------------------------------------------------------------------------
// conv2d / conv2d_transpose, padding='same' semantics supplied by the caller as pad_t/pad_l.
// GEMM view: M = B*Ho*Wo output pixels, N = Co, K = kh*kw*Ci.  BMxBN tile per CTA, BK = 16.
// ---------------------------------------------------------------------------------------------
// PHASE (stride-2 input-gradient form only): one launch per output parity (ph_y, ph_x).  Only the filter taps whose parity
// meets the stride are enumerated (1, 2, 2 or 4 of the 9 taps of a 3x3 kernel) instead of multiplying the others by zero.
template <int BM, int BN, int TM, int TN, bool PHASE>
__global__ void __launch_bounds__((BM / TM) * (BN / TN)) conv_f32_kernel(const ConvF32 p) {
    constexpr int BK = 16;
    constexpr int NT = (BM / TM) * (BN / TN);
    constexpr int A_ROWS = NT / BK;          // pixel rows gathered per pass
    constexpr int A_PASSES = BM / A_ROWS;
    constexpr int B_ROWS = NT / BN;          // k rows loaded per pass
    constexpr int B_PASSES = (BK + B_ROWS - 1) / B_ROWS;
    static_assert(NT % BK == 0 && BM % A_ROWS == 0 && NT % BN == 0, "tile/loader mismatch");

    __shared__ float As[BK][BM + 4];
    __shared__ float Bs[BK][BN + 4];

    const int tid = threadIdx.x;
    // PHASE: the pixel grid of this launch is the (Hp x Wp) sub-lattice oy = 2 i + ph_y, ox = 2 j + ph_x
    const int Hp = PHASE ? (p.Ho - p.ph_y + 1) / 2 : p.Ho, Wp = PHASE ? (p.Wo - p.ph_x + 1) / 2 : p.Wo;
    const int ky0 = PHASE ? (p.ph_y + p.pad_t) & 1 : 0, kx0 = PHASE ? (p.ph_x + p.pad_l) & 1 : 0;
    const int nky = PHASE ? (p.kh - ky0 + 1) / 2 : p.kh, nkx = PHASE ? (p.kw - kx0 + 1) / 2 : p.kw;
    const long long M = (long long)p.B * Hp * Wp;
    const int K = nky * nkx * p.Ci;
    const long long m0 = (long long)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;

    const int a_k = tid % BK;
    const int a_r = tid / BK;
    int a_n[A_PASSES], a_oy[A_PASSES], a_ox[A_PASSES];
#pragma unroll
    for (int j = 0; j < A_PASSES; ++j) {
        long long m = m0 + a_r + j * A_ROWS;
        if (m < M) {
            int ox = (int)(m % Wp);
            long long t = m / Wp;
            a_ox[j] = PHASE ? 2 * ox + p.ph_x : ox;
            a_oy[j] = PHASE ? 2 * (int)(t % Hp) + p.ph_y : (int)(t % Hp);
            a_n[j] = (int)(t / Hp);
        } else {
            a_n[j] = -1; a_oy[j] = 0; a_ox[j] = 0;
        }
    }
    const int b_c = tid % BN;
    const int b_r = tid / BN;

    const int tx = tid % (BN / TN);
    const int ty = tid / (BN / TN);
    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < K; k0 += BK) {
        // ---- gather A (im2col on the fly) ----
        {
            const int kk = k0 + a_k;
            int ky = 0, kx = 0, ci = 0;
            const bool kvalid = kk < K;
            if (kvalid) {
                int tap = kk / p.Ci;
                ci = kk - tap * p.Ci;
                ky = tap / nkx;
                kx = tap - ky * nkx;
                if (PHASE) { ky = ky0 + 2 * ky; kx = kx0 + 2 * kx; }
            }
#pragma unroll
            for (int j = 0; j < A_PASSES; ++j) {
                float v = 0.f;
                if (kvalid && a_n[j] >= 0) {
                    int iy, ix;
                    bool ok;
                    if (!p.transposed) {
                        iy = a_oy[j] * p.stride - p.pad_t + ky;
                        ix = a_ox[j] * p.stride - p.pad_l + kx;
                        ok = iy >= 0 && iy < p.Hi && ix >= 0 && ix < p.Wi;
                    } else {
                        int ty_ = a_oy[j] + p.pad_t - ky;
                        int tx_ = a_ox[j] + p.pad_l - kx;
                        ok = ty_ >= 0 && tx_ >= 0 && (ty_ % p.stride) == 0 && (tx_ % p.stride) == 0;
                        iy = ty_ / p.stride;
                        ix = tx_ / p.stride;
                        ok = ok && iy < p.Hi && ix < p.Wi;
                    }
                    if (ok) {
                        v = __ldg(p.x + (((long long)a_n[j] * p.Hi + iy) * p.Wi + ix) * p.Ci + ci);
                        v = v * p.in_scale + p.in_shift;
                    }
                }
                As[a_k][a_r + j * A_ROWS] = v;
            }
        }
        // ---- load B (weights) ----
#pragma unroll
        for (int j = 0; j < B_PASSES; ++j) {
            int r = b_r + j * B_ROWS;
            if (r < BK) {
                int kk = k0 + r;
                int co = n0 + b_c;
                float v = 0.f;
                if (kk < K && co < p.Co) {
                    int tap = kk / p.Ci;
                    int ci = kk - tap * p.Ci;
                    if (PHASE) { const int ty_ = tap / nkx; tap = (ky0 + 2 * ty_) * p.kw + kx0 + 2 * (tap - ty_ * nkx); }
                    v = __ldg(p.w + tap * p.w_tap + ci * p.w_ci + co * p.w_co);
                }
                Bs[r][b_c] = v;
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[TM], b[TN];
#pragma unroll
            for (int i = 0; i < TM; ++i) a[i] = As[k][ty * TM + i];
#pragma unroll
            for (int j = 0; j < TN; ++j) b[j] = Bs[k][tx * TN + j];
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }

#pragma unroll
    for (int i = 0; i < TM; ++i) {
        long long m = m0 + ty * TM + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            int co = n0 + tx * TN + j;
            if (co >= p.Co) continue;
            float v = acc[i][j];
            if (p.bias) v += __ldg(p.bias + co);
            v = apply_act(v, p.act1);
            if (p.post_scale) v = v * __ldg(p.post_scale + co) + __ldg(p.post_shift + co);
            v = apply_act(v, p.act2);
            long long o = m * p.Co + co;
            if (PHASE) {
                const int ox = (int)(m % Wp);
                const long long t = m / Wp;
                o = ((((long long)(t / Hp)) * p.Ho + 2 * (int)(t % Hp) + p.ph_y) * p.Wo + 2 * ox + p.ph_x) * p.Co + co;
            }
            if (p.residual) v += __ldg(p.residual + o);
            if (p.out_tf32) { uint32_t r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v)); v = __uint_as_float(r); }
            p.y[o] = v;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Direct stride-1 convolution for thin outputs (Co <= 32) on large images: the 9x9 stem, the 16 -> 3 head (forward and
// input gradient) and the loss model's 64 -> 3 input gradient.  The implicit-GEMM kernel above wastes its 16- / 64-wide
// column tile on them and pays an im2col index computation per gathered element.
// A CTA owns a 32x32 output tile; per chunk of CK input channels it stages the (32+K-1)^2 halo channel-planar in shared
// memory (lanes read consecutive pixels: conflict-free) plus the chunk's weights (read as warp-wide broadcasts).
// A thread owns a column of 4 output pixels x CO channels: for one (column tap, channel) it loads the 4+K-1 inputs of its
// column once and slides the K row taps over them in registers: (4+K-1) + K*CO/4 loads per 4*K*CO FMAs.
// ---------------------------------------------------------------------------------------------
template <int CO, int KS>
__global__ void __launch_bounds__(256) conv_direct_f32_kernel(const ConvF32 p, int CK, int tiles_x) {
    constexpr int T = 32, PX = 4, HH = T + KS - 1, HWP = HH | 1, NX = PX + KS - 1;
    extern __shared__ float direct_smem[];
    const int plane = HH * HWP + ((HH * HWP) % 2 == 0 ? 1 : 0);              // odd plane stride: the chunk loader writes ck-fastest
    float* xs = direct_smem;                                                   // [CK][HH][HWP]
    float* ws = direct_smem + ((CK * plane + 3) & ~3);                         // [KS*KS][CK][CO], 16-byte aligned
    const int tid = threadIdx.x, tx = tid & 31, tg = tid >> 5;
    const int n = blockIdx.y, co0 = blockIdx.z * CO;                           // wider layers: CO output channels per CTA
    const int x0 = (blockIdx.x % tiles_x) * T, y0 = (blockIdx.x / tiles_x) * T;
    // halo origin in input coordinates and the tap order along the halo (the input-gradient form walks the taps backwards)
    const int hy0 = p.transposed ? y0 + p.pad_t - (KS - 1) : y0 - p.pad_t;
    const int hx0 = p.transposed ? x0 + p.pad_l - (KS - 1) : x0 - p.pad_l;
    float acc[PX][CO];
#pragma unroll
    for (int j = 0; j < PX; ++j)
#pragma unroll
        for (int c = 0; c < CO; ++c) acc[j][c] = 0.f;
    const float* xn = p.x + (long long)n * p.Hi * p.Wi * p.Ci;
    for (int c0 = 0; c0 < p.Ci; c0 += CK) {
        __syncthreads();
        for (int e = tid; e < CK * HH * HH; e += 256) {
            const int ck = e % CK, pix = e / CK, hx = pix % HH, hy = pix / HH;
            const int iy = hy0 + hy, ix = hx0 + hx;
            float v = 0.f;
            if (c0 + ck < p.Ci && iy >= 0 && iy < p.Hi && ix >= 0 && ix < p.Wi)
                v = fmaf(__ldg(xn + ((long long)iy * p.Wi + ix) * p.Ci + c0 + ck), p.in_scale, p.in_shift);
            xs[ck * plane + hy * HWP + hx] = v;
        }
        for (int e = tid; e < KS * KS * CK * CO; e += 256) {
            const int co = e % CO, ck = (e / CO) % CK, t = e / (CO * CK);
            const int dy = t / KS, dx = t % KS;
            const int tap = p.transposed ? (KS - 1 - dy) * KS + (KS - 1 - dx) : t;
            ws[e] = (co0 + co < p.Co && c0 + ck < p.Ci) ? __ldg(p.w + tap * p.w_tap + (c0 + ck) * p.w_ci + (co0 + co) * p.w_co) : 0.f;
        }
        __syncthreads();
        for (int dx = 0; dx < KS; ++dx) {
            for (int ck = 0; ck < CK; ++ck) {
                float xv[NX];
                const float* xc = xs + ck * plane + (tg * PX) * HWP + tx + dx;
#pragma unroll
                for (int j = 0; j < NX; ++j) xv[j] = xc[j * HWP];
#pragma unroll
                for (int dy = 0; dy < KS; ++dy) {
                    const float4* w4 = reinterpret_cast<const float4*>(ws + ((dy * KS + dx) * CK + ck) * CO);
#pragma unroll
                    for (int c = 0; c < CO / 4; ++c) {
                        const float4 w = w4[c];
#pragma unroll
                        for (int j = 0; j < PX; ++j) {
                            acc[j][4 * c + 0] = fmaf(xv[j + dy], w.x, acc[j][4 * c + 0]);
                            acc[j][4 * c + 1] = fmaf(xv[j + dy], w.y, acc[j][4 * c + 1]);
                            acc[j][4 * c + 2] = fmaf(xv[j + dy], w.z, acc[j][4 * c + 2]);
                            acc[j][4 * c + 3] = fmaf(xv[j + dy], w.w, acc[j][4 * c + 3]);
                        }
                    }
                }
            }
        }
    }
    const int ox = x0 + tx;
    if (ox >= p.Wo) return;
#pragma unroll
    for (int j = 0; j < PX; ++j) {
        const int oy = y0 + tg * PX + j;
        if (oy >= p.Ho) continue;
        const long long o0 = (((long long)n * p.Ho + oy) * p.Wo + ox) * p.Co + co0;
        float v[CO];
#pragma unroll
        for (int c = 0; c < CO; ++c) {
            const int cc = co0 + c < p.Co ? co0 + c : p.Co - 1;             // clamped: values beyond Co are never stored
            float t = acc[j][c];
            if (p.bias) t += __ldg(p.bias + cc);
            t = apply_act(t, p.act1);
            if (p.post_scale) t = t * __ldg(p.post_scale + cc) + __ldg(p.post_shift + cc);
            t = apply_act(t, p.act2);
            if (p.residual) t += __ldg(p.residual + o0 - co0 + cc);
            if (p.out_tf32) { uint32_t r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(t)); t = __uint_as_float(r); }
            v[c] = t;
        }
        if (p.Co % 4 == 0) {                                                   // 128-bit stores: a pixel's CO channels are contiguous
#pragma unroll
            for (int c = 0; c < CO; c += 4)
                if (co0 + c < p.Co) *reinterpret_cast<float4*>(p.y + o0 + c) = make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]);
        } else {
#pragma unroll
            for (int c = 0; c < CO; ++c)
                if (co0 + c < p.Co) p.y[o0 + c] = v[c];
        }
    }
}

template <int CO, int KS>
static cudaError_t launch_conv_direct(const ConvF32& p, cudaStream_t s) {
    constexpr int HH = 32 + KS - 1, HWP = HH | 1;
    const int plane = HH * HWP + ((HH * HWP) % 2 == 0 ? 1 : 0);
    // channels per chunk: as many as fit in ~64 KB together with their weights (3 CTAs per SM when CO is small)
    int ck = p.Ci < 16 ? p.Ci : 16;
    auto bytes = [&](int k) { return (size_t)(((k * plane + 3) & ~3) + KS * KS * k * CO) * sizeof(float); };
    const size_t budget = CO >= 32 ? 160 * 1024 : 72 * 1024;         // CO = 32 is register-limited to one CTA per SM anyway
    while (ck > 1 && bytes(ck) > budget) --ck;
    ck = ceil_div(p.Ci, ceil_div(p.Ci, ck));                         // equal chunks
    const size_t smem = bytes(ck);
    static SmemAttrCache configured;
    if (cudaError_t e = ensure_dynamic_smem(conv_direct_f32_kernel<CO, KS>, smem, configured)) return e;
    const int tiles_x = ceil_div(p.Wo, 32), tiles_y = ceil_div(p.Ho, 32);
    dim3 grid((unsigned)(tiles_x * tiles_y), (unsigned)p.B, (unsigned)ceil_div(p.Co, CO));
    conv_direct_f32_kernel<CO, KS><<<grid, 256, smem, s>>>(p, ck, tiles_x);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// Pointwise (1x1, stride 1) convolution = a plain GEMM over the pixels: the predictor's expand / project layers and their
// input gradients.  The implicit-GEMM kernel above spends its time on im2col index arithmetic and scalar accesses there
// (0.44 TB/s on a 16 -> 64 layer); here a pixel row is loaded with 128-bit accesses and stored the same way.
// 64 pixels x 64 output channels per CTA, 4 x 4 register tiles, K in chunks of 32.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) conv1x1_f32_kernel(const ConvF32 p) {
    constexpr int BM = 64, BN = 64, BK = 32;
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN + 4];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const long long M = (long long)p.B * p.Ho * p.Wo;
    const long long m0 = (long long)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN, K = p.Ci;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int k0 = 0; k0 < K; k0 += BK) {
#pragma unroll
        for (int e = tid; e < BM * BK / 4; e += 256) {            // 8 float4 per pixel row of the chunk
            const int px = e >> 3, k4 = (e & 7) * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (m0 + px < M && k0 + k4 < K) {
                v = __ldg(reinterpret_cast<const float4*>(p.x + (m0 + px) * K + k0 + k4));
                v.x = v.x * p.in_scale + p.in_shift; v.y = v.y * p.in_scale + p.in_shift;
                v.z = v.z * p.in_scale + p.in_shift; v.w = v.w * p.in_scale + p.in_shift;
            }
            As[k4][px] = v.x; As[k4 + 1][px] = v.y; As[k4 + 2][px] = v.z; As[k4 + 3][px] = v.w;
        }
#pragma unroll
        for (int e = tid; e < BK * BN; e += 256) {
            const int k = e / BN, c = e % BN;
            Bs[k][c] = (k0 + k < K && n0 + c < p.Co) ? __ldg(p.w + (long long)(k0 + k) * p.w_ci + (long long)(n0 + c) * p.w_co) : 0.f;
        }
        __syncthreads();
#pragma unroll 8
        for (int k = 0; k < BK; ++k) {
            const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
    const int co0 = n0 + tx * 4;
    if (co0 >= p.Co) return;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long long m = m0 + ty * 4 + i;
        if (m >= M) continue;
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int co = co0 + j < p.Co ? co0 + j : p.Co - 1;   // clamped: never stored
            float t = acc[i][j];
            if (p.bias) t += __ldg(p.bias + co);
            t = apply_act(t, p.act1);
            if (p.post_scale) t = t * __ldg(p.post_scale + co) + __ldg(p.post_shift + co);
            t = apply_act(t, p.act2);
            if (p.residual) t += __ldg(p.residual + m * p.Co + co);
            if (p.out_tf32) { uint32_t r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(t)); t = __uint_as_float(r); }
            v[j] = t;
        }
        float* o = p.y + m * p.Co + co0;
        if (p.Co % 4 == 0) *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
        else {
#pragma unroll
            for (int j = 0; j < 4; ++j) if (co0 + j < p.Co) o[j] = v[j];
        }
    }
}

cudaError_t launch_conv_f32(const ConvF32& p, cudaStream_t s) {
    long long M = (long long)p.B * p.Ho * p.Wo;
    if (M == 0) return cudaSuccess;
    static const bool direct_off = [] { const char* e = ab_env("RST_CONV_DIRECT"); return e && e[0] == '0'; }();
    if (!direct_off && p.kh == 1 && p.kw == 1 && p.stride == 1 && p.pad_t == 0 && p.pad_l == 0 && p.Hi == p.Ho && p.Wi == p.Wo &&
        p.Ci % 4 == 0 && M >= 4096 && (reinterpret_cast<uintptr_t>(p.x) & 15) == 0 && (reinterpret_cast<uintptr_t>(p.y) & 15) == 0) {
        dim3 grid((unsigned)((M + 63) / 64), (unsigned)ceil_div(p.Co, 64));
        conv1x1_f32_kernel<<<grid, 256, 0, s>>>(p);
        return cudaGetLastError();
    }
    if (!direct_off && p.stride == 1 && p.kh == p.kw && p.Ho >= 64 && p.Wo >= 64 && p.B <= 65535) {
        if (p.kh == 9 && p.Co <= 4) return launch_conv_direct<4, 9>(p, s);
        if (p.kh == 9 && p.Co <= 16) return launch_conv_direct<16, 9>(p, s);
        if (p.kh == 9 && p.Co <= 32) return launch_conv_direct<32, 9>(p, s);
        if (p.kh == 3 && p.Co <= 4) return launch_conv_direct<4, 3>(p, s);
    }
    static const bool phases_off = [] { const char* e = ab_env("RST_CONV_PHASES"); return e && e[0] == '0'; }();
    if (p.transposed && p.stride == 2 && !phases_off) {
        // four launches, one per output parity, each over its own quarter of the pixels and its own subset of the taps
        for (int ph = 0; ph < 4; ++ph) {
            ConvF32 q = p;
            q.ph_y = ph >> 1; q.ph_x = ph & 1;
            const long long Mp = (long long)p.B * ((p.Ho - q.ph_y + 1) / 2) * ((p.Wo - q.ph_x + 1) / 2);
            if (Mp == 0) continue;
            if (p.Co > 32) {
                dim3 grid((unsigned)((Mp + 63) / 64), (unsigned)ceil_div(p.Co, 64));
                conv_f32_kernel<64, 64, 4, 4, true><<<grid, 256, 0, s>>>(q);
            } else if (p.Co > 16) {
                dim3 grid((unsigned)((Mp + 127) / 128), 1);
                conv_f32_kernel<128, 32, 4, 4, true><<<grid, 256, 0, s>>>(q);
            } else {
                dim3 grid((unsigned)((Mp + 127) / 128), 1);
                conv_f32_kernel<128, 16, 4, 2, true><<<grid, 256, 0, s>>>(q);
            }
        }
        return cudaGetLastError();
    }
    if (p.Co > 32) {
        dim3 grid((unsigned)((M + 63) / 64), (unsigned)ceil_div(p.Co, 64));
        conv_f32_kernel<64, 64, 4, 4, false><<<grid, 256, 0, s>>>(p);
    } else if (p.Co > 16) {                                       // 17..32 output channels: no half-empty 64-wide tile
        dim3 grid((unsigned)((M + 127) / 128), 1);
        conv_f32_kernel<128, 32, 4, 4, false><<<grid, 256, 0, s>>>(p);
    } else {
        dim3 grid((unsigned)((M + 127) / 128), 1);
        conv_f32_kernel<128, 16, 4, 2, false><<<grid, 256, 0, s>>>(p);
    }
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// depthwise conv (Keras DepthwiseConv2D, MobileNetV3 blocks).  One thread per output element.
// ---------------------------------------------------------------------------------------------
__global__ void depthwise_f32_kernel(const DepthwiseF32 p, long long total) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    int c = (int)(idx % p.C);
    long long t = idx / p.C;
    int ox = (int)(t % p.Wo);
    t /= p.Wo;
    int oy = (int)(t % p.Ho);
    int n = (int)(t / p.Ho);
    float acc = 0.f;
    for (int ky = 0; ky < p.k; ++ky) {
        int iy = oy * p.stride - p.pad_t + ky;
        if (iy < 0 || iy >= p.Hi) continue;
        for (int kx = 0; kx < p.k; ++kx) {
            int ix = ox * p.stride - p.pad_l + kx;
            if (ix < 0 || ix >= p.Wi) continue;
            acc = fmaf(__ldg(p.x + (((long long)n * p.Hi + iy) * p.Wi + ix) * p.C + c),
                       __ldg(p.w + (ky * p.k + kx) * p.C + c), acc);
        }
    }
    if (p.post_scale) acc = acc * __ldg(p.post_scale + c) + __ldg(p.post_shift + c);
    p.y[idx] = apply_act(acc, p.act);
}

cudaError_t launch_depthwise_f32(const DepthwiseF32& p, cudaStream_t s) {
    long long total = (long long)p.B * p.Ho * p.Wo * p.C;
    if (total == 0) return cudaSuccess;
    depthwise_f32_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(p, total);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// per-(n,c) sum / sum of squares over the H*W pixels, double accumulation
// ---------------------------------------------------------------------------------------------
__global__ void moments_f32_kernel(const float* __restrict__ x, double* __restrict__ stats, int P, int C,
                                   int pix_per_block) {
    extern __shared__ double sm[];
    const int n = blockIdx.y;
    const int p0 = blockIdx.x * pix_per_block;
    const int p1 = min(P, p0 + pix_per_block);
    const float* xb = x + (long long)n * P * C;
    if (C <= (int)blockDim.x) {
        const int G = blockDim.x / C;
        const int c = threadIdx.x % C;
        const int g = threadIdx.x / C;
        double s = 0.0, q = 0.0;
        if (g < G) {
            // four independent loads and accumulator pairs in flight (one dependent fp64 chain per load is latency bound)
            double s1 = 0.0, q1 = 0.0, s2 = 0.0, q2 = 0.0, s3 = 0.0, q3 = 0.0;
            int pp = p0 + g;
            for (; pp + 3 * G < p1; pp += 4 * G) {
                const float f0 = __ldg(xb + (long long)pp * C + c), f1 = __ldg(xb + (long long)(pp + G) * C + c);
                const float f2 = __ldg(xb + (long long)(pp + 2 * G) * C + c), f3 = __ldg(xb + (long long)(pp + 3 * G) * C + c);
                const double v0 = f0, v1 = f1, v2 = f2, v3 = f3;
                s += v0; q += v0 * v0; s1 += v1; q1 += v1 * v1; s2 += v2; q2 += v2 * v2; s3 += v3; q3 += v3 * v3;
            }
            for (; pp < p1; pp += G) {
                double v = (double)__ldg(xb + (long long)pp * C + c);
                s += v;
                q += v * v;
            }
            s += s1 + s2 + s3;
            q += q1 + q2 + q3;
        }
        sm[threadIdx.x] = s;
        sm[blockDim.x + threadIdx.x] = q;
        __syncthreads();
        if (g == 0) {
            for (int j = 1; j < G; ++j) {
                s += sm[j * C + c];
                q += sm[blockDim.x + j * C + c];
            }
            atomicAdd(&stats[((long long)n * C + c) * 2 + 0], s);
            atomicAdd(&stats[((long long)n * C + c) * 2 + 1], q);
        }
    } else {
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            double s = 0.0, q = 0.0;
            for (int pp = p0; pp < p1; ++pp) {
                double v = (double)__ldg(xb + (long long)pp * C + c);
                s += v;
                q += v * v;
            }
            atomicAdd(&stats[((long long)n * C + c) * 2 + 0], s);
            atomicAdd(&stats[((long long)n * C + c) * 2 + 1], q);
        }
    }
}

cudaError_t launch_moments_f32(const float* x, double* stats, int B, int P, int C, cudaStream_t s) {
    if (B == 0 || P == 0) return cudaSuccess;
    int threads = C <= 256 ? C * (256 / C) : 256;
    // enough CTAs to fill the GPU (8 x 148) even for one small tensor, long enough runs to amortise the atomics
    int pix_per_block = 2048;
    while (pix_per_block > 256 && (long long)ceil_div(P, pix_per_block) * B < 148 * 8) pix_per_block /= 2;
    dim3 grid((unsigned)ceil_div(P, pix_per_block), (unsigned)B);
    moments_f32_kernel<<<grid, threads, 2 * threads * sizeof(double), s>>>(x, stats, P, C, pix_per_block);
    return cudaGetLastError();
}

__global__ void zero_f64_kernel(double* p, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = 0.0;
}
cudaError_t launch_zero_f64(double* p, long long n, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    zero_f64_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(p, n);
    return cudaGetLastError();
}

__global__ void stats_to_mean_kernel(const double* stats, float* mean, long long n, double invP) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) mean[i] = (float)(stats[i * 2] * invP);
}
cudaError_t launch_stats_to_mean(const double* stats, float* mean, int B, int P, int C, cudaStream_t s) {
    long long n = (long long)B * C;
    if (n == 0) return cudaSuccess;
    stats_to_mean_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(stats, mean, n, 1.0 / (double)P);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// conditional instance norm apply; follows the reference's operation order:
//   xh = x*inv + (-mean*inv);  y = bias + xh*scale   (styleTransfer.py:65-68)
// ---------------------------------------------------------------------------------------------
template <bool XBF, bool YBF>
__global__ void cin_apply_kernel(const CinApply p, int pix_per_block) {
    extern __shared__ float smf[];
    const int C = p.C;
    float* s_inv = smf;              // [C]
    float* s_nmi = smf + C;          // [C]   -mean*inv
    float* s_scale = smf + 2 * C;    // [S][C]
    float* s_bias = s_scale + p.num_styles * C;
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
    const long long base = (long long)n * p.P * C;
    const long long e0 = (long long)blockIdx.x * pix_per_block * C;
    const long long e1 = min((long long)p.P * C, e0 + (long long)pix_per_block * C);
    const bool blend = p.num_styles == 2 && p.weights != nullptr;
    for (long long e = e0 + threadIdx.x; e < e1; e += blockDim.x) {
        int c = (int)(e % C);
        float xv;
        if (XBF) xv = __bfloat162float(((const __nv_bfloat16*)p.x)[base + e]);
        else xv = ((const float*)p.x)[base + e];
        float scale, bias;
        if (blend) {
            long long pix = e / C;
            const float* w = p.weights + ((long long)n * p.P + pix) * 2;
            float w0 = __ldg(w), w1 = __ldg(w + 1);
            // reduce_sum over the style axis of params*weights (styleTransfer.py:41-42)
            scale = s_scale[c] * w0 + s_scale[C + c] * w1;
            bias = s_bias[c] * w0 + s_bias[C + c] * w1;
        } else {
            scale = s_scale[c];
            bias = s_bias[c];
        }
        float xh = xv * s_inv[c] + s_nmi[c];
        float v = bias + xh * scale;
        v = apply_act(v, p.act);
        if (p.residual) {
            if (YBF) v += __bfloat162float(((const __nv_bfloat16*)p.residual)[base + e]);
            else v += ((const float*)p.residual)[base + e];
        }
        if (YBF) ((__nv_bfloat16*)p.y)[base + e] = __float2bfloat16(v);
        else ((float*)p.y)[base + e] = v;
    }
}

// fp32 -> fp32 with P * C % 4 == 0: one float4 (four consecutive elements, possibly across a pixel boundary: C = 3) per thread
// and iteration, 32-bit index arithmetic; per element the same operations in the same order as cin_apply_kernel above
// (results are bit-identical).
__global__ void __launch_bounds__(256) cin_apply_f32x4_kernel(const CinApply p, int vec_per_block) {
    extern __shared__ float smf[];
    const int C = p.C;
    float* s_inv = smf;
    float* s_nmi = smf + C;
    float* s_scale = smf + 2 * C;
    float* s_bias = s_scale + p.num_styles * C;
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
    const long long base4 = (long long)n * p.P * C / 4;
    const float4* x4 = reinterpret_cast<const float4*>(p.x) + base4;
    const float4* r4 = p.residual ? reinterpret_cast<const float4*>(p.residual) + base4 : nullptr;
    float4* y4 = reinterpret_cast<float4*>(p.y) + base4;
    const int total4 = (int)((long long)p.P * C / 4);               // P * C < 2^31 (checked by the launcher)
    const int v0 = blockIdx.x * vec_per_block, v1 = min(total4, v0 + vec_per_block);
    const bool blend = p.num_styles == 2 && p.weights != nullptr;
    const float2* w2 = blend ? reinterpret_cast<const float2*>(p.weights) + (long long)n * p.P : nullptr;
    for (int v = v0 + threadIdx.x; v < v1; v += blockDim.x) {
        int pix = (4 * v) / C, c = 4 * v - pix * C;
        const float4 xv = x4[v];
        const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
        float rs[4] = {0.f, 0.f, 0.f, 0.f};
        if (r4) { const float4 rv = r4[v]; rs[0] = rv.x; rs[1] = rv.y; rs[2] = rv.z; rs[3] = rv.w; }
        float2 w = make_float2(0.f, 0.f);
        if (blend) w = __ldg(w2 + pix);
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float scale, bias;
            if (blend) {
                scale = s_scale[c] * w.x + s_scale[C + c] * w.y;
                bias = s_bias[c] * w.x + s_bias[C + c] * w.y;
            } else {
                scale = s_scale[c];
                bias = s_bias[c];
            }
            const float xh = xs[j] * s_inv[c] + s_nmi[c];
            float t = bias + xh * scale;
            t = apply_act(t, p.act);
            if (r4) t += rs[j];
            o[j] = t;
            if (++c == C) {                                          // next pixel
                c = 0; ++pix;
                if (blend && j < 3 && pix < p.P) w = __ldg(w2 + pix);
            }
        }
        y4[v] = make_float4(o[0], o[1], o[2], o[3]);
    }
}

cudaError_t launch_cin_apply(const CinApply& p, cudaStream_t s) {
    if (p.B == 0 || p.P == 0) return cudaSuccess;
    if (!p.x_bf16 && !p.y_bf16 && ((long long)p.P * p.C) % 4 == 0 && (long long)p.P * p.C < (1ll << 31) &&
        (reinterpret_cast<uintptr_t>(p.x) & 15) == 0 && (reinterpret_cast<uintptr_t>(p.y) & 15) == 0 &&
        (!p.residual || (reinterpret_cast<uintptr_t>(p.residual) & 15) == 0)) {
        // 128 KB read per CTA on large tensors; small ones (batch 1 at the bottleneck: 3.7 MB) are cut into ~4 CTAs per SM so that
        // the pass is not a serial chain of load -> store round trips on a fifth of the SMs
        const long long total4 = (long long)p.P * p.C / 4;
        int vec_per_block = (int)(total4 * p.B / (4 * 148));
        vec_per_block = vec_per_block < 512 ? 512 : vec_per_block > 8192 ? 8192 : (vec_per_block + 255) / 256 * 256;
        dim3 grid((unsigned)ceil_div((int)((long long)p.P * p.C / 4), vec_per_block), (unsigned)p.B);
        size_t smem = (size_t)(2 + 2 * p.num_styles) * p.C * sizeof(float);
        cin_apply_f32x4_kernel<<<grid, 256, smem, s>>>(p, vec_per_block);
        return cudaGetLastError();
    }
    int pix_per_block = max(1, 16384 / p.C);
    dim3 grid((unsigned)ceil_div(p.P, pix_per_block), (unsigned)p.B);
    size_t smem = (size_t)(2 + 2 * p.num_styles) * p.C * sizeof(float);
    if (p.x_bf16 && p.y_bf16) cin_apply_kernel<true, true><<<grid, 256, smem, s>>>(p, pix_per_block);
    else if (p.x_bf16) cin_apply_kernel<true, false><<<grid, 256, smem, s>>>(p, pix_per_block);
    else if (p.y_bf16) cin_apply_kernel<false, true><<<grid, 256, smem, s>>>(p, pix_per_block);
    else cin_apply_kernel<false, false><<<grid, 256, smem, s>>>(p, pix_per_block);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// style-weight handling
// ---------------------------------------------------------------------------------------------
// (pixels, S-1) -> (pixels, S) = [1 - sum(w), w...]   (styleTransfer.py:297-302)
__global__ void weights_concat_kernel(const float* __restrict__ w_in, float* __restrict__ w_out, long long pixels,
                                      int sm1) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= pixels) return;
    float sum = 0.f;
    for (int k = 0; k < sm1; ++k) {
        float v = w_in[i * sm1 + k];
        sum += v;
        w_out[i * (sm1 + 1) + 1 + k] = v;
    }
    w_out[i * (sm1 + 1)] = 1.f - sum;
}
// Two styles: the concat [1 - w, w] (styleTransfer.py:297-302) and its first two AvgPool2D(2) levels (:335-345) in ONE pass.
// A thread owns a 4 x 4 block of the weight map: four float4 loads, then 16 + 4 + 1 two-channel pixels out (the pooled levels
// are averages of averages, as the reference's chained pooling layers compute them).
__global__ void weights_pyramid3_kernel(const float* __restrict__ w_in, float* __restrict__ l0, float* __restrict__ l1,
                                        float* __restrict__ l2, int H, int W, long long blocks) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= blocks) return;
    const int bw = W / 4, bh = H / 4;
    const int bx = (int)(i % bw);
    const long long t = i / bw;
    const int by = (int)(t % bh);
    const long long n = t / bh;
    float4 r[4];
#pragma unroll
    for (int y = 0; y < 4; ++y) r[y] = *reinterpret_cast<const float4*>(w_in + ((n * H + 4 * by + y) * W + 4 * bx));
    float m1[2][2][2];                                    // [row][col][channel] of level 1
#pragma unroll
    for (int y = 0; y < 4; ++y) {
        const float w1[4] = {r[y].x, r[y].y, r[y].z, r[y].w};
        float4* o = reinterpret_cast<float4*>(l0 + ((n * H + 4 * by + y) * W + 4 * bx) * 2);
        o[0] = make_float4(1.f - w1[0], w1[0], 1.f - w1[1], w1[1]);
        o[1] = make_float4(1.f - w1[2], w1[2], 1.f - w1[3], w1[3]);
    }
#pragma unroll
    for (int yy = 0; yy < 2; ++yy)
#pragma unroll
        for (int xx = 0; xx < 2; ++xx) {
            const float a[4] = {xx ? r[2 * yy].z : r[2 * yy].x, xx ? r[2 * yy].w : r[2 * yy].y,
                                xx ? r[2 * yy + 1].z : r[2 * yy + 1].x, xx ? r[2 * yy + 1].w : r[2 * yy + 1].y};
            // the same summation order as the generic 2x2 pooling kernel: top-left + top-right + bottom-left + bottom-right
            m1[yy][xx][0] = ((1.f - a[0]) + (1.f - a[1]) + (1.f - a[2]) + (1.f - a[3])) * 0.25f;
            m1[yy][xx][1] = (a[0] + a[1] + a[2] + a[3]) * 0.25f;
        }
    const int H1 = H / 2, W1 = W / 2;
#pragma unroll
    for (int yy = 0; yy < 2; ++yy)
        *reinterpret_cast<float4*>(l1 + ((n * H1 + 2 * by + yy) * W1 + 2 * bx) * 2) =
            make_float4(m1[yy][0][0], m1[yy][0][1], m1[yy][1][0], m1[yy][1][1]);
    *reinterpret_cast<float2*>(l2 + ((n * (H / 4) + by) * (W / 4) + bx) * 2) =
        make_float2((m1[0][0][0] + m1[0][1][0] + m1[1][0][0] + m1[1][1][0]) * 0.25f,
                    (m1[0][0][1] + m1[0][1][1] + m1[1][0][1] + m1[1][1][1]) * 0.25f);
}
cudaError_t launch_weights_pyramid3(const float* w_in, float* l0, float* l1, float* l2, int B, int H, int W, cudaStream_t s) {
    const long long blocks = (long long)B * (H / 4) * (W / 4);
    if (blocks == 0) return cudaSuccess;
    weights_pyramid3_kernel<<<(unsigned)((blocks + 255) / 256), 256, 0, s>>>(w_in, l0, l1, l2, H, W, blocks);
    return cudaGetLastError();
}

cudaError_t launch_weights_concat(const float* w_in, float* w_out, long long pixels, int sm1, cudaStream_t s) {
    if (pixels == 0) return cudaSuccess;
    weights_concat_kernel<<<(unsigned)((pixels + 255) / 256), 256, 0, s>>>(w_in, w_out, pixels, sm1);
    return cudaGetLastError();
}

// V channels per thread (V = 4: 128-bit accesses when C % 4 == 0 and the tensors are 16-byte aligned)
template <bool MAX, int V>
__global__ void pool2_f32_kernel(const float* __restrict__ x, float* __restrict__ y, int Hi, int Wi, int C,
                                 long long total) {
    const long long idx = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * V;
    if (idx >= total) return;
    const int Ho = Hi / 2, Wo = Wi / 2;
    int c = (int)(idx % C);
    long long t = idx / C;
    int ox = (int)(t % Wo);
    t /= Wo;
    int oy = (int)(t % Ho);
    long long n = t / Ho;
    const float* b = x + ((n * Hi + 2 * oy) * Wi + 2 * ox) * C + c;
    if (V == 4) {
        const float4 v00 = *reinterpret_cast<const float4*>(b), v01 = *reinterpret_cast<const float4*>(b + C);
        const float4 v10 = *reinterpret_cast<const float4*>(b + (long long)Wi * C), v11 = *reinterpret_cast<const float4*>(b + (long long)Wi * C + C);
        float4 o;
        o.x = MAX ? fmaxf(fmaxf(v00.x, v01.x), fmaxf(v10.x, v11.x)) : (v00.x + v01.x + v10.x + v11.x) * 0.25f;
        o.y = MAX ? fmaxf(fmaxf(v00.y, v01.y), fmaxf(v10.y, v11.y)) : (v00.y + v01.y + v10.y + v11.y) * 0.25f;
        o.z = MAX ? fmaxf(fmaxf(v00.z, v01.z), fmaxf(v10.z, v11.z)) : (v00.z + v01.z + v10.z + v11.z) * 0.25f;
        o.w = MAX ? fmaxf(fmaxf(v00.w, v01.w), fmaxf(v10.w, v11.w)) : (v00.w + v01.w + v10.w + v11.w) * 0.25f;
        *reinterpret_cast<float4*>(y + idx) = o;
    } else {
        float v00 = b[0], v01 = b[C], v10 = b[(long long)Wi * C], v11 = b[(long long)Wi * C + C];
        y[idx] = MAX ? fmaxf(fmaxf(v00, v01), fmaxf(v10, v11)) : (v00 + v01 + v10 + v11) * 0.25f;
    }
}
template <bool MAX>
static cudaError_t launch_pool2(const float* x, float* y, int B, int Hi, int Wi, int C, cudaStream_t s) {
    long long total = (long long)B * (Hi / 2) * (Wi / 2) * C;
    if (total == 0) return cudaSuccess;
    const bool vec = C % 4 == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
    if (vec) pool2_f32_kernel<MAX, 4><<<(unsigned)((total / 4 + 255) / 256), 256, 0, s>>>(x, y, Hi, Wi, C, total);
    else pool2_f32_kernel<MAX, 1><<<(unsigned)((total + 255) / 256), 256, 0, s>>>(x, y, Hi, Wi, C, total);
    return cudaGetLastError();
}
cudaError_t launch_avgpool2_f32(const float* x, float* y, int B, int Hi, int Wi, int C, cudaStream_t s) {
    return launch_pool2<false>(x, y, B, Hi, Wi, C, s);
}
cudaError_t launch_maxpool2_f32(const float* x, float* y, int B, int Hi, int Wi, int C, cudaStream_t s) {
    return launch_pool2<true>(x, y, B, Hi, Wi, C, s);
}

__global__ void scale_channels_kernel(const float* __restrict__ x, const float* __restrict__ z,
                                      float* __restrict__ y, long long PC, int C, long long total) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    long long n = idx / PC;
    int c = (int)(idx % C);
    y[idx] = x[idx] * __ldg(z + n * C + c);
}
cudaError_t launch_scale_channels(const float* x, const float* z, float* y, int B, int P, int C, cudaStream_t s) {
    long long total = (long long)B * P * C;
    if (total == 0) return cudaSuccess;
    scale_channels_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(x, z, y, (long long)P * C, C, total);
    return cudaGetLastError();
}

// _apply_style_weights (styleTransfer.py:36-44): out[b,p,f] = sum_s w[b,p,s]*params[b,0,s,f], S == 2
__global__ void apply_style_weights_kernel(const float* __restrict__ w, const float* __restrict__ params,
                                           float* __restrict__ out, long long P, int F, long long total) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    int f = (int)(idx % F);
    long long bp = idx / F;
    long long b = bp / P;
    const float* pb = params + b * 2 * F;
    out[idx] = pb[f] * w[bp * 2] + pb[F + f] * w[bp * 2 + 1];
}
cudaError_t launch_apply_style_weights(const float* w, const float* params, float* out, int B, long long P, int F,
                                       cudaStream_t s) {
    long long total = (long long)B * P * F;
    if (total == 0) return cudaSuccess;
    apply_style_weights_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(w, params, out, P, F, total);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// Gram matrix (styleLoss.py:11-18), fp32 CUDA-core version: G[b] = F^T F / P, split over pixels.
// ---------------------------------------------------------------------------------------------
__global__ void gram_f32_kernel(const float* __restrict__ x, float* __restrict__ g, int P, int C, int pix_per_split,
                                float invP) {
    __shared__ float Fc[32][33];
    __shared__ float Fd[32][33];
    const int c0 = blockIdx.x * 32, d0 = blockIdx.y * 32;
    const int splits = ceil_div(P, pix_per_split);
    const int b = blockIdx.z / splits, sp = blockIdx.z % splits;
    const int p0 = sp * pix_per_split, p1 = min(P, p0 + pix_per_split);
    const float* xb = x + (long long)b * P * C;
    const int tx = threadIdx.x, ty = threadIdx.y;   // 32 x 8
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int pp = p0; pp < p1; pp += 32) {
        for (int r = ty; r < 32; r += 8) {
            int pix = pp + r;
            Fc[r][tx] = (pix < p1 && c0 + tx < C) ? xb[(long long)pix * C + c0 + tx] : 0.f;
            Fd[r][tx] = (pix < p1 && d0 + tx < C) ? xb[(long long)pix * C + d0 + tx] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < 32; ++r) {
            float d = Fd[r][tx];
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[i] = fmaf(Fc[r][ty + 8 * i], d, acc[i]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int c = c0 + ty + 8 * i, d = d0 + tx;
        if (c < C && d < C) atomicAdd(&g[((long long)b * C + c) * C + d], acc[i] * invP);
    }
}
// C % 64 == 0 (the VGG16 taps): 64 x 64 tile of G per CTA, 4 x 4 register tile per thread, 128-bit shared-memory reads
__global__ void __launch_bounds__(256) gram64_f32_kernel(const float* __restrict__ x, float* __restrict__ g, int P, int C,
                                                         int pix_per_split, float invP) {
    __shared__ __align__(16) float Fc[32][68];
    __shared__ __align__(16) float Fd[32][68];
    const int c0 = blockIdx.x * 64, d0 = blockIdx.y * 64;
    const int splits = ceil_div(P, pix_per_split);
    const int b = blockIdx.z / splits, sp = blockIdx.z % splits;
    const int p0 = sp * pix_per_split, p1 = min(P, p0 + pix_per_split);
    const float* xb = x + (long long)b * P * C;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int pp = p0; pp < p1; pp += 32) {
        // 32 pixels x 64 channels per operand = 512 float4: two per thread and operand, coalesced along the channels
#pragma unroll
        for (int e = threadIdx.x; e < 512; e += 256) {
            const int r = e >> 4, c4 = (e & 15) * 4, pix = pp + r;
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f), d = a;
            if (pix < p1) {
                a = __ldg(reinterpret_cast<const float4*>(xb + (long long)pix * C + c0 + c4));
                d = __ldg(reinterpret_cast<const float4*>(xb + (long long)pix * C + d0 + c4));
            }
            *reinterpret_cast<float4*>(&Fc[r][c4]) = a;
            *reinterpret_cast<float4*>(&Fd[r][c4]) = d;
        }
        __syncthreads();
#pragma unroll 8
        for (int r = 0; r < 32; ++r) {
            const float4 a = *reinterpret_cast<const float4*>(&Fc[r][ty * 4]);
            const float4 d = *reinterpret_cast<const float4*>(&Fd[r][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, dv[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], dv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
            atomicAdd(&g[((long long)b * C + c0 + ty * 4 + i) * C + d0 + tx * 4 + j], acc[i][j] * invP);
}

cudaError_t launch_gram_f32(const float* x, float* g, int B, int P, int C, cudaStream_t s) {
    if (B == 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(g, 0, (size_t)B * C * C * sizeof(float), s);
    if (e != cudaSuccess) return e;
    if (C % 64 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
        const int tiles = (C / 64) * (C / 64);
        int pix_per_split = 4096;
        while (pix_per_split > 256 && (long long)ceil_div(P, pix_per_split) * B * tiles < 148 * 4) pix_per_split /= 2;
        const int splits = ceil_div(P, pix_per_split);
        dim3 grid((unsigned)(C / 64), (unsigned)(C / 64), (unsigned)(B * splits));
        gram64_f32_kernel<<<grid, 256, 0, s>>>(x, g, P, C, pix_per_split, 1.f / (float)P);
        return cudaGetLastError();
    }
    int pix_per_split = 4096;
    int splits = ceil_div(P, pix_per_split);
    dim3 grid((unsigned)ceil_div(C, 32), (unsigned)ceil_div(C, 32), (unsigned)(B * splits));
    gram_f32_kernel<<<grid, dim3(32, 8), 0, s>>>(x, g, P, C, pix_per_split, 1.f / (float)P);
    return cudaGetLastError();
}

}  // namespace rst
