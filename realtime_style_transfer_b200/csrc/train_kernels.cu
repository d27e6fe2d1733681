// fp32 kernels for the training step: weight gradients, normalisation forward (training mode) and backward,
// activation backward, RMSprop.  Correctness-first CUDA-core kernels (the tensor-core training path is future work).
#include "train_kernels.cuh"

#include <cstdlib>
#include <initializer_list>

namespace rst {

static inline unsigned nblk(long long n) { return (unsigned)((n + 255) / 256); }

// V consecutive floats per thread (V = 4: one 128-bit access) for the elementwise passes over NHWC tensors
template <int V> __device__ __forceinline__ void ldv(const float* p, float* v);
template <> __device__ __forceinline__ void ldv<1>(const float* p, float* v) { v[0] = *p; }
template <> __device__ __forceinline__ void ldv<4>(const float* p, float* v) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
template <int V> __device__ __forceinline__ void stv(float* p, const float* v);
template <> __device__ __forceinline__ void stv<1>(float* p, const float* v) { *p = v[0]; }
template <> __device__ __forceinline__ void stv<4>(float* p, const float* v) { *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); }
// V = 4 needs whole groups of 4 channels and 16-byte aligned tensors
static inline bool vec4_ok(long long total, int C, std::initializer_list<const void*> ptrs) {
    if (total % 4 != 0 || C % 4 != 0) return false;
    for (const void* q : ptrs) if (reinterpret_cast<uintptr_t>(q) & 15) return false;
    return true;
}

// ---------------------------------------------------------------------------------------------------------------
// weight gradient of Conv2D / Conv2DTranspose ('same'), split over base pixels with fp32 atomics.
//   conv : dW[ky,kx,ci,co] = sum_{n,oy,ox} x[n, oy*s - pt + ky, ox*s - pl + kx, ci] * g[n,oy,ox,co]     (base = output grid)
//   convT: dW[ky,kx,co,ci] = sum_{n,i,j}  x[n,i,j,ci] * g[n, i*s + ky - pt, j*s + kx - pl, co]         (base = input grid)
// One CTA = one filter tap x 32 input channels x 32 output channels x a slab of base pixels.
// ---------------------------------------------------------------------------------------------------------------
// TM x TN = input-channel x output-channel tile of one tap; 16 x 16 threads, each a (TM/16) x (TN/16) register tile.
// RMODE 1 (Conv2D): the M dimension of a CTA is a block of (kx, ci) pairs of ONE filter row ky -- for a fixed output
// pixel these kw*Ci inputs are contiguous in NHWC memory, so thin layers (17 input channels, 81 taps) fill the tile.
// RMODE 2 (Conv2DTranspose): the same on the gradient side, N = block of (kx, co) pairs (16 -> 3 head: 27 columns).
template <int TM, int TN, int RMODE>
__global__ void __launch_bounds__(256) wgrad_f32_kernel(const WgradF32 p, int ci_blocks, int pix_per_split) {
    constexpr int RM = TM / 16, RN = TN / 16, PIX = 32;
    __shared__ __align__(16) float As[PIX][TM + 4];      // [pixel][ci]  (ROWMODE: [pixel][kx*Ci + ci])
    __shared__ __align__(16) float Gs[PIX][TN + 4];      // [pixel][co]
    constexpr bool ROWMODE = RMODE == 1, GROW = RMODE == 2;
    __shared__ int xcol[PIX];                             // ROWMODE / GROW: column of kx = 0 in x / g
    const int tap = blockIdx.x / ci_blocks, ci0 = (blockIdx.x % ci_blocks) * TM, c0 = blockIdx.y * TN;
    const int ky = (ROWMODE || GROW) ? tap : tap / p.kw, kx = (ROWMODE || GROW) ? 0 : tap - ky * p.kw;
    const int row_len = p.kw * (GROW ? p.Co : p.Ci);      // ROWMODE: M extent, GROW: N extent
    const long long NP = (long long)p.B * p.Hb * p.Wb;
    const long long p0 = (long long)blockIdx.z * pix_per_split;
    const long long p1 = min(NP, p0 + pix_per_split);
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    float acc[RM][RN];
#pragma unroll
    for (int i = 0; i < RM; ++i)
#pragma unroll
        for (int j = 0; j < RN; ++j) acc[i][j] = 0.f;
    __shared__ long long xoff[PIX], goff[PIX];            // element offset of each pixel's channel 0 in x / g, -1 = zero
    for (long long pp = p0; pp < p1; pp += PIX) {
        if (threadIdx.x < PIX) {
            const long long pix = pp + threadIdx.x;
            long long xo = -1, go = -1;
            if (pix < p1) {
                const int bx = (int)(pix % p.Wb);
                const long long t = pix / p.Wb;
                const int by = (int)(t % p.Hb), n = (int)(t / p.Hb);
                int xy = by, xx = bx, gy_ = by, gx_ = bx;
                if (!p.transposed) { xy = by * p.stride - p.pad_t + ky; xx = bx * p.stride - p.pad_l + kx; }
                else { gy_ = by * p.stride + ky - p.pad_t; gx_ = bx * p.stride + kx - p.pad_l; }
                if (ROWMODE) {
                    // offset of (row, column of kx = 0, channel 0); the column may lie left of the image, elements are checked
                    if (xy >= 0 && xy < p.Hx) xo = (((long long)n * p.Hx + xy) * p.Wx + xx) * p.Ci + (1LL << 40);
                    xcol[threadIdx.x] = xx;
                } else if (xy >= 0 && xy < p.Hx && xx >= 0 && xx < p.Wx) xo = (((long long)n * p.Hx + xy) * p.Wx + xx) * p.Ci;
                if (GROW) {
                    if (gy_ >= 0 && gy_ < p.Hg) go = (((long long)n * p.Hg + gy_) * p.Wg + gx_) * p.Co + (1LL << 40);
                    xcol[threadIdx.x] = gx_;
                } else
                if (gy_ >= 0 && gy_ < p.Hg && gx_ >= 0 && gx_ < p.Wg) go = (((long long)n * p.Hg + gy_) * p.Wg + gx_) * p.Co;
            }
            xoff[threadIdx.x] = xo; goff[threadIdx.x] = go;
        }
        __syncthreads();
        // cooperative loads: consecutive threads read consecutive channels of one pixel (coalesced NHWC reads)
#pragma unroll
        for (int e = threadIdx.x; e < PIX * TM; e += 256) {
            const int q = e / TM, c = e % TM;
            const long long xo = xoff[q];
            if (ROWMODE) {
                const int m = ci0 + c, col = xcol[q] + m / p.Ci;          // m / Ci is loop-invariant per thread (256 % TM == 0)
                As[q][c] = (xo >= 0 && m < row_len && col >= 0 && col < p.Wx)
                               ? fmaf(__ldg(p.x + (xo - (1LL << 40)) + m), p.in_scale, p.in_shift) : 0.f;
            } else
            As[q][c] = (xo >= 0 && ci0 + c < p.Ci) ? fmaf(__ldg(p.x + xo + ci0 + c), p.in_scale, p.in_shift) : 0.f;
        }
#pragma unroll
        for (int e = threadIdx.x; e < PIX * TN; e += 256) {
            const int q = e / TN, c = e % TN;
            const long long go = goff[q];
            if (GROW) {
                const int nn = c0 + c, col = xcol[q] + nn / p.Co;
                Gs[q][c] = (go >= 0 && nn < row_len && col >= 0 && col < p.Wg) ? __ldg(p.g + (go - (1LL << 40)) + nn) : 0.f;
            } else
            Gs[q][c] = (go >= 0 && c0 + c < p.Co) ? __ldg(p.g + go + c0 + c) : 0.f;
        }
        __syncthreads();
#pragma unroll 8
        for (int q = 0; q < PIX; ++q) {
            float a[RM], g[RN];
            if (RM % 4 == 0) {
#pragma unroll
                for (int h = 0; h < RM / 4; ++h) {       // RM = 8: columns ty*4 and 64 + ty*4 (conflict-free float4 reads)
                    const float4 t = *reinterpret_cast<const float4*>(&As[q][h * 64 + ty * 4]);
                    a[4 * h] = t.x; a[4 * h + 1] = t.y; a[4 * h + 2] = t.z; a[4 * h + 3] = t.w;
                }
            } else { const float2 t = *reinterpret_cast<const float2*>(&As[q][ty * 2]); a[0] = t.x; a[1] = t.y; }
            if (RN % 4 == 0) {
#pragma unroll
                for (int h = 0; h < RN / 4; ++h) {
                    const float4 t = *reinterpret_cast<const float4*>(&Gs[q][h * 64 + tx * 4]);
                    g[4 * h] = t.x; g[4 * h + 1] = t.y; g[4 * h + 2] = t.z; g[4 * h + 3] = t.w;
                }
            } else { const float2 t = *reinterpret_cast<const float2*>(&Gs[q][tx * 2]); g[0] = t.x; g[1] = t.y; }
#pragma unroll
            for (int i = 0; i < RM; ++i)
#pragma unroll
                for (int j = 0; j < RN; ++j) acc[i][j] = fmaf(a[i], g[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < RM; ++i)
#pragma unroll
        for (int j = 0; j < RN; ++j) {
            // register i of a thread is tile row (i / 4) * 64 + ty * 4 + i % 4 when the tile is read as float4 halves (RM = 8)
            const int cin = ci0 + (RM % 4 == 0 ? (i / 4) * 64 + ty * 4 + i % 4 : ty * RM + i);
            const int co = c0 + (RN % 4 == 0 ? (j / 4) * 64 + tx * 4 + j % 4 : tx * RN + j);
            if (ROWMODE) {
                if (cin < row_len && co < p.Co) atomicAdd(p.dw + ((long long)ky * row_len + cin) * p.Co + co, acc[i][j]);
            } else if (GROW) {
                if (cin < p.Ci && co < row_len) atomicAdd(p.dw + ((long long)ky * row_len + co) * p.Ci + cin, acc[i][j]);
            } else if (cin < p.Ci && co < p.Co) {
                const long long idx = !p.transposed ? ((long long)tap * p.Ci + cin) * p.Co + co : ((long long)tap * p.Co + co) * p.Ci + cin;
                atomicAdd(p.dw + idx, acc[i][j]);
            }
        }
}

// ---------------------------------------------------------------------------------------------------------------
// Weight gradient of the thin 9x9 stride-1 layers at full resolution (stem 17 -> 32, head 16 -> 3): a sliding-window kernel.
//   dW[ky, kx, a, b] = sum_{n,y,x} P[n, y + ky - pt, x + kx - pl, a] * Q[n, y, x, b]
// conv: P = layer input (A = Ci), Q = output gradient (NB = Co); convT: P = output gradient (A = Co), Q = layer input (NB = Ci).
// A thread owns dW[ky, 0..KS-1, a, 4 consecutive b]: walking along an image row it loads ONE new P value and one float4 of Q
// per pixel and keeps the KS-wide P window in registers (KS * 4 FMAs per 2 shared-memory loads; the GEMM-style kernel above
// manages 8 per 2).  A CTA covers KYB filter rows and stages 4 x 72-pixel tiles (+ halo) in shared memory; out-of-image
// pixels are stored as zeros, so no masks are needed in the inner loop.  CTAs are persistent: one atomicAdd per weight each.
// ---------------------------------------------------------------------------------------------------------------
struct WgradDirect {
    const float* P; const float* Q; float* dw;
    int B, H, W, NB, pad_t, pad_l;
    float p_scale, p_shift;
};
template <int A, int NBG, int KS, int KYB>
__global__ void __launch_bounds__((KYB * A * NBG + 31) / 32 * 32)
wgrad_direct_f32_kernel(const WgradDirect p, int tiles_x, int tiles_y, int blocks_per_group) {
    constexpr int TH = 4, TW = 8 * KS, PW = TW + KS - 1, PH = TH + KYB - 1, NB4 = NBG * 4, NT = KYB * A * NBG;
    extern __shared__ __align__(16) float wd_smem[];
    float* Qs = wd_smem;                                   // [TH][TW][NB4]
    float* Ps = wd_smem + TH * TW * NB4;                   // [PH][PW][A]
    const int tid = threadIdx.x, nthreads = blockDim.x;
    const int kg = blockIdx.x / blocks_per_group, bi = blockIdx.x % blocks_per_group;
    const int ky0 = kg * KYB;
    const bool active = tid < NT;
    const int bg = tid % NBG, a = (tid / NBG) % A, kyl = active ? tid / (NBG * A) : 0;
    float acc[KS][4];
#pragma unroll
    for (int k = 0; k < KS; ++k) { acc[k][0] = 0.f; acc[k][1] = 0.f; acc[k][2] = 0.f; acc[k][3] = 0.f; }
    const long long total_tiles = (long long)p.B * tiles_y * tiles_x;
    for (long long t = bi; t < total_tiles; t += blocks_per_group) {
        const int txi = (int)(t % tiles_x);
        const long long rr = t / tiles_x;
        const int tyi = (int)(rr % tiles_y), n = (int)(rr / tiles_y);
        const int x0 = txi * TW, y0 = tyi * TH;
        const long long img = (long long)n * p.H * p.W;
        __syncthreads();
        for (int e = tid; e < TH * TW * NB4; e += nthreads) {
            const int b = e % NB4, pix = e / NB4, gx = x0 + pix % TW, gy = y0 + pix / TW;
            Qs[e] = (b < p.NB && gy < p.H && gx < p.W) ? __ldg(p.Q + (img + (long long)gy * p.W + gx) * p.NB + b) : 0.f;
        }
        for (int e = tid; e < PH * PW * A; e += nthreads) {
            const int c = e % A, pix = e / A, px = x0 + pix % PW - p.pad_l, py = y0 + pix / PW + ky0 - p.pad_t;
            Ps[e] = (py >= 0 && py < p.H && px >= 0 && px < p.W)
                        ? fmaf(__ldg(p.P + (img + (long long)py * p.W + px) * A + c), p.p_scale, p.p_shift) : 0.f;
        }
        __syncthreads();
        if (active) {
#pragma unroll 1
            for (int r = 0; r < TH; ++r) {
                const float* prow = Ps + (r + kyl) * PW * A + a;
                const float4* qrow = reinterpret_cast<const float4*>(Qs + r * TW * NB4) + bg;
                float w[KS];
#pragma unroll
                for (int c = 0; c < KS - 1; ++c) w[c] = prow[c * A];
#pragma unroll 1
                for (int x9 = 0; x9 < TW; x9 += KS) {
#pragma unroll
                    for (int u = 0; u < KS; ++u) {
                        const int xx = x9 + u;
                        w[(u + KS - 1) % KS] = prow[(xx + KS - 1) * A];
                        const float4 q = qrow[xx * NBG];
#pragma unroll
                        for (int kx = 0; kx < KS; ++kx) {
                            const float pv = w[(u + kx) % KS];
                            acc[kx][0] = fmaf(pv, q.x, acc[kx][0]); acc[kx][1] = fmaf(pv, q.y, acc[kx][1]);
                            acc[kx][2] = fmaf(pv, q.z, acc[kx][2]); acc[kx][3] = fmaf(pv, q.w, acc[kx][3]);
                        }
                    }
                }
            }
        }
    }
    if (!active) return;
#pragma unroll
    for (int kx = 0; kx < KS; ++kx)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int b = bg * 4 + j;
            if (b < p.NB) atomicAdd(p.dw + ((long long)((ky0 + kyl) * KS + kx) * A + a) * p.NB + b, acc[kx][j]);
        }
}
template <int A, int NBG, int KS, int KYB>
static cudaError_t launch_wgrad_direct(const WgradDirect& p, cudaStream_t s) {
    constexpr int TH = 4, TW = 8 * KS, PW = TW + KS - 1, PH = TH + KYB - 1, NT = KYB * A * NBG, THREADS = (NT + 31) / 32 * 32;
    const size_t smem = (size_t)(TH * TW * NBG * 4 + PH * PW * A) * sizeof(float);
    static SmemAttrCache configured;
    if (cudaError_t e = ensure_dynamic_smem(wgrad_direct_f32_kernel<A, NBG, KS, KYB>, smem, configured)) return e;
    static int per_sm = 0;                                     // a property of the kernel and the architecture, not of the device
    if (!per_sm) {
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, wgrad_direct_f32_kernel<A, NBG, KS, KYB>, THREADS, smem);
        if (e != cudaSuccess) return e;
        if (per_sm < 1) per_sm = 1;
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int tiles_x = ceil_div(p.W, TW), tiles_y = ceil_div(p.H, TH), groups = KS / KYB;
    const long long tiles = (long long)p.B * tiles_x * tiles_y;
    long long bpg = (long long)sms * per_sm / groups;
    if (bpg > tiles) bpg = tiles;
    if (bpg < 1) bpg = 1;
    wgrad_direct_f32_kernel<A, NBG, KS, KYB><<<(unsigned)(bpg * groups), THREADS, smem, s>>>(p, tiles_x, tiles_y, (int)bpg);
    return cudaGetLastError();
}
// thin 9x9 stride-1 'same' layers on images of at least 64 x 64 pixels; false = not handled here
static bool try_wgrad_direct(const WgradF32& p, cudaStream_t s, cudaError_t* e) {
    static const bool off = [] { const char* v = ab_env("RST_WGRAD_DIRECT"); return v && v[0] == '0'; }();
    if (off || p.stride != 1 || p.kh != 9 || p.kw != 9 || p.Hx != p.Hg || p.Wx != p.Wg || p.Hx < 64 || p.Wx < 64) return false;
    WgradDirect d{};
    d.dw = p.dw; d.B = p.B; d.H = p.Hx; d.W = p.Wx; d.pad_t = p.pad_t; d.pad_l = p.pad_l;
    if (!p.transposed) {
        if (p.Co != 32) return false;
        d.P = p.x; d.Q = p.g; d.NB = p.Co; d.p_scale = p.in_scale; d.p_shift = p.in_shift;
        if (p.Ci == 17) *e = launch_wgrad_direct<17, 8, 9, 3>(d, s);
        else if (p.Ci == 18) *e = launch_wgrad_direct<18, 8, 9, 3>(d, s);
        else if (p.Ci == 3) *e = launch_wgrad_direct<3, 8, 9, 9>(d, s);
        else return false;
        return true;
    }
    if (p.Co != 3 || p.Ci != 16 || p.in_scale != 1.f || p.in_shift != 0.f) return false;
    d.P = p.g; d.Q = p.x; d.NB = p.Ci; d.p_scale = 1.f; d.p_shift = 0.f;
    *e = launch_wgrad_direct<3, 4, 9, 9>(d, s);
    return true;
}

cudaError_t launch_wgrad_f32(const WgradF32& p, cudaStream_t s) {
    const long long NP = (long long)p.B * p.Hb * p.Wb;
    if (NP == 0) return cudaSuccess;
    { cudaError_t e = cudaSuccess; if (try_wgrad_direct(p, s, &e)) return e; }
    // row-contiguous tiles work for any stride: the kw taps of one filter row touch kw ADJACENT pixels of the tap-shifted tensor
    static const bool rows_s1_only = [] { const char* e = ab_env("RST_WGRAD_ROWS_S1"); return e && e[0] == '1'; }();
    const int rmode = (p.stride != 1 && rows_s1_only) ? 0 : (p.transposed ? 2 : 1);
    const int m_extent = rmode == 1 ? p.kw * p.Ci : p.Ci, n_extent = rmode == 2 ? p.kw * p.Co : p.Co;
    const bool wide_m = m_extent > 32, wide_n = n_extent > 32;
    const bool big = rmode == 1 && m_extent % 128 == 0 && n_extent % 128 == 0;     // trunk 128 -> 128: 8 x 8 register tiles
    const int tm = big ? 128 : wide_m ? 64 : 32, tn = big ? 128 : wide_n ? 64 : 32;
    const int ci_blocks = ceil_div(m_extent, tm);
    const int groups = rmode ? p.kh : p.kh * p.kw;
    // enough pixel slabs to fill the GPU a few times over, but long enough to amortise the atomics
    const long long tiles = (long long)groups * ci_blocks * ceil_div(n_extent, tn);
    long long splits = (148LL * 8 + tiles - 1) / tiles;
    int pix_per_split = (int)((NP + splits - 1) / splits);
    pix_per_split = (pix_per_split + 31) / 32 * 32;
    if (pix_per_split < 512) pix_per_split = 512;
    dim3 grid((unsigned)(groups * ci_blocks), (unsigned)ceil_div(n_extent, tn), (unsigned)((NP + pix_per_split - 1) / pix_per_split));
#define RST_WGRAD(TM_, TN_)                                                                                              \
    do {                                                                                                                  \
        if (rmode == 1) wgrad_f32_kernel<TM_, TN_, 1><<<grid, 256, 0, s>>>(p, ci_blocks, pix_per_split);                 \
        else if (rmode == 2) wgrad_f32_kernel<TM_, TN_, 2><<<grid, 256, 0, s>>>(p, ci_blocks, pix_per_split);            \
        else wgrad_f32_kernel<TM_, TN_, 0><<<grid, 256, 0, s>>>(p, ci_blocks, pix_per_split);                            \
    } while (0)
    if (big) wgrad_f32_kernel<128, 128, 1><<<grid, 256, 0, s>>>(p, ci_blocks, pix_per_split);
    else if (wide_m && wide_n) RST_WGRAD(64, 64);
    else if (wide_m) RST_WGRAD(64, 32);
    else if (wide_n) RST_WGRAD(32, 64);
    else RST_WGRAD(32, 32);
#undef RST_WGRAD
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------
// normalisation helpers.  stats (G, C, 2) doubles [sum, sumsq] per group g (CIN: g = sample, BatchNorm: one group).
// ---------------------------------------------------------------------------------------------------------------
// per-(n,c) sums -> per-c sums (BatchNorm over the batch)
__global__ void reduce_over_batch_kernel(const double* __restrict__ in, double* __restrict__ out, int B, int C2) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= C2) return;
    double s = 0.0;
    for (int n = 0; n < B; ++n) s += in[(long long)n * C2 + i];
    out[i] = s;
}
cudaError_t launch_reduce_over_batch(const double* in, double* out, int B, int C2, cudaStream_t s) {
    reduce_over_batch_kernel<<<nblk(C2), 256, 0, s>>>(in, out, B, C2);
    return cudaGetLastError();
}

// mean / inv-std per (group, channel); optional affine coefficients y = x*a + b with a = inv*scale, b = bias - mean*a.
// scale/bias: per (group, channel) with strides (CIN: style params) or per channel (BatchNorm: gamma/beta, gstride 0).
// BatchNorm also updates the moving statistics: mm = m*mm + (1-m)*mean, mv = m*mv + (1-m)*var*N/(N-1) (Keras fused BN).
__global__ void norm_finalize_kernel(const double* __restrict__ stats, int G, int C, double count, float eps,
                                     const float* __restrict__ scale, const float* __restrict__ bias, long long gstride,
                                     float* __restrict__ mean, float* __restrict__ inv, float* __restrict__ a,
                                     float* __restrict__ b, float* moving_mean, float* moving_var, float momentum) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= G * C) return;
    const int g = i / C, c = i % C;
    const double m = stats[(long long)i * 2] / count;
    double var = stats[(long long)i * 2 + 1] / count - m * m;
    if (var < 0.0) var = 0.0;
    const float iv = (float)(1.0 / sqrt(var + (double)eps));
    mean[i] = (float)m;
    inv[i] = iv;
    const float sc = scale[g * gstride + c], bi = bias[g * gstride + c];
    a[i] = iv * sc;
    b[i] = bi - (float)m * iv * sc;
    if (moving_mean) {
        moving_mean[c] = momentum * moving_mean[c] + (1.f - momentum) * (float)m;
        const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
        moving_var[c] = momentum * moving_var[c] + (1.f - momentum) * (float)unbiased;
    }
}
cudaError_t launch_norm_finalize(const double* stats, int G, int C, double count, float eps, const float* scale,
                                 const float* bias, long long gstride, float* mean, float* inv, float* a, float* b,
                                 float* moving_mean, float* moving_var, float momentum, cudaStream_t s) {
    norm_finalize_kernel<<<nblk((long long)G * C), 256, 0, s>>>(stats, G, C, count, eps, scale, bias, gstride, mean, inv, a, b,
                                                               moving_mean, moving_var, momentum);
    return cudaGetLastError();
}

// y = act(x*a[g,c] + b[g,c]) (+ residual);  g = sample index when per_sample, else 0
template <int V>
__global__ void affine_act_kernel(const float* __restrict__ x, float* __restrict__ y, const float* __restrict__ a,
                                  const float* __restrict__ b, const float* __restrict__ residual, long long PC, int C,
                                  int per_sample, int act, long long total) {
    const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * V;
    if (i >= total) return;
    const int c = (int)(i % C);
    const long long gidx = per_sample ? (i / PC) * C + c : c;
    float xv[V], rv[V], yv[V];
    ldv<V>(x + i, xv);
    if (residual) ldv<V>(residual + i, rv);
#pragma unroll
    for (int j = 0; j < V; ++j) {
        float v = fmaf(xv[j], a[gidx + j], b[gidx + j]);
        if (act == ACT_RELU) v = fmaxf(v, 0.f);
        else if (act == ACT_SIGMOID) v = 1.f / (1.f + expf(-v));
        else if (act == ACT_HSIGMOID) v = fminf(fmaxf(v + 3.f, 0.f), 6.f) * (1.f / 6.f);
        else if (act == ACT_HSWISH) v = v * fminf(fmaxf(v + 3.f, 0.f), 6.f) * (1.f / 6.f);
        if (residual) v += rv[j];
        yv[j] = v;
    }
    stv<V>(y + i, yv);
}
cudaError_t launch_affine_act(const float* x, float* y, const float* a, const float* b, const float* residual, int B, long long P,
                              int C, int per_sample, int act, cudaStream_t s) {
    const long long total = (long long)B * P * C;
    if (total == 0) return cudaSuccess;
    if (vec4_ok(total, C, {x, y, residual}))
        affine_act_kernel<4><<<nblk(total / 4), 256, 0, s>>>(x, y, a, b, residual, P * C, C, per_sample, act, total);
    else
        affine_act_kernel<1><<<nblk(total), 256, 0, s>>>(x, y, a, b, residual, P * C, C, per_sample, act, total);
    return cudaGetLastError();
}

// g *= act'(out)   (ReLU: out > 0;  sigmoid: out*(1-out))
template <int V>
__global__ void act_bwd_kernel(float* __restrict__ g, const float* __restrict__ out, int act, long long n) {
    const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * V;
    if (i >= n) return;
    float ov[V], gv[V];
    ldv<V>(out + i, ov);
    ldv<V>(g + i, gv);
#pragma unroll
    for (int j = 0; j < V; ++j) {
        const float o = ov[j];
        if (act == ACT_RELU) { if (!(o > 0.f)) gv[j] = 0.f; }
        else if (act == ACT_SIGMOID) gv[j] *= o * (1.f - o);
        else if (act == ACT_HSIGMOID) gv[j] = (o > 0.f && o < 1.f) ? gv[j] * (1.f / 6.f) : 0.f;
    }
    stv<V>(g + i, gv);
}
cudaError_t launch_act_bwd(float* g, const float* out, int act, long long n, cudaStream_t s) {
    if (n == 0 || act == ACT_NONE) return cudaSuccess;
    if (vec4_ok(n, 4, {g, out})) act_bwd_kernel<4><<<nblk(n / 4), 256, 0, s>>>(g, out, act, n);
    else act_bwd_kernel<1><<<nblk(n), 256, 0, s>>>(g, out, act, n);
    return cudaGetLastError();
}

// g *= act'(u), u = x*a[g,c] + b[g,c] recomputed from the saved pre-normalisation tensor
template <int V>
__global__ void affine_act_bwd_kernel(float* __restrict__ g, const float* __restrict__ x, const float* __restrict__ a,
                                      const float* __restrict__ b, long long PC, int C, int per_sample, int act, long long total) {
    const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * V;
    if (i >= total) return;
    const int c = (int)(i % C);
    const long long gidx = per_sample ? (i / PC) * C + c : c;
    float xv[V], gv[V];
    ldv<V>(x + i, xv);
    ldv<V>(g + i, gv);
#pragma unroll
    for (int j = 0; j < V; ++j) {
        const float u = fmaf(xv[j], a[gidx + j], b[gidx + j]);
        float d = 1.f;
        if (act == ACT_RELU) d = u > 0.f ? 1.f : 0.f;
        else if (act == ACT_SIGMOID) { const float sg = 1.f / (1.f + expf(-u)); d = sg * (1.f - sg); }
        else if (act == ACT_HSIGMOID) d = (u > -3.f && u < 3.f) ? 1.f / 6.f : 0.f;
        else if (act == ACT_HSWISH) d = u <= -3.f ? 0.f : (u >= 3.f ? 1.f : (2.f * u + 3.f) / 6.f);
        gv[j] *= d;
    }
    stv<V>(g + i, gv);
}
cudaError_t launch_affine_act_bwd(float* g, const float* x, const float* a, const float* b, int B, long long P, int C,
                                  int per_sample, int act, cudaStream_t s) {
    const long long total = (long long)B * P * C;
    if (total == 0 || act == ACT_NONE) return cudaSuccess;
    if (vec4_ok(total, C, {g, x}))
        affine_act_bwd_kernel<4><<<nblk(total / 4), 256, 0, s>>>(g, x, a, b, P * C, C, per_sample, act, total);
    else
        affine_act_bwd_kernel<1><<<nblk(total), 256, 0, s>>>(g, x, a, b, P * C, C, per_sample, act, total);
    return cudaGetLastError();
}

// per-(n,c): r[.,0] = sum_p g, r[.,1] = sum_p g * xhat, xhat = (x - mean)*inv   (mean/inv indexed per sample or per channel)
__global__ void norm_bwd_reduce_kernel(const float* __restrict__ g, const float* __restrict__ x, const float* __restrict__ mean,
                                       const float* __restrict__ inv, double* __restrict__ r, int P, int C, int per_sample,
                                       int pix_per_block) {
    extern __shared__ double sm[];
    const int n = blockIdx.y;
    const int p0 = blockIdx.x * pix_per_block, p1 = min(P, p0 + pix_per_block);
    const long long base = (long long)n * P * C;
    const int T = blockDim.x;
    if (C <= T) {
        const int G = T / C, c = threadIdx.x % C, gi = threadIdx.x / C;
        double s1 = 0.0, s2 = 0.0;
        if (gi < G) {
            const float m = mean ? mean[per_sample ? n * C + c : c] : 0.f, iv = inv ? inv[per_sample ? n * C + c : c] : 1.f;
            double t1 = 0.0, t2 = 0.0;
            int pp = p0 + gi;
            for (; pp + G < p1; pp += 2 * G) {                      // two independent load pairs / accumulator chains in flight
                const float ga = g[base + (long long)pp * C + c], gb = g[base + (long long)(pp + G) * C + c];
                const float xa = (x[base + (long long)pp * C + c] - m) * iv, xb = (x[base + (long long)(pp + G) * C + c] - m) * iv;
                s1 += ga; s2 += (double)ga * xa;
                t1 += gb; t2 += (double)gb * xb;
            }
            for (; pp < p1; pp += G) {
                const float gv = g[base + (long long)pp * C + c];
                const float xh = (x[base + (long long)pp * C + c] - m) * iv;
                s1 += gv;
                s2 += (double)gv * xh;
            }
            s1 += t1; s2 += t2;
        }
        sm[threadIdx.x] = s1;
        sm[T + threadIdx.x] = s2;
        __syncthreads();
        if (gi == 0) {
            for (int j = 1; j < G; ++j) { s1 += sm[j * C + c]; s2 += sm[T + j * C + c]; }
            atomicAdd(&r[((long long)n * C + c) * 2], s1);
            atomicAdd(&r[((long long)n * C + c) * 2 + 1], s2);
        }
    } else {
        for (int c = threadIdx.x; c < C; c += T) {
            const float m = mean ? mean[per_sample ? n * C + c : c] : 0.f, iv = inv ? inv[per_sample ? n * C + c : c] : 1.f;
            double s1 = 0.0, s2 = 0.0;
            for (int pp = p0; pp < p1; ++pp) {
                const float gv = g[base + (long long)pp * C + c];
                s1 += gv;
                s2 += (double)gv * ((x[base + (long long)pp * C + c] - m) * iv);
            }
            atomicAdd(&r[((long long)n * C + c) * 2], s1);
            atomicAdd(&r[((long long)n * C + c) * 2 + 1], s2);
        }
    }
}
cudaError_t launch_norm_bwd_reduce(const float* g, const float* x, const float* mean, const float* inv, double* r, int B, int P,
                                   int C, int per_sample, cudaStream_t s) {
    if (B == 0 || P == 0) return cudaSuccess;
    const int threads = C <= 256 ? C * (256 / C) : 256;
    // enough CTAs to fill the GPU even for the small, wide tensors of the predictor (a CTA walks its pixels serially)
    int pix_per_block = 2048;
    while (pix_per_block > 64 && (long long)ceil_div(P, pix_per_block) * B < 148 * 8) pix_per_block /= 2;
    dim3 grid((unsigned)ceil_div(P, pix_per_block), (unsigned)B);
    norm_bwd_reduce_kernel<<<grid, threads, 2 * threads * sizeof(double), s>>>(g, x, mean, inv, r, P, C, per_sample, pix_per_block);
    return cudaGetLastError();
}

// gx = a[g,c] * (g - r1/N - xhat * r2/N);   r indexed per sample (CIN) or per channel (BatchNorm, already reduced over n)
template <int V>
__global__ void norm_bwd_apply_kernel(const float* __restrict__ g, const float* __restrict__ x, const float* __restrict__ mean,
                                      const float* __restrict__ inv, const float* __restrict__ a, const double* __restrict__ r,
                                      float* __restrict__ gx, long long PC, int C, int per_sample, double invN, int accumulate,
                                      long long total) {
    const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * V;
    if (i >= total) return;
    const int c = (int)(i % C);
    const long long g0 = per_sample ? (i / PC) * C + c : c;
    float gv[V], xv[V], ov[V];
    ldv<V>(g + i, gv);
    ldv<V>(x + i, xv);
    if (accumulate) ldv<V>(gx + i, ov);
#pragma unroll
    for (int j = 0; j < V; ++j) {
        const long long gi = g0 + j;
        const float xh = (xv[j] - mean[gi]) * inv[gi];
        const float m1 = (float)(r[gi * 2] * invN), m2 = (float)(r[gi * 2 + 1] * invN);
        const float v = a[gi] * (gv[j] - m1 - xh * m2);
        ov[j] = accumulate ? ov[j] + v : v;
    }
    stv<V>(gx + i, ov);
}
cudaError_t launch_norm_bwd_apply(const float* g, const float* x, const float* mean, const float* inv, const float* a,
                                  const double* r, float* gx, int B, long long P, int C, int per_sample, double count,
                                  int accumulate, cudaStream_t s) {
    const long long total = (long long)B * P * C;
    if (total == 0) return cudaSuccess;
    if (vec4_ok(total, C, {g, x, gx}))
        norm_bwd_apply_kernel<4><<<nblk(total / 4), 256, 0, s>>>(g, x, mean, inv, a, r, gx, P * C, C, per_sample, 1.0 / count, accumulate, total);
    else
        norm_bwd_apply_kernel<1><<<nblk(total), 256, 0, s>>>(g, x, mean, inv, a, r, gx, P * C, C, per_sample, 1.0 / count, accumulate, total);
    return cudaGetLastError();
}

// CIN parameter gradients: d scale[n,c] = r[n,c,1], d bias[n,c] = r[n,c,0], accumulated into the (B, Ptotal) style-parameter
// gradient at columns [off, off+C) and [off+C, off+2C)
__global__ void cin_param_grad_kernel(const double* __restrict__ r, float* __restrict__ pg, int B, int C, long long ptotal, int off) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * C) return;
    const int n = i / C, c = i % C;
    pg[n * ptotal + off + c] += (float)r[(long long)i * 2 + 1];
    pg[n * ptotal + off + C + c] += (float)r[(long long)i * 2];
}
cudaError_t launch_cin_param_grad(const double* r, float* pg, int B, int C, long long ptotal, int off, cudaStream_t s) {
    if (B * C == 0) return cudaSuccess;
    cin_param_grad_kernel<<<nblk((long long)B * C), 256, 0, s>>>(r, pg, B, C, ptotal, off);
    return cudaGetLastError();
}

// out[i] (+)= (float) in[i*stride + offset]   (pull sums out of the double reduction buffers)
__global__ void gather_f64_kernel(const double* __restrict__ in, float* __restrict__ out, int n, int stride, int offset, int accumulate) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float v = (float)in[(long long)i * stride + offset];
    out[i] = accumulate ? out[i] + v : v;
}
cudaError_t launch_gather_f64(const double* in, float* out, int n, int stride, int offset, int accumulate, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    gather_f64_kernel<<<nblk(n), 256, 0, s>>>(in, out, n, stride, offset, accumulate);
    return cudaGetLastError();
}

// a[i] += b[i]
__global__ void add_inplace_kernel(float* __restrict__ a, const float* __restrict__ b, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] += b[i];
}
cudaError_t launch_add_inplace(float* a, const float* b, long long n, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    add_inplace_kernel<<<nblk(n), 256, 0, s>>>(a, b, n);
    return cudaGetLastError();
}
// global-average-pool backward: gx[n,p,c] = g[n,c] / P
__global__ void gap_bwd_kernel(const float* __restrict__ g, float* __restrict__ gx, long long PC, int C, float invP, int accumulate,
                               long long total) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const float v = g[(i / PC) * C + (i % C)] * invP;
    gx[i] = accumulate ? gx[i] + v : v;
}
cudaError_t launch_gap_bwd(const float* g, float* gx, int B, long long P, int C, int accumulate, cudaStream_t s) {
    const long long total = (long long)B * P * C;
    if (total == 0) return cudaSuccess;
    gap_bwd_kernel<<<nblk(total), 256, 0, s>>>(g, gx, P * C, C, 1.f / (float)P, accumulate, total);
    return cudaGetLastError();
}
// squeeze-excite multiply backward w.r.t. the feature map: gx (+)= g * z[n,c]
__global__ void scale_channels_bwd_kernel(const float* __restrict__ g, const float* __restrict__ z, float* __restrict__ gx,
                                          long long PC, int C, int accumulate, long long total) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const float v = g[i] * z[(i / PC) * C + (i % C)];
    gx[i] = accumulate ? gx[i] + v : v;
}
cudaError_t launch_scale_channels_bwd(const float* g, const float* z, float* gx, int B, long long P, int C, int accumulate,
                                      cudaStream_t s) {
    const long long total = (long long)B * P * C;
    if (total == 0) return cudaSuccess;
    scale_channels_bwd_kernel<<<nblk(total), 256, 0, s>>>(g, z, gx, P * C, C, accumulate, total);
    return cudaGetLastError();
}

// depthwise conv input gradient: gx[n,iy,ix,c] (+)= sum_k g[n,(iy+pt-ky)/s,(ix+pl-kx)/s,c] * w[k,c]
__global__ void depthwise_dgrad_kernel(const DepthwiseF32 p, const float* __restrict__ g, float* __restrict__ gx, int accumulate,
                                       long long total) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int c = (int)(idx % p.C);
    long long t = idx / p.C;
    const int ix = (int)(t % p.Wi);
    t /= p.Wi;
    const int iy = (int)(t % p.Hi);
    const int n = (int)(t / p.Hi);
    float acc = 0.f;
    for (int ky = 0; ky < p.k; ++ky) {
        const int ty = iy + p.pad_t - ky;
        if (ty < 0 || ty % p.stride) continue;
        const int oy = ty / p.stride;
        if (oy >= p.Ho) continue;
        for (int kx = 0; kx < p.k; ++kx) {
            const int tx = ix + p.pad_l - kx;
            if (tx < 0 || tx % p.stride) continue;
            const int ox = tx / p.stride;
            if (ox >= p.Wo) continue;
            acc = fmaf(__ldg(g + (((long long)n * p.Ho + oy) * p.Wo + ox) * p.C + c), __ldg(p.w + (ky * p.k + kx) * p.C + c), acc);
        }
    }
    gx[idx] = accumulate ? gx[idx] + acc : acc;
}
cudaError_t launch_depthwise_dgrad(const DepthwiseF32& p, const float* g, float* gx, int accumulate, cudaStream_t s) {
    const long long total = (long long)p.B * p.Hi * p.Wi * p.C;
    if (total == 0) return cudaSuccess;
    depthwise_dgrad_kernel<<<nblk(total), 256, 0, s>>>(p, g, gx, accumulate, total);
    return cudaGetLastError();
}
// depthwise conv weight gradient: dw[k,c] += sum_{n,oy,ox} x[n,oy*s-pt+ky,ox*s-pl+kx,c] * g[n,oy,ox,c]   (dw zeroed by the caller)
__global__ void depthwise_wgrad_kernel(const DepthwiseF32 p, const float* __restrict__ g, float* __restrict__ dw, int pix_per_block) {
    const int tap = blockIdx.x, ky = tap / p.k, kx = tap - ky * p.k;
    const long long NP = (long long)p.B * p.Ho * p.Wo;
    const long long p0 = (long long)blockIdx.y * pix_per_block, p1 = min(NP, p0 + pix_per_block);
    for (int c = threadIdx.x; c < p.C; c += blockDim.x) {
        float acc = 0.f;
        for (long long pix = p0; pix < p1; ++pix) {
            const int ox = (int)(pix % p.Wo);
            const long long t = pix / p.Wo;
            const int oy = (int)(t % p.Ho);
            const int n = (int)(t / p.Ho);
            const int iy = oy * p.stride - p.pad_t + ky, ix = ox * p.stride - p.pad_l + kx;
            if (iy < 0 || iy >= p.Hi || ix < 0 || ix >= p.Wi) continue;
            acc = fmaf(__ldg(p.x + (((long long)n * p.Hi + iy) * p.Wi + ix) * p.C + c), __ldg(g + pix * p.C + c), acc);
        }
        atomicAdd(dw + tap * p.C + c, acc);
    }
}
cudaError_t launch_depthwise_wgrad(const DepthwiseF32& p, const float* g, float* dw, cudaStream_t s) {
    const long long NP = (long long)p.B * p.Ho * p.Wo;
    if (NP == 0) return cudaSuccess;
    const int pix_per_block = 64;
    dim3 grid((unsigned)(p.k * p.k), (unsigned)((NP + pix_per_block - 1) / pix_per_block));
    depthwise_wgrad_kernel<<<grid, 128, 0, s>>>(p, g, dw, pix_per_block);
    return cudaGetLastError();
}

// Keras RMSprop (TF 2.9, momentum 0, not centred): rms = rho*rms + (1-rho)*g^2;  w -= lr * g / (sqrt(rms) + eps)
__global__ void rmsprop_kernel(float* __restrict__ w, const float* __restrict__ g, float* __restrict__ rms, float lr, float rho,
                               float eps, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float gv = g[i];
    const float r = rho * rms[i] + (1.f - rho) * gv * gv;
    rms[i] = r;
    w[i] -= lr * gv / (sqrtf(r) + eps);
}
cudaError_t launch_rmsprop(float* w, const float* g, float* rms, float lr, float rho, float eps, long long n, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    rmsprop_kernel<<<nblk(n), 256, 0, s>>>(w, g, rms, lr, rho, eps, n);
    return cudaGetLastError();
}

}  // namespace rst
