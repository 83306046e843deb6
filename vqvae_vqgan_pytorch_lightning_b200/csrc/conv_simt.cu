// fp32 SIMT implicit-GEMM convolution (strict-parity path; also serves the 3-channel edge layers of the bf16
// path).  NHWC activations, packed weights wp[(kh,kw,ci)][co].
//
//   forward : C[m = (n,oh,ow)][co] = sum_k A[m][k = (kh,kw,ci)] * wp[k][co]      A gathered on the fly (no im2col)
//   wgrad   : dwp[m = (kh,kw,ci)][co] = sum_p A'[m][p = (n,oh,ow)] * dy[p][co]    split over p, fp32 atomics
//
// Block tile 128 x 128 x 16, 256 threads, 8 x 8 register micro-tile per thread, register-staged double
// buffering (global loads of tile t+1 are in flight while tile t is multiplied).  The tensor-core path for
// the wide layers is conv_tc.cu; this kernel's roof is the fp32 FMA pipe (148 SMs x 128 FMA/clk).
#include "common.cuh"

namespace {

constexpr int BM = 128, BN = 128, BK = 16, PADM = 4;

struct ConvGeom {
    int N, H, W, Ci, Co, KH, KW, pad, stride, OH, OW;
    int up;      // > 1: the input is (virtually) zero-upsampled by `up` -- transposed-conv gather for strided dgrad
    int64_t M;   // N*OH*OW
    int K;       // KH*KW*Ci
};

__device__ __forceinline__ float apply_act(float v, int act, float alpha) {
    if (act == VQB_ACT_TANH) return tanhf(v);
    if (act == VQB_ACT_SILU) return silu_f(v);
    if (act == VQB_ACT_LRELU) return v > 0.f ? v : v * alpha;
    if (act == VQB_ACT_RELU) return fmaxf(v, 0.f);
    return v;
}

__device__ __forceinline__ void mma_tile(const float (*As)[BM + PADM], const float (*Bs)[BN], float (&acc)[8][8], int ty, int tx) {
#pragma unroll
    for (int k = 0; k < BK; ++k) {
        float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
        float4 a1 = *reinterpret_cast<const float4*>(&As[k][64 + ty * 4]);
        float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
        float4 b1 = *reinterpret_cast<const float4*>(&Bs[k][64 + tx * 4]);
        float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
}

// ---------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------
template <typename TIn, typename TOut>
__global__ void __launch_bounds__(256, 2)
conv_fwd_simt_kernel(const TIn* __restrict__ x, const float* __restrict__ wp, const float* __restrict__ bias,
                     const TOut* __restrict__ residual, TOut* __restrict__ y, ConvGeom g, int act, float alpha, float gain) {
    __shared__ __align__(16) float As[2][BK][BM + PADM];
    __shared__ __align__(16) float Bs[2][BK][BN];
    const int tid = threadIdx.x;
    const int ty = tid >> 4, tx = tid & 15;
    const int64_t m0 = (int64_t)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const bool vecA = (g.Ci % 4 == 0);
    const bool vecB = (g.Co % 4 == 0);

    // A loader: rows ar, ar+64; k offset akc..akc+3
    const int ar = tid >> 2, akc = (tid & 3) * 4;
    int a_ih0[2], a_iw0[2];
    int64_t a_base[2];
    bool a_ok[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        int64_t m = m0 + ar + r * 64;
        a_ok[r] = m < g.M;
        int64_t mm = a_ok[r] ? m : 0;
        int ow = (int)(mm % g.OW); int64_t t = mm / g.OW; int oh = (int)(t % g.OH); int n = (int)(t / g.OH);
        a_ih0[r] = oh * g.stride - g.pad;
        a_iw0[r] = ow * g.stride - g.pad;
        a_base[r] = (int64_t)n * g.H * g.W * g.Ci;
    }
    // B loader: rows bk, bk+8; n offset bn..bn+3
    const int bk = tid >> 5, bn = (tid & 31) * 4;

    float a_reg[2][4], b_reg[2][4];
    auto load_tiles = [&](int kt) {
        const int k0 = kt * BK;
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            int k = k0 + akc;
            if (vecA) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (a_ok[r] && k < g.K) {
                    int tap = k / g.Ci, ci = k - tap * g.Ci;
                    int kh = tap / g.KW, kw = tap - kh * g.KW;
                    int ih = a_ih0[r] + kh, iw = a_iw0[r] + kw;
                    bool ok = ih >= 0 && iw >= 0;
                    if (g.up > 1) { ok = ok && (ih % g.up == 0) && (iw % g.up == 0); ih /= g.up; iw /= g.up; }
                    if (ok && ih < g.H && iw < g.W)
                        v = ld4(x + a_base[r] + ((int64_t)ih * g.W + iw) * g.Ci + ci);
                }
                a_reg[r][0] = v.x; a_reg[r][1] = v.y; a_reg[r][2] = v.z; a_reg[r][3] = v.w;
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float v = 0.f;
                    int kk = k + j;
                    if (a_ok[r] && kk < g.K) {
                        int tap = kk / g.Ci, ci = kk - tap * g.Ci;
                        int kh = tap / g.KW, kw = tap - kh * g.KW;
                        int ih = a_ih0[r] + kh, iw = a_iw0[r] + kw;
                        bool ok = ih >= 0 && iw >= 0;
                        if (g.up > 1) { ok = ok && (ih % g.up == 0) && (iw % g.up == 0); ih /= g.up; iw /= g.up; }
                        if (ok && ih < g.H && iw < g.W)
                            v = ld1(x + a_base[r] + ((int64_t)ih * g.W + iw) * g.Ci + ci);
                    }
                    a_reg[r][j] = v;
                }
            }
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            int k = k0 + bk + r * 8;
            int n = n0 + bn;
            if (vecB) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (k < g.K && n < g.Co) v = *reinterpret_cast<const float4*>(wp + (int64_t)k * g.Co + n);
                b_reg[r][0] = v.x; b_reg[r][1] = v.y; b_reg[r][2] = v.z; b_reg[r][3] = v.w;
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) b_reg[r][j] = (k < g.K && n + j < g.Co) ? wp[(int64_t)k * g.Co + n + j] : 0.f;
            }
        }
    };
    auto store_tiles = [&](int buf) {
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int j = 0; j < 4; ++j) As[buf][akc + j][ar + r * 64] = a_reg[r][j];
#pragma unroll
        for (int r = 0; r < 2; ++r)
            *reinterpret_cast<float4*>(&Bs[buf][bk + r * 8][bn]) = make_float4(b_reg[r][0], b_reg[r][1], b_reg[r][2], b_reg[r][3]);
    };

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    const int KT = (g.K + BK - 1) / BK;
    load_tiles(0);
    store_tiles(0);
    __syncthreads();
    for (int kt = 0; kt < KT; ++kt) {
        const int cur = kt & 1;
        if (kt + 1 < KT) load_tiles(kt + 1);
        mma_tile(As[cur], Bs[cur], acc, ty, tx);
        if (kt + 1 < KT) store_tiles(cur ^ 1);
        __syncthreads();
    }

    // epilogue
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        int64_t m = m0 + ((i < 4) ? (ty * 4 + i) : (64 + ty * 4 + i - 4));
        if (m >= g.M) continue;
#pragma unroll
        for (int jh = 0; jh < 2; ++jh) {
            int n = n0 + jh * 64 + tx * 4;
            if (n >= g.Co) continue;
            float v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float t = acc[i][jh * 4 + j];
                if (bias && n + j < g.Co) t += bias[n + j];
                t = apply_act(t, act, alpha) * gain;
                v[j] = t;
            }
            int64_t off = m * g.Co + n;
            if (vecB) {
                if (residual) { float4 r = ld4(residual + off); v[0] += r.x; v[1] += r.y; v[2] += r.z; v[3] += r.w; }
                st4(y + off, make_float4(v[0], v[1], v[2], v[3]));
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (n + j < g.Co) {
                        float t = v[j];
                        if (residual) t += ld1(residual + off + j);
                        st1(y + off + j, t);
                    }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// weight gradient
// ---------------------------------------------------------------------------------------------------
template <typename TIn, typename TG>
__global__ void __launch_bounds__(256, 2)
conv_wgrad_simt_kernel(const TIn* __restrict__ x, const TG* __restrict__ dy, float* __restrict__ dwp, ConvGeom g,
                       int64_t p_chunk) {
    __shared__ __align__(16) float As[2][BK][BM + PADM];
    __shared__ __align__(16) float Bs[2][BK][BN];
    const int tid = threadIdx.x;
    const int ty = tid >> 4, tx = tid & 15;
    const int m0 = blockIdx.x * BM;      // over K = (kh,kw,ci)
    const int n0 = blockIdx.y * BN;      // over Co
    const int64_t pbeg = (int64_t)blockIdx.z * p_chunk;
    int64_t pend = pbeg + p_chunk; if (pend > g.M) pend = g.M;
    const bool vecA = (g.Ci % 4 == 0);
    const bool vecB = (g.Co % 4 == 0);

    // loaders: p rows lk, lk+8 ; m / n offset lc..lc+3
    const int lk = tid >> 5, lc = (tid & 31) * 4;
    // per-thread decomposition of its A columns m = m0+lc+j -> (tap, ci)
    int a_kh[4], a_kw[4], a_ci[4];
    bool a_mok[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        int m = m0 + lc + j;
        a_mok[j] = m < g.K;
        int mm = a_mok[j] ? m : 0;
        int tap = mm / g.Ci;
        a_ci[j] = mm - tap * g.Ci;
        a_kh[j] = tap / g.KW;
        a_kw[j] = tap - a_kh[j] * g.KW;
    }

    float a_reg[2][4], b_reg[2][4];
    auto load_tiles = [&](int64_t p0) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            int64_t p = p0 + lk + r * 8;
            bool pok = p < pend;
            int64_t pp = pok ? p : 0;
            int ow = (int)(pp % g.OW); int64_t t = pp / g.OW; int oh = (int)(t % g.OH); int n = (int)(t / g.OH);
            int ih0 = oh * g.stride - g.pad, iw0 = ow * g.stride - g.pad;
            const TIn* xb = x + (int64_t)n * g.H * g.W * g.Ci;
            if (vecA) {   // the 4 columns share one tap (Ci % 4 == 0)
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                int ih = ih0 + a_kh[0], iw = iw0 + a_kw[0];
                if (pok && a_mok[0] && ih >= 0 && ih < g.H && iw >= 0 && iw < g.W)
                    v = ld4(xb + ((int64_t)ih * g.W + iw) * g.Ci + a_ci[0]);
                a_reg[r][0] = v.x; a_reg[r][1] = v.y; a_reg[r][2] = v.z; a_reg[r][3] = v.w;
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float v = 0.f;
                    int ih = ih0 + a_kh[j], iw = iw0 + a_kw[j];
                    if (pok && a_mok[j] && ih >= 0 && ih < g.H && iw >= 0 && iw < g.W)
                        v = ld1(xb + ((int64_t)ih * g.W + iw) * g.Ci + a_ci[j]);
                    a_reg[r][j] = v;
                }
            }
            int n_ = n0 + lc;
            if (vecB) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (pok && n_ < g.Co) v = ld4(dy + p * g.Co + n_);
                b_reg[r][0] = v.x; b_reg[r][1] = v.y; b_reg[r][2] = v.z; b_reg[r][3] = v.w;
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) b_reg[r][j] = (pok && n_ + j < g.Co) ? ld1(dy + p * g.Co + n_ + j) : 0.f;
            }
        }
    };
    auto store_tiles = [&](int buf) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            *reinterpret_cast<float4*>(&As[buf][lk + r * 8][lc]) = make_float4(a_reg[r][0], a_reg[r][1], a_reg[r][2], a_reg[r][3]);
            *reinterpret_cast<float4*>(&Bs[buf][lk + r * 8][lc]) = make_float4(b_reg[r][0], b_reg[r][1], b_reg[r][2], b_reg[r][3]);
        }
    };

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    if (pbeg >= pend) return;
    const int64_t PT = (pend - pbeg + BK - 1) / BK;
    load_tiles(pbeg);
    store_tiles(0);
    __syncthreads();
    for (int64_t pt = 0; pt < PT; ++pt) {
        const int cur = (int)(pt & 1);
        if (pt + 1 < PT) load_tiles(pbeg + (pt + 1) * BK);
        mma_tile(As[cur], Bs[cur], acc, ty, tx);
        if (pt + 1 < PT) store_tiles(cur ^ 1);
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        int m = m0 + ((i < 4) ? (ty * 4 + i) : (64 + ty * 4 + i - 4));
        if (m >= g.K) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            int n = n0 + ((j < 4) ? (tx * 4 + j) : (64 + tx * 4 + j - 4));
            if (n < g.Co) atomicAdd(dwp + (int64_t)m * g.Co + n, acc[i][j]);
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// weight gradient of a narrow head (Co <= 4, stride 1, 3x3): decoder.conv_out 128 -> 3 at full resolution.
// The GEMM view (N = 3) wastes a 128-wide tile; here thread <-> input channel keeps the 9 x Co partial sums in
// registers, streams x once (coalesced over channels) and reads the 3x3 window of dy from a zero-padded
// shared-memory copy of three dy rows (broadcast reads).  HBM-bound: x is read exactly once.
// ---------------------------------------------------------------------------------------------------
template <typename TIn, typename TG>
__global__ void __launch_bounds__(128)
conv_wgrad_narrow_kernel(const TIn* __restrict__ x, const TG* __restrict__ dy, float* __restrict__ dwp, int N, int H, int W,
                         int Ci, int Co) {
    extern __shared__ float4 dys[];                 // [3][W + 2] (co padded to 4), zero borders
    const int ci = blockIdx.y * 128 + threadIdx.x;
    const bool ci_ok = ci < Ci;
    float acc[9][4];
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[t][c] = 0.f;
    const int WP = W + 2;
    for (int row = blockIdx.x; row < N * H; row += gridDim.x) {
        const int n = row / H, h = row - n * H;
        __syncthreads();
        for (int i = threadIdx.x; i < 3 * WP; i += 128) {
            int r = i / WP, wp = i - r * WP;
            int hh = h + 1 - r;                      // tap kh pairs input row h with output row h - kh + 1  (kh = r)
            int ww = wp - 1;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (hh >= 0 && hh < H && ww >= 0 && ww < W) {
                const TG* src = dy + (((int64_t)n * H + hh) * W + ww) * Co;
                v.x = ld1(src);
                if (Co > 1) v.y = ld1(src + 1);
                if (Co > 2) v.z = ld1(src + 2);
                if (Co > 3) v.w = ld1(src + 3);
            }
            dys[i] = v;
        }
        __syncthreads();
        if (ci_ok) {
            const TIn* xr = x + (((int64_t)n * H + h) * W) * Ci + ci;
            for (int w = 0; w < W; ++w) {
                const float xv = ld1(xr + (int64_t)w * Ci);
#pragma unroll
                for (int kh = 0; kh < 3; ++kh)
#pragma unroll
                    for (int kw = 0; kw < 3; ++kw) {
                        // input pixel (h, w) meets output pixel (h - kh + 1, w - kw + 1): padded column index w - kw + 2
                        const float4 g = dys[kh * WP + (w - kw + 2)];
                        acc[kh * 3 + kw][0] = fmaf(xv, g.x, acc[kh * 3 + kw][0]);
                        acc[kh * 3 + kw][1] = fmaf(xv, g.y, acc[kh * 3 + kw][1]);
                        acc[kh * 3 + kw][2] = fmaf(xv, g.z, acc[kh * 3 + kw][2]);
                        acc[kh * 3 + kw][3] = fmaf(xv, g.w, acc[kh * 3 + kw][3]);
                    }
            }
        }
    }
    if (ci_ok) {
#pragma unroll
        for (int t = 0; t < 9; ++t)
#pragma unroll
            for (int c = 0; c < 4; ++c)
                if (c < Co) atomicAdd(dwp + ((int64_t)t * Ci + ci) * Co + c, acc[t][c]);
    }
}

// ---------------------------------------------------------------------------------------------------
// narrow INPUT (Ci <= 4, 3x3, stride 1): encoder.conv_in 3 -> 128 and the dgrad of decoder.conv_out.
// thread <-> output channel; the 9 x Ci weights of that channel live in registers; three zero-padded input rows
// are staged in shared memory and read as broadcasts.  HBM-bound: y is written once, x read once.
// ---------------------------------------------------------------------------------------------------
template <typename TIn, typename TOut>
__global__ void __launch_bounds__(128)
conv_fwd_narrow_ci_kernel(const TIn* __restrict__ x, const float* __restrict__ wp, const float* __restrict__ bias,
                          TOut* __restrict__ y, int N, int H, int W, int Ci, int Co) {
    extern __shared__ float4 xs[];                  // [3][W + 2]
    const int co = blockIdx.y * 128 + threadIdx.x;
    const bool ok = co < Co;
    float wr[9][4];
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int c = 0; c < 4; ++c) wr[t][c] = (ok && c < Ci) ? wp[((int64_t)t * Ci + c) * Co + co] : 0.f;
    const float b = (ok && bias) ? bias[co] : 0.f;
    const int WP = W + 2;
    for (int row = blockIdx.x; row < N * H; row += gridDim.x) {
        const int n = row / H, h = row - n * H;
        __syncthreads();
        for (int i = threadIdx.x; i < 3 * WP; i += 128) {
            int r = i / WP, wp_ = i - r * WP;
            int hh = h + r - 1, ww = wp_ - 1;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (hh >= 0 && hh < H && ww >= 0 && ww < W) {
                const TIn* src = x + (((int64_t)n * H + hh) * W + ww) * Ci;
                v.x = ld1(src);
                if (Ci > 1) v.y = ld1(src + 1);
                if (Ci > 2) v.z = ld1(src + 2);
                if (Ci > 3) v.w = ld1(src + 3);
            }
            xs[i] = v;
        }
        __syncthreads();
        if (ok) {
            TOut* yr = y + (((int64_t)n * H + h) * W) * Co + co;
            for (int w = 0; w < W; ++w) {
                float acc = b;
#pragma unroll
                for (int kh = 0; kh < 3; ++kh)
#pragma unroll
                    for (int kw = 0; kw < 3; ++kw) {
                        const float4 v = xs[kh * WP + w + kw];
                        acc = fmaf(v.x, wr[kh * 3 + kw][0], acc);
                        acc = fmaf(v.y, wr[kh * 3 + kw][1], acc);
                        acc = fmaf(v.z, wr[kh * 3 + kw][2], acc);
                        acc = fmaf(v.w, wr[kh * 3 + kw][3], acc);
                    }
                st1(yr + (int64_t)w * Co, acc);
            }
        }
    }
}

template <typename TIn, typename TG>
__global__ void __launch_bounds__(128)
conv_wgrad_narrow_ci_kernel(const TIn* __restrict__ x, const TG* __restrict__ dy, float* __restrict__ dwp, int N, int H, int W,
                            int Ci, int Co) {
    extern __shared__ float4 xs[];                  // [3][W + 2]
    const int co = blockIdx.y * 128 + threadIdx.x;
    const bool ok = co < Co;
    float acc[9][4];
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[t][c] = 0.f;
    const int WP = W + 2;
    for (int row = blockIdx.x; row < N * H; row += gridDim.x) {
        const int n = row / H, h = row - n * H;
        __syncthreads();
        for (int i = threadIdx.x; i < 3 * WP; i += 128) {
            int r = i / WP, wp_ = i - r * WP;
            int hh = h + r - 1, ww = wp_ - 1;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (hh >= 0 && hh < H && ww >= 0 && ww < W) {
                const TIn* src = x + (((int64_t)n * H + hh) * W + ww) * Ci;
                v.x = ld1(src);
                if (Ci > 1) v.y = ld1(src + 1);
                if (Ci > 2) v.z = ld1(src + 2);
                if (Ci > 3) v.w = ld1(src + 3);
            }
            xs[i] = v;
        }
        __syncthreads();
        if (ok) {
            const TG* gr = dy + (((int64_t)n * H + h) * W) * Co + co;
            for (int w = 0; w < W; ++w) {
                const float g = ld1(gr + (int64_t)w * Co);
#pragma unroll
                for (int kh = 0; kh < 3; ++kh)
#pragma unroll
                    for (int kw = 0; kw < 3; ++kw) {
                        const float4 v = xs[kh * WP + w + kw];
                        acc[kh * 3 + kw][0] = fmaf(v.x, g, acc[kh * 3 + kw][0]);
                        acc[kh * 3 + kw][1] = fmaf(v.y, g, acc[kh * 3 + kw][1]);
                        acc[kh * 3 + kw][2] = fmaf(v.z, g, acc[kh * 3 + kw][2]);
                        acc[kh * 3 + kw][3] = fmaf(v.w, g, acc[kh * 3 + kw][3]);
                    }
            }
        }
    }
    if (ok) {
#pragma unroll
        for (int t = 0; t < 9; ++t)
#pragma unroll
            for (int c = 0; c < 4; ++c)
                if (c < Ci) atomicAdd(dwp + ((int64_t)t * Ci + c) * Co + co, acc[t][c]);
    }
}

// ---- pointwise (1x1) convolution with a narrow input (Ci <= 4): the discriminator's fromrgb layer 3 -> 128 at full
// resolution.  HBM-bound (write y once / read dy once): threadIdx.x owns 8 consecutive output channels (its Ci x 8 weights
// live in registers), threadIdx.y / blockIdx.x stride over pixels; 16-byte stores / loads along channels.
template <typename T> __device__ __forceinline__ void pw_store8(T* p, const float* v);
template <> __device__ __forceinline__ void pw_store8<float>(float* p, const float* v) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
template <> __device__ __forceinline__ void pw_store8<bf16>(bf16* p, const float* v) {
    uint4 u;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = u;
}
template <typename T> __device__ __forceinline__ void pw_load8(const T* p, float* v);
template <> __device__ __forceinline__ void pw_load8<float>(const float* p, float* v) {
    float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <> __device__ __forceinline__ void pw_load8<bf16>(const bf16* p, float* v) {
    uint4 u = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}

template <typename TIn, typename TOut>
__global__ void pw_narrow_ci_fwd_kernel(const TIn* __restrict__ x, const float* __restrict__ wp, const float* __restrict__ bias,
                                        TOut* __restrict__ y, int64_t P, int Ci, int Co, int act, float alpha, float gain) {
    const int c0 = threadIdx.x * 8;
    float wr[4][8], bv[8];
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int u = 0; u < 8; ++u) wr[c][u] = (c < Ci) ? wp[(int64_t)c * Co + c0 + u] : 0.f;      // wp = [(ci)][co]
#pragma unroll
    for (int u = 0; u < 8; ++u) bv[u] = bias ? bias[c0 + u] : 0.f;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.y + threadIdx.y; p < P; p += (int64_t)gridDim.x * blockDim.y) {
        float xv[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int c = 0; c < 4; ++c) if (c < Ci) xv[c] = ld1(x + p * Ci + c);
        float o[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            float t = bv[u];
#pragma unroll
            for (int c = 0; c < 4; ++c) t = fmaf(xv[c], wr[c][u], t);
            o[u] = apply_act(t, act, alpha) * gain;
        }
        pw_store8<TOut>(y + p * Co + c0, o);
    }
}

// dwp[(ci)][co] += sum_p x[p][ci] * dy[p][co]
template <typename TIn, typename TG>
__global__ void pw_narrow_ci_wgrad_kernel(const TIn* __restrict__ x, const TG* __restrict__ dy, float* __restrict__ dwp, int64_t P,
                                          int Ci, int Co, int rows_per_block) {
    extern __shared__ float shw[];                               // [ty][4][Co]
    const int c0 = threadIdx.x * 8;
    float acc[4][8];
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int u = 0; u < 8; ++u) acc[c][u] = 0.f;
    const int64_t p0 = (int64_t)blockIdx.x * rows_per_block;
    int64_t p1 = p0 + rows_per_block; if (p1 > P) p1 = P;
#pragma unroll 4
    for (int64_t p = p0 + threadIdx.y; p < p1; p += blockDim.y) {
        float g[8], xv[4] = {0.f, 0.f, 0.f, 0.f};
        pw_load8<TG>(dy + p * Co + c0, g);
#pragma unroll
        for (int c = 0; c < 4; ++c) if (c < Ci) xv[c] = ld1(x + p * Ci + c);
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int u = 0; u < 8; ++u) acc[c][u] = fmaf(xv[c], g[u], acc[c][u]);
    }
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int u = 0; u < 8; ++u) shw[((size_t)threadIdx.y * 4 + c) * Co + c0 + u] = acc[c][u];
    __syncthreads();
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    for (int i = tid; i < Ci * Co; i += blockDim.x * blockDim.y) {
        const int c = i / Co, co = i - c * Co;
        float t = 0.f;
        for (int yy = 0; yy < (int)blockDim.y; ++yy) t += shw[((size_t)yy * 4 + c) * Co + co];
        atomicAdd(dwp + (int64_t)c * Co + co, t);
    }
}

inline int make_geom(ConvGeom& g, int N, int H, int W, int Ci, int Co, int KH, int KW, int pad, int stride) {
    if (!(N > 0 && H > 0 && W > 0 && Ci > 0 && Co > 0 && KH > 0 && KW > 0 && pad >= 0 && stride >= 1)) {
        vqb_set_error("conv2d: bad geometry N=%d H=%d W=%d Ci=%d Co=%d KH=%d KW=%d pad=%d stride=%d", N, H, W, Ci, Co, KH, KW, pad, stride);
        return VQB_ERR_ARG;
    }
    g.N = N; g.H = H; g.W = W; g.Ci = Ci; g.Co = Co; g.KH = KH; g.KW = KW; g.pad = pad; g.stride = stride; g.up = 1;
    g.OH = (H + 2 * pad - KH) / stride + 1;
    g.OW = (W + 2 * pad - KW) / stride + 1;
    if (g.OH <= 0 || g.OW <= 0) { vqb_set_error("conv2d: empty output"); return VQB_ERR_ARG; }
    g.M = (int64_t)N * g.OH * g.OW;
    g.K = KH * KW * Ci;
    return VQB_OK;
}

}  // namespace

int vqb_conv2d_fwd_simt(const void* x, int x_dtype, const float* wp, const float* bias, const void* residual, void* y,
                        int y_dtype, int N, int H, int W, int Ci, int Co, int KH, int KW, int pad, int stride, int act,
                        float alpha, float gain, cudaStream_t stream) {
    ConvGeom g;
    int rc = make_geom(g, N, H, W, Ci, Co, KH, KW, pad, stride); if (rc) return rc;
    if (Ci <= 4 && KH == 3 && KW == 3 && pad == 1 && stride == 1 && act == VQB_ACT_NONE && gain == 1.0f && !residual &&
        (size_t)3 * (W + 2) * 16 <= 48 * 1024) {
        int rows = N * H;
        dim3 ngrid(rows < 148 * 8 ? rows : 148 * 8, (Co + 127) / 128);
        size_t sm = (size_t)3 * (W + 2) * sizeof(float4);
        VQB_DISPATCH_1(x_dtype, TIn, VQB_DISPATCH_1(y_dtype, TOut,
            (conv_fwd_narrow_ci_kernel<TIn, TOut><<<ngrid, 128, sm, stream>>>((const TIn*)x, wp, bias, (TOut*)y, N, H, W, Ci, Co));))
        VQB_CHECK_LAUNCH("conv2d_fwd_narrow_ci");
        return VQB_OK;
    }
    if (Ci <= 4 && KH == 1 && KW == 1 && pad == 0 && stride == 1 && !residual && Co % 8 == 0 && Co / 8 <= 256 &&
        (int64_t)N * H * W >= 4096) {
        const int tx = Co / 8, ty = 256 / tx > 0 ? 256 / tx : 1;
        const int64_t P = (int64_t)N * H * W;
        int64_t blocks = ceil_div64(P, ty); if (blocks > 148 * 16) blocks = 148 * 16;
        VQB_DISPATCH_1(x_dtype, TIn, VQB_DISPATCH_1(y_dtype, TOut,
            (pw_narrow_ci_fwd_kernel<TIn, TOut><<<(unsigned)blocks, dim3(tx, ty), 0, stream>>>((const TIn*)x, wp, bias, (TOut*)y, P, Ci, Co, act, alpha, gain));))
        VQB_CHECK_LAUNCH("conv2d_fwd_pw_narrow_ci");
        return VQB_OK;
    }
    dim3 grid((unsigned)ceil_div64(g.M, BM), (unsigned)((Co + BN - 1) / BN));
    VQB_DISPATCH_1(x_dtype, TIn, VQB_DISPATCH_1(y_dtype, TOut,
        (conv_fwd_simt_kernel<TIn, TOut><<<grid, 256, 0, stream>>>((const TIn*)x, wp, bias, (const TOut*)residual, (TOut*)y, g, act, alpha, gain));))
    VQB_CHECK_LAUNCH("conv2d_fwd_simt");
    return VQB_OK;
}

// dgrad of a (possibly strided) convolution: dx[N,H,W,Ci] from dy[N,OH,OW,Co] and the dgrad-packed weight
// wd[((KH-1-kh)*KW+(KW-1-kw))*Co+co][ci]; a stride-s forward conv becomes a gather over the zero-upsampled dy.
int vqb_conv2d_dgrad_simt(const void* dy, int dy_dtype, const float* wd, void* dx, int dx_dtype, int N, int H, int W, int Ci,
                          int Co, int KH, int KW, int pad, int stride, cudaStream_t stream) {
    ConvGeom g;
    const int OH = (H + 2 * pad - KH) / stride + 1, OW = (W + 2 * pad - KW) / stride + 1;
    if (!(N > 0 && OH > 0 && OW > 0 && Ci > 0 && Co > 0)) { vqb_set_error("conv2d_dgrad: bad geometry"); return VQB_ERR_ARG; }
    g.N = N; g.H = OH; g.W = OW; g.Ci = Co; g.Co = Ci; g.KH = KH; g.KW = KW;
    g.pad = KH - 1 - pad; g.stride = 1; g.up = stride; g.OH = H; g.OW = W;
    g.M = (int64_t)N * H * W; g.K = KH * KW * Co;
    if (KH != KW) { vqb_set_error("conv2d_dgrad: square kernels only"); return VQB_ERR_UNSUPPORTED; }
    dim3 grid((unsigned)ceil_div64(g.M, BM), (unsigned)((Ci + BN - 1) / BN));
    VQB_DISPATCH_1(dy_dtype, TIn, VQB_DISPATCH_1(dx_dtype, TOut,
        (conv_fwd_simt_kernel<TIn, TOut><<<grid, 256, 0, stream>>>((const TIn*)dy, wd, nullptr, (const TOut*)nullptr, (TOut*)dx, g, VQB_ACT_NONE, 0.f, 1.f));))
    VQB_CHECK_LAUNCH("conv2d_dgrad_simt");
    return VQB_OK;
}

int vqb_conv2d_wgrad_simt(const void* x, int x_dtype, const void* dy, int dy_dtype, float* dwp, int N, int H, int W, int Ci,
                          int Co, int KH, int KW, int pad, int stride, cudaStream_t stream) {
    ConvGeom g;
    int rc = make_geom(g, N, H, W, Ci, Co, KH, KW, pad, stride); if (rc) return rc;
    if (Co <= 4 && KH == 3 && KW == 3 && pad == 1 && stride == 1 && (size_t)3 * (W + 2) * 16 <= 48 * 1024) {
        int rows = N * H;
        dim3 grid(rows < 148 * 8 ? rows : 148 * 8, (Ci + 127) / 128);
        size_t sm = (size_t)3 * (W + 2) * sizeof(float4);
        VQB_DISPATCH_1(x_dtype, TIn, VQB_DISPATCH_1(dy_dtype, TG,
            (conv_wgrad_narrow_kernel<TIn, TG><<<grid, 128, sm, stream>>>((const TIn*)x, (const TG*)dy, dwp, N, H, W, Ci, Co));))
        VQB_CHECK_LAUNCH("conv2d_wgrad_narrow");
        return VQB_OK;
    }
    if (Ci <= 4 && KH == 3 && KW == 3 && pad == 1 && stride == 1 && (size_t)3 * (W + 2) * 16 <= 48 * 1024) {
        int rows = N * H;
        dim3 grid(rows < 148 * 8 ? rows : 148 * 8, (Co + 127) / 128);
        size_t sm = (size_t)3 * (W + 2) * sizeof(float4);
        VQB_DISPATCH_1(x_dtype, TIn, VQB_DISPATCH_1(dy_dtype, TG,
            (conv_wgrad_narrow_ci_kernel<TIn, TG><<<grid, 128, sm, stream>>>((const TIn*)x, (const TG*)dy, dwp, N, H, W, Ci, Co));))
        VQB_CHECK_LAUNCH("conv2d_wgrad_narrow_ci");
        return VQB_OK;
    }
    if (Ci <= 4 && KH == 1 && KW == 1 && pad == 0 && stride == 1 && Co % 8 == 0 && Co / 8 <= 256 && (int64_t)N * H * W >= 4096) {
        const int tx = Co / 8, ty = 256 / tx > 0 ? 256 / tx : 1;
        const int64_t P = (int64_t)N * H * W;
        int rows = (int)ceil_div64(P, (int64_t)148 * 8); if (rows < ty * 16) rows = ty * 16;
        const unsigned blocks = (unsigned)ceil_div64(P, rows);
        const size_t sm = sizeof(float) * ty * 4 * Co;
        VQB_DISPATCH_1(x_dtype, TIn, VQB_DISPATCH_1(dy_dtype, TG,
            (pw_narrow_ci_wgrad_kernel<TIn, TG><<<blocks, dim3(tx, ty), sm, stream>>>((const TIn*)x, (const TG*)dy, dwp, P, Ci, Co, rows));))
        VQB_CHECK_LAUNCH("conv2d_wgrad_pw_narrow_ci");
        return VQB_OK;
    }
    int gm = (g.K + BM - 1) / BM, gn = (Co + BN - 1) / BN;
    // split the pixel reduction so that the grid is ~4 waves of 148 SMs x 2 resident CTAs
    int64_t want = (int64_t)148 * 2 * 4;
    int64_t splits = want / ((int64_t)gm * gn); if (splits < 1) splits = 1;
    int64_t max_splits = ceil_div64(g.M, 64); if (splits > max_splits) splits = max_splits;
    int64_t chunk = ceil_div64(ceil_div64(g.M, splits), BK) * BK;
    splits = ceil_div64(g.M, chunk);
    dim3 grid(gm, gn, (unsigned)splits);
    VQB_DISPATCH_1(x_dtype, TIn, VQB_DISPATCH_1(dy_dtype, TG,
        (conv_wgrad_simt_kernel<TIn, TG><<<grid, 256, 0, stream>>>((const TIn*)x, (const TG*)dy, dwp, g, chunk));))
    VQB_CHECK_LAUNCH("conv2d_wgrad_simt");
    return VQB_OK;
}
