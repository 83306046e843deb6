// GroupNorm with UNBIASED variance (reference: vqvae/modules/autoencoder.py:25-39) fused with SiLU, NHWC.
//
// HBM-bound.  Forward = one statistics pass (read x) + one apply pass (read x, write y).  Backward = one
// reduction pass (read x, dy) + one apply pass (read x, dy, write dx).  Thread mapping: threadIdx.x owns VEC
// consecutive channels (fixed for the whole kernel, so per-channel parameters live in registers and global
// loads of one pixel row are fully coalesced), threadIdx.y strides over pixels.  Partial sums are fp32 over
// <= 32 pixels per thread, then promoted to double for the block / grid combine (E[x^2]-mu^2 is evaluated in
// double, so there is no catastrophic cancellation at n = 4*65536 elements per group).
#include "common.cuh"

namespace {

struct GnLaunch {
    dim3 grid, block;
    int vec, ppb;
};

inline GnLaunch gn_launch(int N, int HW, int C, int G) {
    GnLaunch L;
    L.vec = (C % 4 == 0 && (C / G) % 4 == 0 && C / 4 >= G) ? 4 : 1;   // VEC=4 keeps a float4 inside one group
    int tx = C / L.vec;
    int ty = 256 / tx; if (ty < 1) ty = 1;
    if (ty > 32) ty = 32;
    L.block = dim3(tx, ty);
    L.ppb = ty * 32;                   // pixels per block: 32 per thread row
    if (L.ppb > HW) L.ppb = ((HW + ty - 1) / ty) * ty;
    L.grid = dim3((HW + L.ppb - 1) / L.ppb, N);
    return L;
}

template <typename T, int VEC>
__device__ __forceinline__ void ldv(const T* p, float (&v)[VEC]) {
    if (VEC == 4) { float4 t = ld4(p); v[0] = t.x; v[1 % VEC] = t.y; v[2 % VEC] = t.z; v[3 % VEC] = t.w; }
    else v[0] = ld1(p);
}
template <typename T, int VEC>
__device__ __forceinline__ void stv(T* p, const float (&v)[VEC]) {
    if (VEC == 4) st4(p, make_float4(v[0], v[1 % VEC], v[2 % VEC], v[3 % VEC]));
    else st1(p, v[0]);
}

// ---- forward statistics ---------------------------------------------------------------------------
template <typename T, int VEC>
__global__ void gn_stats_kernel(const T* __restrict__ x, double* __restrict__ sums, int HW, int C, int G, int ppb) {
    extern __shared__ double sh[];   // [2][blockDim.y][blockDim.x]
    const int b = blockIdx.y;
    const int c0 = threadIdx.x * VEC;
    const int p0 = blockIdx.x * ppb;
    int p1 = p0 + ppb; if (p1 > HW) p1 = HW;
    float s = 0.f, ss = 0.f;
    const T* xb = x + (int64_t)b * HW * C + c0;
    for (int p = p0 + threadIdx.y; p < p1; p += blockDim.y) {
        float v[VEC];
        ldv<T, VEC>(xb + (int64_t)p * C, v);
#pragma unroll
        for (int j = 0; j < VEC; ++j) { s += v[j]; ss += v[j] * v[j]; }
    }
    const int tx = blockDim.x, ty = blockDim.y;
    sh[threadIdx.y * tx + threadIdx.x] = (double)s;
    sh[(ty + threadIdx.y) * tx + threadIdx.x] = (double)ss;
    __syncthreads();
    const int tid = threadIdx.y * tx + threadIdx.x;
    const int cg = C / G;                         // channels per group
    const int tpg = (cg >= VEC) ? cg / VEC : 1;   // threads (in x) per group
    const int gpt = (cg >= VEC) ? 1 : VEC / cg;   // groups per thread (only when VEC=1 -> 1)
    (void)gpt;
    if (tid < G) {
        double a = 0.0, q = 0.0;
        for (int y = 0; y < ty; ++y)
            for (int t = 0; t < tpg; ++t) {
                a += sh[y * tx + tid * tpg + t];
                q += sh[(ty + y) * tx + tid * tpg + t];
            }
        atomicAdd(&sums[((int64_t)b * G + tid) * 2 + 0], a);
        atomicAdd(&sums[((int64_t)b * G + tid) * 2 + 1], q);
    }
}

__global__ void gn_finalize_kernel(const double* __restrict__ sums, float* __restrict__ stats, int total, double n, double eps) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    double s = sums[2 * i], q = sums[2 * i + 1];
    double mean = s / n;
    double var = (q - s * mean) / (n - 1.0);      // unbiased: torch.var default (autoencoder.py:31)
    if (var < 0.0) var = 0.0;
    stats[2 * i] = (float)mean;
    stats[2 * i + 1] = (float)(1.0 / sqrt(var + eps));
}

// ---- forward apply -----------------------------------------------------------------------------------
template <typename TI, typename TO, int VEC>
__global__ void gn_apply_kernel(const TI* __restrict__ x, const float* __restrict__ stats, const float* __restrict__ gamma,
                                const float* __restrict__ beta, TO* __restrict__ y, int HW, int C, int G, int ppb, int act) {
    const int b = blockIdx.y;
    const int c0 = threadIdx.x * VEC;
    const int cg = C / G;
    float mean[VEC], rstd[VEC], ga[VEC], be[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
        int g = (c0 + j) / cg;
        mean[j] = stats[((int64_t)b * G + g) * 2];
        rstd[j] = stats[((int64_t)b * G + g) * 2 + 1];
        ga[j] = gamma[c0 + j];
        be[j] = beta[c0 + j];
    }
    const int p0 = blockIdx.x * ppb;
    int p1 = p0 + ppb; if (p1 > HW) p1 = HW;
    const int64_t base = (int64_t)b * HW * C + c0;
    for (int p = p0 + threadIdx.y; p < p1; p += blockDim.y) {
        float v[VEC], o[VEC];
        ldv<TI, VEC>(x + base + (int64_t)p * C, v);
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            float t = (v[j] - mean[j]) * rstd[j] * ga[j] + be[j];
            o[j] = (act == VQB_ACT_SILU) ? silu_f(t) : t;
        }
        stv<TO, VEC>(y + base + (int64_t)p * C, o);
    }
}

// ---- backward reduce ---------------------------------------------------------------------------------
template <typename TI, typename TG, int VEC>
__global__ void gn_bwd_reduce_kernel(const TI* __restrict__ x, const TG* __restrict__ dy, const float* __restrict__ stats,
                                     const float* __restrict__ gamma, const float* __restrict__ beta,
                                     double* __restrict__ part, int HW, int C, int G, int ppb, int act) {
    extern __shared__ double sh[];   // [2][ty][tx*VEC]
    const int b = blockIdx.y;
    const int c0 = threadIdx.x * VEC;
    const int cg = C / G;
    float mean[VEC], rstd[VEC], ga[VEC], be[VEC], a[VEC], q[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
        int g = (c0 + j) / cg;
        mean[j] = stats[((int64_t)b * G + g) * 2];
        rstd[j] = stats[((int64_t)b * G + g) * 2 + 1];
        ga[j] = gamma[c0 + j];
        be[j] = beta[c0 + j];
        a[j] = 0.f; q[j] = 0.f;
    }
    const int p0 = blockIdx.x * ppb;
    int p1 = p0 + ppb; if (p1 > HW) p1 = HW;
    const int64_t base = (int64_t)b * HW * C + c0;
    for (int p = p0 + threadIdx.y; p < p1; p += blockDim.y) {
        float v[VEC], g[VEC];
        ldv<TI, VEC>(x + base + (int64_t)p * C, v);
        ldv<TG, VEC>(dy + base + (int64_t)p * C, g);
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            float xh = (v[j] - mean[j]) * rstd[j];
            float ds = g[j];
            if (act == VQB_ACT_SILU) ds *= silu_grad_f(xh * ga[j] + be[j]);
            a[j] += ds;
            q[j] += ds * xh;
        }
    }
    const int tx = blockDim.x, ty = blockDim.y, row = tx * VEC;
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
        sh[threadIdx.y * row + c0 + j] = (double)a[j];
        sh[(ty + threadIdx.y) * row + c0 + j] = (double)q[j];
    }
    __syncthreads();
    const int tid = threadIdx.y * tx + threadIdx.x;
    for (int c = tid; c < C; c += tx * ty) {
        double sa = 0.0, sq = 0.0;
        for (int y = 0; y < ty; ++y) { sa += sh[y * row + c]; sq += sh[(ty + y) * row + c]; }
        atomicAdd(&part[((int64_t)b * C + c) * 2 + 0], sa);
        atomicAdd(&part[((int64_t)b * C + c) * 2 + 1], sq);
    }
}

// coef[b][g] and parameter grads
__global__ void gn_bwd_finalize_kernel(const double* __restrict__ part, const float* __restrict__ gamma,
                                       float* __restrict__ coef, float* __restrict__ dgamma, float* __restrict__ dbeta,
                                       int N, int C, int G, double n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int cg = C / G;
    if (i < N * G) {
        int b = i / G, g = i % G;
        double s1 = 0.0, s2 = 0.0;
        for (int c = g * cg; c < (g + 1) * cg; ++c) {
            double ga = (double)gamma[c];
            s1 += ga * part[((int64_t)b * C + c) * 2];
            s2 += ga * part[((int64_t)b * C + c) * 2 + 1];
        }
        coef[2 * i] = (float)(s1 / n);
        coef[2 * i + 1] = (float)(s2 / (n - 1.0));
    }
    if (i < C) {
        double db = 0.0, dg = 0.0;
        for (int b = 0; b < N; ++b) {
            db += part[((int64_t)b * C + i) * 2];
            dg += part[((int64_t)b * C + i) * 2 + 1];
        }
        dbeta[i] = (float)db;
        dgamma[i] = (float)dg;
    }
}

template <typename TI, typename TG, typename TO, int VEC>
__global__ void gn_bwd_apply_kernel(const TI* __restrict__ x, const TG* __restrict__ dy, const float* __restrict__ stats,
                                    const float* __restrict__ gamma, const float* __restrict__ beta,
                                    const float* __restrict__ coef, TO* __restrict__ dx, int HW, int C, int G, int ppb,
                                    int act) {
    const int b = blockIdx.y;
    const int c0 = threadIdx.x * VEC;
    const int cg = C / G;
    float mean[VEC], rstd[VEC], ga[VEC], be[VEC], k1[VEC], k2[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
        int g = (c0 + j) / cg;
        mean[j] = stats[((int64_t)b * G + g) * 2];
        rstd[j] = stats[((int64_t)b * G + g) * 2 + 1];
        k1[j] = coef[((int64_t)b * G + g) * 2];
        k2[j] = coef[((int64_t)b * G + g) * 2 + 1];
        ga[j] = gamma[c0 + j];
        be[j] = beta[c0 + j];
    }
    const int p0 = blockIdx.x * ppb;
    int p1 = p0 + ppb; if (p1 > HW) p1 = HW;
    const int64_t base = (int64_t)b * HW * C + c0;
    for (int p = p0 + threadIdx.y; p < p1; p += blockDim.y) {
        float v[VEC], g[VEC], o[VEC];
        ldv<TI, VEC>(x + base + (int64_t)p * C, v);
        ldv<TG, VEC>(dy + base + (int64_t)p * C, g);
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            float xh = (v[j] - mean[j]) * rstd[j];
            float ds = g[j];
            if (act == VQB_ACT_SILU) ds *= silu_grad_f(xh * ga[j] + be[j]);
            o[j] = rstd[j] * (ds * ga[j] - k1[j] - xh * k2[j]);
        }
        stv<TO, VEC>(dx + base + (int64_t)p * C, o);
    }
}

inline int gn_check(const char* name, int N, int HW, int C, int G) {
    if (!(N > 0 && HW > 0 && C > 0 && G > 0 && C % G == 0)) {
        vqb_set_error("%s: bad shape N=%d HW=%d C=%d G=%d", name, N, HW, C, G);
        return VQB_ERR_ARG;
    }
    { int vec = (C % 4 == 0 && (C / G) % 4 == 0 && C / 4 >= G) ? 4 : 1;
      if (C / vec > 1024) { vqb_set_error("%s: C=%d too large", name, C); return VQB_ERR_UNSUPPORTED; } }
    if ((double)(C / G) * HW < 2.0) { vqb_set_error("%s: unbiased variance needs >= 2 elements per group", name); return VQB_ERR_ARG; }
    return VQB_OK;
}

}  // namespace

#define GN_VEC_DISPATCH(L, ...) \
    if ((L).vec == 4) { constexpr int VEC = 4; __VA_ARGS__ } else { constexpr int VEC = 1; __VA_ARGS__ }

extern "C" int vqb_gn_stats(const void* x, int x_dtype, double* sums, int N, int HW, int C, int G, void* stream) {
    int rc = gn_check("gn_stats", N, HW, C, G); if (rc) return rc;
    VQB_CHECK_ARG(x && sums, "gn_stats: null pointer");
    GnLaunch L = gn_launch(N, HW, C, G);
    VQB_CHECK_ARG((int)(L.block.x * L.block.y) >= G, "gn_stats: block smaller than group count");
    size_t sm = 2 * sizeof(double) * L.block.x * L.block.y;
    GN_VEC_DISPATCH(L, VQB_DISPATCH_1(x_dtype, T, (gn_stats_kernel<T, VEC><<<L.grid, L.block, sm, as_stream(stream)>>>(
                                                      (const T*)x, sums, HW, C, G, L.ppb));))
    VQB_CHECK_LAUNCH("gn_stats");
    return VQB_OK;
}

extern "C" int vqb_gn_finalize(const double* sums, float* stats, int N, int HW, int C, int G, float eps, void* stream) {
    int rc = gn_check("gn_finalize", N, HW, C, G); if (rc) return rc;
    VQB_CHECK_ARG(sums && stats, "gn_finalize: null pointer");
    int total = N * G;
    gn_finalize_kernel<<<(total + 127) / 128, 128, 0, as_stream(stream)>>>(sums, stats, total, (double)(C / G) * HW, (double)eps);
    VQB_CHECK_LAUNCH("gn_finalize");
    return VQB_OK;
}

extern "C" int vqb_gn_apply(const void* x, int x_dtype, const float* stats, const float* gamma, const float* beta, void* y,
                            int y_dtype, int N, int HW, int C, int G, int act, void* stream) {
    int rc = gn_check("gn_apply", N, HW, C, G); if (rc) return rc;
    VQB_CHECK_ARG(x && stats && gamma && beta && y, "gn_apply: null pointer");
    VQB_CHECK_ARG(act == VQB_ACT_NONE || act == VQB_ACT_SILU, "gn_apply: act must be NONE or SILU");
    GnLaunch L = gn_launch(N, HW, C, G);
    GN_VEC_DISPATCH(L, VQB_DISPATCH_1(x_dtype, TI, VQB_DISPATCH_1(y_dtype, TO,
        (gn_apply_kernel<TI, TO, VEC><<<L.grid, L.block, 0, as_stream(stream)>>>((const TI*)x, stats, gamma, beta, (TO*)y, HW, C, G, L.ppb, act));)))
    VQB_CHECK_LAUNCH("gn_apply");
    return VQB_OK;
}

extern "C" int vqb_gn_bwd_reduce(const void* x, int x_dtype, const void* dy, int dy_dtype, const float* stats,
                                 const float* gamma, const float* beta, double* part, int N, int HW, int C, int G, int act,
                                 void* stream) {
    int rc = gn_check("gn_bwd_reduce", N, HW, C, G); if (rc) return rc;
    VQB_CHECK_ARG(x && dy && stats && gamma && beta && part, "gn_bwd_reduce: null pointer");
    GnLaunch L = gn_launch(N, HW, C, G);
    size_t sm = 2 * sizeof(double) * L.block.x * L.block.y * L.vec;
    GN_VEC_DISPATCH(L, VQB_DISPATCH_1(x_dtype, TI, VQB_DISPATCH_1(dy_dtype, TG,
        (gn_bwd_reduce_kernel<TI, TG, VEC><<<L.grid, L.block, sm, as_stream(stream)>>>((const TI*)x, (const TG*)dy, stats, gamma, beta, part, HW, C, G, L.ppb, act));)))
    VQB_CHECK_LAUNCH("gn_bwd_reduce");
    return VQB_OK;
}

extern "C" int vqb_gn_bwd_finalize(const double* part, const float* gamma, float* coef, float* dgamma, float* dbeta, int N,
                                   int HW, int C, int G, void* stream) {
    int rc = gn_check("gn_bwd_finalize", N, HW, C, G); if (rc) return rc;
    VQB_CHECK_ARG(part && gamma && coef && dgamma && dbeta, "gn_bwd_finalize: null pointer");
    int total = (N * G > C) ? N * G : C;
    gn_bwd_finalize_kernel<<<(total + 127) / 128, 128, 0, as_stream(stream)>>>(part, gamma, coef, dgamma, dbeta, N, C, G,
                                                                                (double)(C / G) * HW);
    VQB_CHECK_LAUNCH("gn_bwd_finalize");
    return VQB_OK;
}

extern "C" int vqb_gn_bwd_apply(const void* x, int x_dtype, const void* dy, int dy_dtype, const float* stats,
                                const float* gamma, const float* beta, const float* coef, void* dx, int dx_dtype, int N,
                                int HW, int C, int G, int act, void* stream) {
    int rc = gn_check("gn_bwd_apply", N, HW, C, G); if (rc) return rc;
    VQB_CHECK_ARG(x && dy && stats && gamma && beta && coef && dx, "gn_bwd_apply: null pointer");
    GnLaunch L = gn_launch(N, HW, C, G);
    GN_VEC_DISPATCH(L, VQB_DISPATCH_1(x_dtype, TI, VQB_DISPATCH_1(dy_dtype, TG, VQB_DISPATCH_1(dx_dtype, TO,
        (gn_bwd_apply_kernel<TI, TG, TO, VEC><<<L.grid, L.block, 0, as_stream(stream)>>>((const TI*)x, (const TG*)dy, stats, gamma, beta, coef, (TO*)dx, HW, C, G, L.ppb, act));))))
    VQB_CHECK_LAUNCH("gn_bwd_apply");
    return VQB_OK;
}
