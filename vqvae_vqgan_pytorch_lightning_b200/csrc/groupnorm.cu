// GroupNorm with UNBIASED variance (reference: vqvae/modules/autoencoder.py:25-39) fused with SiLU, NHWC.
//
// HBM-bound.  Forward = one statistics pass (read x) + one apply pass (read x, write y).  Backward = one
// reduction pass (read x, dy) + one apply pass (read x, dy, write dx).  Thread mapping: threadIdx.x owns VEC
// consecutive channels (fixed for the whole kernel, so per-channel parameters live in registers and global
// loads of one pixel row are fully coalesced 16-byte accesses: VEC = 8 for bf16 / wide layers), threadIdx.y strides
// over pixels with a 2-deep unrolled loop (two independent 16-byte loads in flight per thread).  Partial sums are
// fp32 over <= 32 pixels per thread, then promoted to double for the block / grid combine (E[x^2]-mu^2 is
// evaluated in double, so there is no catastrophic cancellation at n = 4*65536 elements per group).
#include "common.cuh"

namespace {

// sigmoid via MUFU ex2 + fast reciprocal: ~1e-6 relative error, a fraction of the instruction count of expf + IEEE division
// (two MUFU ops per element).  At full HBM rate these kernels process ~6 elements / clk / SM, i.e. ~12 of the 16 MUFU
// issue slots per clock -- the SFU pipe, not HBM, paces them.  For bf16 activations (APPROX) the sigmoid is ONE MUFU op,
// 0.5 * tanh.approx(0.5 t) + 0.5 (max abs error ~2.5e-4, an order of magnitude below bf16 rounding of the result).
__device__ __forceinline__ float tanh_approx(float x) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
template <bool APPROX>
__device__ __forceinline__ float fast_sigmoid(float t) {
    if constexpr (APPROX) return fmaf(0.5f, tanh_approx(0.5f * t), 0.5f);
    else return __fdividef(1.0f, 1.0f + __expf(-t));
}
template <bool APPROX> __device__ __forceinline__ float fast_silu(float t) { return t * fast_sigmoid<APPROX>(t); }
// the same two functions of u = t/2 (the halving is folded into the per-channel scale/shift by the callers):
//   silu(t) = u (1 + tanh u);   2 silu'(t) = (1 + tanh u)(1 + u (1 - tanh u))      -- 3 and 5 instructions incl. the MUFU
__device__ __forceinline__ float silu_of_half(float u) { return fmaf(u, tanh_approx(u), u); }
__device__ __forceinline__ float silu_grad2_of_half(float u) {
    const float th = tanh_approx(u);
    const float k = 1.0f + fmaf(-u, th, u);
    return fmaf(th, k, k);
}
template <bool APPROX>
__device__ __forceinline__ float fast_silu_grad(float t) { float s = fast_sigmoid<APPROX>(t); return s * fmaf(t, 1.0f - s, 1.0f); }

struct GnLaunch {
    dim3 grid, block;
    int vec, ppb;
};

inline int gn_vec(int C, int G) {
    const int cg = C / G;
    if (C % 8 == 0 && cg % 4 == 0 && C / 8 <= 1024) return 8;   // each 4-channel half stays inside one group
    if (C % 4 == 0 && cg % 4 == 0 && C / 4 >= G) return 4;
    return 1;
}

inline GnLaunch gn_launch(int N, int HW, int C, int G, int pix_per_thread = 32, int max_vec = 8) {
    GnLaunch L;
    L.vec = gn_vec(C, G);
    if (L.vec > max_vec) L.vec = max_vec;     // cg % 4 == 0 holds whenever gn_vec returned 8, so 4 is valid too
    int tx = C / L.vec;
    int ty = 256 / tx; if (ty < 1) ty = 1;
    if (ty > 32) ty = 32;
    L.block = dim3(tx, ty);
    L.ppb = ty * pix_per_thread;       // pixels per block (reductions use more: fewer double atomics per byte read)
    if (L.ppb > HW) L.ppb = ((HW + ty - 1) / ty) * ty;
    L.grid = dim3((HW + L.ppb - 1) / L.ppb, N);
    return L;
}

template <typename T, int VEC>
__device__ __forceinline__ void ldv(const T* p, float (&v)[VEC]) {
    if constexpr (VEC == 8) {
        if constexpr (sizeof(T) == 2) {
            uint4 u = *reinterpret_cast<const uint4*>(p);
            const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
            for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
        } else {
            float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        }
    } else if constexpr (VEC == 4) {
        float4 t = ld4(p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else {
        v[0] = ld1(p);
    }
}
template <typename T, int VEC>
__device__ __forceinline__ void stv(T* p, const float (&v)[VEC]) {
    if constexpr (VEC == 8) {
        if constexpr (sizeof(T) == 2) {
            uint4 u;
            __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
            for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
            *reinterpret_cast<uint4*>(p) = u;
        } else {
            *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
        }
    } else if constexpr (VEC == 4) {
        st4(p, make_float4(v[0], v[1], v[2], v[3]));
    } else {
        st1(p, v[0]);
    }
}

// ---- forward statistics ---------------------------------------------------------------------------
// accumulation unit = UNIT consecutive channels (4 when VEC >= 4, else 1); a group is C/G/UNIT units
template <typename T, int VEC>
__global__ void gn_stats_kernel(const T* __restrict__ x, double* __restrict__ sums, int HW, int C, int G, int ppb) {
    constexpr int UNIT = (VEC >= 4) ? 4 : 1;
    constexpr int NU = VEC / UNIT;
    extern __shared__ double sh[];   // [2][blockDim.y][blockDim.x * NU]
    const int b = blockIdx.y;
    const int c0 = threadIdx.x * VEC;
    const int p0 = blockIdx.x * ppb;
    int p1 = p0 + ppb; if (p1 > HW) p1 = HW;
    float s[NU], ss[NU];
#pragma unroll
    for (int u = 0; u < NU; ++u) { s[u] = 0.f; ss[u] = 0.f; }
    const T* xb = x + (int64_t)b * HW * C + c0;
#pragma unroll 2
    for (int p = p0 + threadIdx.y; p < p1; p += blockDim.y) {
        float v[VEC];
        ldv<T, VEC>(xb + (int64_t)p * C, v);
#pragma unroll
        for (int j = 0; j < VEC; ++j) { s[j / UNIT] += v[j]; ss[j / UNIT] = fmaf(v[j], v[j], ss[j / UNIT]); }
    }
    const int tx = blockDim.x, ty = blockDim.y, row = tx * NU;
#pragma unroll
    for (int u = 0; u < NU; ++u) {
        sh[threadIdx.y * row + threadIdx.x * NU + u] = (double)s[u];
        sh[(ty + threadIdx.y) * row + threadIdx.x * NU + u] = (double)ss[u];
    }
    __syncthreads();
    const int tid = threadIdx.y * tx + threadIdx.x;
    const int upg = (C / G) / UNIT;               // units per group
    for (int g = tid; g < G; g += tx * ty) {
        double a = 0.0, q = 0.0;
        for (int y = 0; y < ty; ++y)
            for (int t = 0; t < upg; ++t) {
                a += sh[y * row + g * upg + t];
                q += sh[(ty + y) * row + g * upg + t];
            }
        atomicAdd(&sums[((int64_t)b * G + g) * 2 + 0], a);
        atomicAdd(&sums[((int64_t)b * G + g) * 2 + 1], q);
    }
}

// bf16 / VEC = 8 statistics with the cp.async ring of gn_bwd_reduce_async_kernel (declared below): same sums
template <int D>
__global__ void gn_stats_async_kernel(const bf16* __restrict__ x, double* __restrict__ sums, int HW, int C, int G, int ppb);

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


// mean / rstd of group g of image b: from the finalized stats, or (sums != NULL) evaluated here from the raw double sums with
// gn_finalize_kernel's arithmetic -- bit-identical, so the folded two-launch forms equal the four-launch ones
__device__ __forceinline__ void gn_group_stats(const float* __restrict__ stats, const double* __restrict__ sums, int64_t bg, double n,
                                               double eps, float& mean, float& rstd) {
    if (sums) {
        const double s = sums[2 * bg], q = sums[2 * bg + 1];
        const double m = s / n;
        double var = (q - s * m) / (n - 1.0);
        if (var < 0.0) var = 0.0;
        mean = (float)m;
        rstd = (float)(1.0 / sqrt(var + eps));
    } else {
        mean = stats[2 * bg];
        rstd = stats[2 * bg + 1];
    }
}
// per-channel mean / rstd of a thread's VEC consecutive channels (at most VEC / 4 + 1 distinct groups: evaluated once per group)
template <int VEC>
__device__ __forceinline__ void gn_thread_stats(const float* __restrict__ stats, const double* __restrict__ sums, float* __restrict__ stats_out,
                                                bool writer, int b, int c0, int cg, int G, double n, double eps, float (&mean)[VEC],
                                                float (&rstd)[VEC]) {
    int gprev = -1;
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
        const int g = (c0 + j) / cg;
        if (g != gprev) {
            gn_group_stats(stats, sums, (int64_t)b * G + g, n, eps, mean[j], rstd[j]);
            if (stats_out && writer && (c0 + j) % cg == 0) {
                stats_out[((int64_t)b * G + g) * 2] = mean[j];
                stats_out[((int64_t)b * G + g) * 2 + 1] = rstd[j];
            }
            gprev = g;
        } else { mean[j] = mean[j - (j > 0)]; rstd[j] = rstd[j - (j > 0)]; }
    }
}
// coef[b][g] = (sum_c gamma ds / n, sum_c gamma ds xhat / (n - 1)): from the finalized buffer or (part != NULL) from the raw
// per-channel double sums with gn_bwd_finalize_kernel's arithmetic
__device__ __forceinline__ void gn_group_coef(const float* __restrict__ coef, const double* __restrict__ part, const float* __restrict__ gamma,
                                              int b, int g, int cg, int C, int G, double n, float& k1, float& k2) {
    if (part) {
        double s1 = 0.0, s2 = 0.0;
        for (int c = g * cg; c < (g + 1) * cg; ++c) {
            const double ga = (double)gamma[c];
            s1 += ga * part[((int64_t)b * C + c) * 2];
            s2 += ga * part[((int64_t)b * C + c) * 2 + 1];
        }
        k1 = (float)(s1 / n);
        k2 = (float)(s2 / (n - 1.0));
    } else {
        k1 = coef[((int64_t)b * G + g) * 2];
        k2 = coef[((int64_t)b * G + g) * 2 + 1];
    }
}
// dgamma[c] / dbeta[c] = sums of part over the batch (gn_bwd_finalize_kernel's second half), by ONE block of the apply kernel
__device__ __forceinline__ void gn_param_grads(const double* __restrict__ part, float* __restrict__ dgamma, float* __restrict__ dbeta, int acc,
                                               int N, int C, int tid, int nthr) {
    for (int c = tid; c < C; c += nthr) {
        double db = 0.0, dg = 0.0;
        for (int b = 0; b < N; ++b) {
            db += part[((int64_t)b * C + c) * 2];
            dg += part[((int64_t)b * C + c) * 2 + 1];
        }
        if (acc) { dbeta[c] += (float)db; dgamma[c] += (float)dg; }
        else { dbeta[c] = (float)db; dgamma[c] = (float)dg; }
    }
}

// ---- forward apply -----------------------------------------------------------------------------------
template <typename TI, typename TO, int VEC>
__global__ void gn_apply_kernel(const TI* __restrict__ x, const float* __restrict__ stats, const float* __restrict__ gamma,
                                const float* __restrict__ beta, TO* __restrict__ y, int HW, int C, int G, int ppb, int act,
                                const double* __restrict__ sums, float* __restrict__ stats_out, double nel, double eps) {
    const int b = blockIdx.y;
    const int c0 = threadIdx.x * VEC;
    const int cg = C / G;
    float sc[VEC], sh_[VEC];          // y = x * sc + sh_  with sc = rstd*gamma, sh_ = beta - mean*rstd*gamma
    {
        float mean[VEC], rstd[VEC];
        gn_thread_stats<VEC>(stats, sums, stats_out, blockIdx.x == 0 && threadIdx.y == 0, b, c0, cg, G, nel, eps, mean, rstd);
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            sc[j] = rstd[j] * gamma[c0 + j];
            sh_[j] = beta[c0 + j] - mean[j] * sc[j];
        }
    }
    constexpr bool APPROX = (sizeof(TI) == 2 && sizeof(TO) == 2);      // bf16 in and out: single-MUFU SiLU on u = t/2
    const bool half_arg = APPROX && act == VQB_ACT_SILU;
    if (half_arg) {
#pragma unroll
        for (int j = 0; j < VEC; ++j) { sc[j] *= 0.5f; sh_[j] *= 0.5f; }
    }
    const int p0 = blockIdx.x * ppb;
    int p1 = p0 + ppb; if (p1 > HW) p1 = HW;
    const int64_t base = (int64_t)b * HW * C + c0;
#pragma unroll 2
    for (int p = p0 + threadIdx.y; p < p1; p += blockDim.y) {
        float v[VEC], o[VEC];
        ldv<TI, VEC>(x + base + (int64_t)p * C, v);
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            float t = fmaf(v[j], sc[j], sh_[j]);
            if constexpr (APPROX) o[j] = (act == VQB_ACT_SILU) ? silu_of_half(t) : t;
            else o[j] = (act == VQB_ACT_SILU) ? fast_silu<false>(t) : t;
        }
        stv<TO, VEC>(y + base + (int64_t)p * C, o);
    }
}

// ---- backward reduce ---------------------------------------------------------------------------------
template <typename TI, typename TG, int VEC, int UNR>
__global__ void gn_bwd_reduce_kernel(const TI* __restrict__ x, const TG* __restrict__ dy, const float* __restrict__ stats,
                                     const float* __restrict__ gamma, const float* __restrict__ beta,
                                     double* __restrict__ part, int HW, int C, int G, int ppb, int act) {
    extern __shared__ double sh[];   // [2][ty][tx*VEC]
    const int b = blockIdx.y;
    const int c0 = threadIdx.x * VEC;
    const int cg = C / G;
    float sc[VEC], sf[VEC], a[VEC], q[VEC];      // t = x*sc + sf is the pre-activation
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
        int g = (c0 + j) / cg;
        const float mean = stats[((int64_t)b * G + g) * 2];
        const float rstd = stats[((int64_t)b * G + g) * 2 + 1];
        sc[j] = rstd * gamma[c0 + j];
        sf[j] = beta[c0 + j] - mean * sc[j];
        a[j] = 0.f; q[j] = 0.f;
    }
    constexpr bool APPROX = (sizeof(TI) == 2 && sizeof(TG) == 2);
    const bool half_arg = APPROX && act == VQB_ACT_SILU;               // accumulate 2 ds, halve the sums once at the end
    if (half_arg) {
#pragma unroll
        for (int j = 0; j < VEC; ++j) { sc[j] *= 0.5f; sf[j] *= 0.5f; }
    }
    const int p0 = blockIdx.x * ppb;
    int p1 = p0 + ppb; if (p1 > HW) p1 = HW;
    const int64_t base = (int64_t)b * HW * C + c0;
#pragma unroll UNR
    for (int p = p0 + threadIdx.y; p < p1; p += blockDim.y) {
        float v[VEC], g[VEC];
        ldv<TI, VEC>(x + base + (int64_t)p * C, v);
        ldv<TG, VEC>(dy + base + (int64_t)p * C, g);
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            float ds = g[j];
            if constexpr (APPROX) { if (act == VQB_ACT_SILU) ds *= silu_grad2_of_half(fmaf(v[j], sc[j], sf[j])); }
            else { if (act == VQB_ACT_SILU) ds *= fast_silu_grad<false>(fmaf(v[j], sc[j], sf[j])); }
            a[j] += ds;
            q[j] = fmaf(ds, v[j], q[j]);             // sum ds*x; converted to sum ds*xhat below
        }
    }
    if (half_arg) {
#pragma unroll
        for (int j = 0; j < VEC; ++j) { a[j] *= 0.5f; q[j] *= 0.5f; }
    }
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
        int g = (c0 + j) / cg;
        const float mean = stats[((int64_t)b * G + g) * 2];
        const float rstd = stats[((int64_t)b * G + g) * 2 + 1];
        q[j] = rstd * (q[j] - mean * a[j]);
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

// ---- backward reduce, bf16, asynchronous-copy pipeline -----------------------------------------------------
// Same result as gn_bwd_reduce_kernel<bf16, bf16, 8>.  The plain-load version is latency-bound (ncu: >80 % of the stall
// samples on the first use of a loaded register, 3.5 TB/s): the loads a thread keeps in flight are limited by registers.
// Here every thread streams ITS OWN 16-byte pieces of x and dy through a private DEPTH-deep ring in shared memory with
// cp.async (LDGSTS, no registers held while in flight, no inter-thread synchronisation -- only cp.async.wait_group):
// bytes in flight per SM = DEPTH * 32 B * resident threads.
constexpr int GN_ASYNC_DEPTH = 12;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void unpack8(const uint4& u, float (&v)[8]) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}

template <int D>
__global__ void gn_bwd_reduce_async_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy, const float* __restrict__ stats,
                                           const float* __restrict__ gamma, const float* __restrict__ beta,
                                           double* __restrict__ part, int HW, int C, int G, int ppb, int act) {
    constexpr int VEC = 8;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint4* ring = reinterpret_cast<uint4*>(smem_raw);            // [D][2][nthreads]
    double* sh = reinterpret_cast<double*>(smem_raw);            // reused after the loop: [2][ty][tx*VEC]
    const int b = blockIdx.y;
    const int c0 = threadIdx.x * VEC;
    const int cg = C / G;
    const int tx = blockDim.x, ty = blockDim.y, nthr = tx * ty;
    const int tid = threadIdx.y * tx + threadIdx.x;
    float sc[VEC], sf[VEC], a[VEC], q[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
        int g = (c0 + j) / cg;
        const float mean = stats[((int64_t)b * G + g) * 2];
        const float rstd = stats[((int64_t)b * G + g) * 2 + 1];
        sc[j] = rstd * gamma[c0 + j];
        sf[j] = beta[c0 + j] - mean * sc[j];
        a[j] = 0.f; q[j] = 0.f;
    }
    const bool silu = (act == VQB_ACT_SILU);
    if (silu) {
#pragma unroll
        for (int j = 0; j < VEC; ++j) { sc[j] *= 0.5f; sf[j] *= 0.5f; }
    }
    const int p0 = blockIdx.x * ppb;
    int p1 = p0 + ppb; if (p1 > HW) p1 = HW;
    const int64_t base = (int64_t)b * HW * C + c0;
    const int first = p0 + threadIdx.y;
    const int niter = first < p1 ? (p1 - first + ty - 1) / ty : 0;
#pragma unroll
    for (int s = 0; s < D; ++s) {
        if (s < niter) {
            const int64_t off = base + (int64_t)(first + s * ty) * C;
            cp_async16(&ring[(s * 2 + 0) * nthr + tid], x + off);
            cp_async16(&ring[(s * 2 + 1) * nthr + tid], dy + off);
        }
        cp_async_commit();
    }
    int slot = 0;
    for (int it = 0; it < niter; ++it) {
        cp_async_wait<D - 1>();                                   // the group of iteration `it` has landed
        const uint4 ux = ring[(slot * 2 + 0) * nthr + tid];
        const uint4 ug = ring[(slot * 2 + 1) * nthr + tid];
        float v[VEC], g[VEC];
        unpack8(ux, v); unpack8(ug, g);
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            float ds = g[j];
            if (silu) ds *= silu_grad2_of_half(fmaf(v[j], sc[j], sf[j]));
            a[j] += ds;
            q[j] = fmaf(ds, v[j], q[j]);
        }
        if (it + D < niter) {                                     // refill the slot just consumed (its values are in registers)
            const int64_t off = base + (int64_t)(first + (it + D) * ty) * C;
            cp_async16(&ring[(slot * 2 + 0) * nthr + tid], x + off);
            cp_async16(&ring[(slot * 2 + 1) * nthr + tid], dy + off);
        }
        cp_async_commit();
        slot = (slot + 1 == D) ? 0 : slot + 1;
    }
    cp_async_wait<0>();
    if (silu) {
#pragma unroll
        for (int j = 0; j < VEC; ++j) { a[j] *= 0.5f; q[j] *= 0.5f; }
    }
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
        int g = (c0 + j) / cg;
        const float mean = stats[((int64_t)b * G + g) * 2];
        const float rstd = stats[((int64_t)b * G + g) * 2 + 1];
        q[j] = rstd * (q[j] - mean * a[j]);
    }
    __syncthreads();                                              // every thread is done with its ring before the reuse
    const int row = tx * VEC;
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
        sh[threadIdx.y * row + c0 + j] = (double)a[j];
        sh[(ty + threadIdx.y) * row + c0 + j] = (double)q[j];
    }
    __syncthreads();
    for (int c = tid; c < C; c += nthr) {
        double sa = 0.0, sq = 0.0;
        for (int y = 0; y < ty; ++y) { sa += sh[y * row + c]; sq += sh[(ty + y) * row + c]; }
        atomicAdd(&part[((int64_t)b * C + c) * 2 + 0], sa);
        atomicAdd(&part[((int64_t)b * C + c) * 2 + 1], sq);
    }
}

template <int D>
__global__ void gn_stats_async_kernel(const bf16* __restrict__ x, double* __restrict__ sums, int HW, int C, int G, int ppb) {
    constexpr int VEC = 8, NU = 2;                               // accumulation unit = 4 consecutive channels
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint4* ring = reinterpret_cast<uint4*>(smem_raw);            // [D][nthreads]
    double* sh = reinterpret_cast<double*>(smem_raw);            // reused after the loop: [2][ty][tx*NU]
    const int b = blockIdx.y;
    const int c0 = threadIdx.x * VEC;
    const int tx = blockDim.x, ty = blockDim.y, nthr = tx * ty;
    const int tid = threadIdx.y * tx + threadIdx.x;
    const int p0 = blockIdx.x * ppb;
    int p1 = p0 + ppb; if (p1 > HW) p1 = HW;
    const int64_t base = (int64_t)b * HW * C + c0;
    const int first = p0 + threadIdx.y;
    const int niter = first < p1 ? (p1 - first + ty - 1) / ty : 0;
    float s[NU] = {0.f, 0.f}, ss[NU] = {0.f, 0.f};
#pragma unroll
    for (int st = 0; st < D; ++st) {
        if (st < niter) cp_async16(&ring[st * nthr + tid], x + base + (int64_t)(first + st * ty) * C);
        cp_async_commit();
    }
    int slot = 0;
    for (int it = 0; it < niter; ++it) {
        cp_async_wait<D - 1>();
        const uint4 ux = ring[slot * nthr + tid];
        float v[VEC];
        unpack8(ux, v);
#pragma unroll
        for (int j = 0; j < VEC; ++j) { s[j / 4] += v[j]; ss[j / 4] = fmaf(v[j], v[j], ss[j / 4]); }
        if (it + D < niter) cp_async16(&ring[slot * nthr + tid], x + base + (int64_t)(first + (it + D) * ty) * C);
        cp_async_commit();
        slot = (slot + 1 == D) ? 0 : slot + 1;
    }
    cp_async_wait<0>();
    __syncthreads();
    const int row = tx * NU;
#pragma unroll
    for (int u = 0; u < NU; ++u) {
        sh[threadIdx.y * row + threadIdx.x * NU + u] = (double)s[u];
        sh[(ty + threadIdx.y) * row + threadIdx.x * NU + u] = (double)ss[u];
    }
    __syncthreads();
    const int upg = (C / G) / 4;
    for (int g = tid; g < G; g += nthr) {
        double a = 0.0, q = 0.0;
        for (int y = 0; y < ty; ++y)
            for (int t = 0; t < upg; ++t) {
                a += sh[y * row + g * upg + t];
                q += sh[(ty + y) * row + g * upg + t];
            }
        atomicAdd(&sums[((int64_t)b * G + g) * 2 + 0], a);
        atomicAdd(&sums[((int64_t)b * G + g) * 2 + 1], q);
    }
}

// forward apply, bf16 -> bf16, VEC = 8: x through the cp.async ring, y stored directly (stores are fire-and-forget)
template <int D>
__global__ void gn_apply_async_kernel(const bf16* __restrict__ x, const float* __restrict__ stats, const float* __restrict__ gamma,
                                      const float* __restrict__ beta, bf16* __restrict__ y, int HW, int C, int G, int ppb, int act,
                                      const double* __restrict__ sums, float* __restrict__ stats_out, double nel, double eps) {
    constexpr int VEC = 8;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint4* ring = reinterpret_cast<uint4*>(smem_raw);            // [D][nthreads]
    const int b = blockIdx.y;
    const int c0 = threadIdx.x * VEC;
    const int cg = C / G;
    const int tx = blockDim.x, ty = blockDim.y, nthr = tx * ty;
    const int tid = threadIdx.y * tx + threadIdx.x;
    const bool silu = (act == VQB_ACT_SILU);
    float sc[VEC], sh_[VEC];
    {
        float mean[VEC], rstd[VEC];
        gn_thread_stats<VEC>(stats, sums, stats_out, blockIdx.x == 0 && threadIdx.y == 0, b, c0, cg, G, nel, eps, mean, rstd);
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            sc[j] = rstd[j] * gamma[c0 + j];
            sh_[j] = beta[c0 + j] - mean[j] * sc[j];
            if (silu) { sc[j] *= 0.5f; sh_[j] *= 0.5f; }
        }
    }
    const int p0 = blockIdx.x * ppb;
    int p1 = p0 + ppb; if (p1 > HW) p1 = HW;
    const int64_t base = (int64_t)b * HW * C + c0;
    const int first = p0 + threadIdx.y;
    const int niter = first < p1 ? (p1 - first + ty - 1) / ty : 0;
#pragma unroll
    for (int st = 0; st < D; ++st) {
        if (st < niter) cp_async16(&ring[st * nthr + tid], x + base + (int64_t)(first + st * ty) * C);
        cp_async_commit();
    }
    int slot = 0;
    for (int it = 0; it < niter; ++it) {
        cp_async_wait<D - 1>();
        const uint4 ux = ring[slot * nthr + tid];
        float v[VEC], o[VEC];
        unpack8(ux, v);
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            float t = fmaf(v[j], sc[j], sh_[j]);
            o[j] = silu ? silu_of_half(t) : t;
        }
        stv<bf16, VEC>(y + base + (int64_t)(first + it * ty) * C, o);
        if (it + D < niter) cp_async16(&ring[slot * nthr + tid], x + base + (int64_t)(first + (it + D) * ty) * C);
        cp_async_commit();
        slot = (slot + 1 == D) ? 0 : slot + 1;
    }
    cp_async_wait<0>();
}

// backward apply, all-bf16, VEC = 8: x, dy (and the skip gradient `add`) through the cp.async ring
template <int D, bool HAS_ADD>
__global__ void gn_bwd_apply_async_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy, const float* __restrict__ stats,
                                          const float* __restrict__ gamma, const float* __restrict__ beta,
                                          const float* __restrict__ coef, const bf16* __restrict__ add, bf16* __restrict__ dx,
                                          int HW, int C, int G, int ppb, int act, const double* __restrict__ part,
                                          float* __restrict__ dgamma, float* __restrict__ dbeta, int acc, int N, double nel) {
    constexpr int VEC = 8, NT = HAS_ADD ? 3 : 2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint4* ring = reinterpret_cast<uint4*>(smem_raw);            // [D][NT][nthreads]
    const int b = blockIdx.y;
    const int c0 = threadIdx.x * VEC;
    const int cg = C / G;
    const int tx = blockDim.x, ty = blockDim.y, nthr = tx * ty;
    const int tid = threadIdx.y * tx + threadIdx.x;
    const bool silu = (act == VQB_ACT_SILU);
    float sc[VEC], sf[VEC], ca[VEC], cb[VEC], cc[VEC];       // dx = ds*ca + x*cb + cc
    if (part && dgamma && blockIdx.x == 0 && blockIdx.y == 0) gn_param_grads(part, dgamma, dbeta, acc, N, C, tid, nthr);
    {
        float k1 = 0.f, k2 = 0.f;
        int gprev = -1;
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            int g = (c0 + j) / cg;
            const float mean = stats[((int64_t)b * G + g) * 2];
            const float rstd = stats[((int64_t)b * G + g) * 2 + 1];
            if (g != gprev) { gn_group_coef(coef, part, gamma, b, g, cg, C, G, nel, k1, k2); gprev = g; }
            sc[j] = rstd * gamma[c0 + j];
            sf[j] = beta[c0 + j] - mean * sc[j];
            ca[j] = sc[j];
            cb[j] = -rstd * rstd * k2;
            cc[j] = -rstd * k1 - mean * cb[j];
            if (silu) { sc[j] *= 0.5f; sf[j] *= 0.5f; ca[j] *= 0.5f; }
        }
    }
    const int p0 = blockIdx.x * ppb;
    int p1 = p0 + ppb; if (p1 > HW) p1 = HW;
    const int64_t base = (int64_t)b * HW * C + c0;
    const int first = p0 + threadIdx.y;
    const int niter = first < p1 ? (p1 - first + ty - 1) / ty : 0;
    auto issue = [&](int slot, int it) {
        const int64_t off = base + (int64_t)(first + it * ty) * C;
        cp_async16(&ring[(slot * NT + 0) * nthr + tid], x + off);
        cp_async16(&ring[(slot * NT + 1) * nthr + tid], dy + off);
        if constexpr (HAS_ADD) cp_async16(&ring[(slot * NT + 2) * nthr + tid], add + off);
    };
#pragma unroll
    for (int st = 0; st < D; ++st) {
        if (st < niter) issue(st, st);
        cp_async_commit();
    }
    int slot = 0;
    for (int it = 0; it < niter; ++it) {
        cp_async_wait<D - 1>();
        float v[VEC], g[VEC], o[VEC];
        unpack8(ring[(slot * NT + 0) * nthr + tid], v);
        unpack8(ring[(slot * NT + 1) * nthr + tid], g);
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            float ds = g[j];
            if (silu) ds *= silu_grad2_of_half(fmaf(v[j], sc[j], sf[j]));
            o[j] = fmaf(ds, ca[j], fmaf(v[j], cb[j], cc[j]));
        }
        if constexpr (HAS_ADD) {
            float r[VEC];
            unpack8(ring[(slot * NT + 2) * nthr + tid], r);
#pragma unroll
            for (int j = 0; j < VEC; ++j) o[j] += r[j];
        }
        stv<bf16, VEC>(dx + base + (int64_t)(first + it * ty) * C, o);
        if (it + D < niter) issue(slot, it + D);
        cp_async_commit();
        slot = (slot + 1 == D) ? 0 : slot + 1;
    }
    cp_async_wait<0>();
}


// ---- backward, ONE cooperative launch: reduce + apply with the second read of x / dy served by L2 ------------------------------
// The two-launch backward reads x and dy twice from HBM (reduce: 2T, apply: 2T + skip gradient + dx) because a whole batch
// (2 x 1 GB at 256^2 x 128 channels x 64 images) passes between the two uses of an image.  Here a persistent grid walks the batch
// image by image: every CTA reduces ITS pixel range of image b (fp32 partials -> double atomics into part[b][c]), signals a
// per-image counter, then applies image b - 1 -- whose x / dy (33 MB) it streamed one phase earlier and are still resident in the
// 126 MB L2 (the apply phase walks its range in REVERSE order, most recently used lines first).  The grid-wide dependency is a
// counter per image (release: __threadfence + atomicAdd; acquire: one thread spins, then __syncthreads); the launch is cooperative
// (all CTAs co-resident by contract).  Same arithmetic as gn_bwd_reduce_async_kernel + gn_bwd_apply_async_kernel with the finalize
// folded in (gn_group_coef, gn_param_grads).
template <int D, bool HAS_ADD>
__global__ void __launch_bounds__(256) gn_bwd_fused_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy, const float* __restrict__ stats,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta,
                                                           double* __restrict__ part, int* __restrict__ counters, const bf16* __restrict__ add,
                                                           bf16* __restrict__ dx, float* __restrict__ dgamma, float* __restrict__ dbeta, int acc,
                                                           int N, int HW, int C, int G, int act, double nel) {
    constexpr int VEC = 8, NT = HAS_ADD ? 3 : 2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint4* ring = reinterpret_cast<uint4*>(smem_raw);            // [D][NT][nthreads]
    double* sh = reinterpret_cast<double*>(smem_raw);            // block reduction (the ring is drained at that point): [2][ty][tx*VEC]
    const int tx = blockDim.x, ty = blockDim.y, nthr = tx * ty;
    const int tid = threadIdx.y * tx + threadIdx.x;
    const int c0 = threadIdx.x * VEC;
    const int cg = C / G;
    const bool silu = (act == VQB_ACT_SILU);
    // this CTA's pixel range of EVERY image
    const int per = (HW + gridDim.x - 1) / gridDim.x;
    const int p0 = blockIdx.x * per;
    int p1 = p0 + per; if (p1 > HW) p1 = HW;
    const int first = p0 + threadIdx.y;
    const int niter = first < p1 ? (p1 - first + ty - 1) / ty : 0;
    float gam[VEC], bet[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) { gam[j] = gamma[c0 + j]; bet[j] = beta[c0 + j]; }

    for (int b = 0; b <= N; ++b) {
        if (b < N) {
            // ================= reduce image b =================
            float sc[VEC], sf[VEC], a[VEC], q[VEC], mean[VEC], rstd[VEC];
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
                const int g = (c0 + j) / cg;
                mean[j] = stats[((int64_t)b * G + g) * 2];
                rstd[j] = stats[((int64_t)b * G + g) * 2 + 1];
                sc[j] = rstd[j] * gam[j];
                sf[j] = bet[j] - mean[j] * sc[j];
                if (silu) { sc[j] *= 0.5f; sf[j] *= 0.5f; }
                a[j] = 0.f; q[j] = 0.f;
            }
            const int64_t base = (int64_t)b * HW * C + c0;
#pragma unroll
            for (int s = 0; s < D; ++s) {
                if (s < niter) {
                    const int64_t off = base + (int64_t)(first + s * ty) * C;
                    cp_async16(&ring[(s * NT + 0) * nthr + tid], x + off);
                    cp_async16(&ring[(s * NT + 1) * nthr + tid], dy + off);
                }
                cp_async_commit();
            }
            int slot = 0;
            for (int it = 0; it < niter; ++it) {
                cp_async_wait<D - 1>();
                float v[VEC], g[VEC];
                unpack8(ring[(slot * NT + 0) * nthr + tid], v);
                unpack8(ring[(slot * NT + 1) * nthr + tid], g);
#pragma unroll
                for (int j = 0; j < VEC; ++j) {
                    float ds = g[j];
                    if (silu) ds *= silu_grad2_of_half(fmaf(v[j], sc[j], sf[j]));
                    a[j] += ds;
                    q[j] = fmaf(ds, v[j], q[j]);
                }
                if (it + D < niter) {
                    const int64_t off = base + (int64_t)(first + (it + D) * ty) * C;
                    cp_async16(&ring[(slot * NT + 0) * nthr + tid], x + off);
                    cp_async16(&ring[(slot * NT + 1) * nthr + tid], dy + off);
                }
                cp_async_commit();
                slot = (slot + 1 == D) ? 0 : slot + 1;
            }
            cp_async_wait<0>();
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
                if (silu) { a[j] *= 0.5f; q[j] *= 0.5f; }
                q[j] = rstd[j] * (q[j] - mean[j] * a[j]);
            }
            __syncthreads();                                      // every thread is done with its ring before the reuse
            const int row = tx * VEC;
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
                sh[threadIdx.y * row + c0 + j] = (double)a[j];
                sh[(ty + threadIdx.y) * row + c0 + j] = (double)q[j];
            }
            __syncthreads();
            for (int c = tid; c < C; c += nthr) {
                double sa = 0.0, sq = 0.0;
                for (int y = 0; y < ty; ++y) { sa += sh[y * row + c]; sq += sh[(ty + y) * row + c]; }
                atomicAdd(&part[((int64_t)b * C + c) * 2 + 0], sa);
                atomicAdd(&part[((int64_t)b * C + c) * 2 + 1], sq);
            }
            __threadfence();                                      // the sums are visible device-wide before the counter moves
            __syncthreads();
            if (tid == 0) atomicAdd(&counters[b], 1);
        }
        if (b >= 1) {
            // ================= apply image b - 1 =================
            const int ib = b - 1;
            if (tid == 0) {
                const volatile int* cnt = counters + ib;
                while (*cnt < (int)gridDim.x) { __nanosleep(64); }
                __threadfence();
            }
            __syncthreads();
            if (ib == N - 1 && blockIdx.x == 0 && dgamma) gn_param_grads(part, dgamma, dbeta, acc, N, C, tid, nthr);
            float sc[VEC], sf[VEC], ca[VEC], cb[VEC], cc[VEC];       // dx = ds*ca + x*cb + cc
            {
                float k1 = 0.f, k2 = 0.f;
                int gprev = -1;
#pragma unroll
                for (int j = 0; j < VEC; ++j) {
                    const int g = (c0 + j) / cg;
                    const float mean = stats[((int64_t)ib * G + g) * 2];
                    const float rstd = stats[((int64_t)ib * G + g) * 2 + 1];
                    if (g != gprev) { gn_group_coef(nullptr, part, gamma, ib, g, cg, C, G, nel, k1, k2); gprev = g; }
                    sc[j] = rstd * gam[j];
                    sf[j] = bet[j] - mean * sc[j];
                    ca[j] = sc[j];
                    cb[j] = -rstd * rstd * k2;
                    cc[j] = -rstd * k1 - mean * cb[j];
                    if (silu) { sc[j] *= 0.5f; sf[j] *= 0.5f; ca[j] *= 0.5f; }
                }
            }
            const int64_t base = (int64_t)ib * HW * C + c0;
            // iteration `it` handles pixel first + (niter - 1 - it) * ty: the range is walked backwards
            auto issue = [&](int slot, int it) {
                const int64_t off = base + (int64_t)(first + (niter - 1 - it) * ty) * C;
                cp_async16(&ring[(slot * NT + 0) * nthr + tid], x + off);
                cp_async16(&ring[(slot * NT + 1) * nthr + tid], dy + off);
                if constexpr (HAS_ADD) cp_async16(&ring[(slot * NT + 2) * nthr + tid], add + off);
            };
#pragma unroll
            for (int st = 0; st < D; ++st) {
                if (st < niter) issue(st, st);
                cp_async_commit();
            }
            int slot = 0;
            for (int it = 0; it < niter; ++it) {
                cp_async_wait<D - 1>();
                float v[VEC], g[VEC], o[VEC];
                unpack8(ring[(slot * NT + 0) * nthr + tid], v);
                unpack8(ring[(slot * NT + 1) * nthr + tid], g);
#pragma unroll
                for (int j = 0; j < VEC; ++j) {
                    float ds = g[j];
                    if (silu) ds *= silu_grad2_of_half(fmaf(v[j], sc[j], sf[j]));
                    o[j] = fmaf(ds, ca[j], fmaf(v[j], cb[j], cc[j]));
                }
                if constexpr (HAS_ADD) {
                    float r[VEC];
                    unpack8(ring[(slot * NT + 2) * nthr + tid], r);
#pragma unroll
                    for (int j = 0; j < VEC; ++j) o[j] += r[j];
                }
                stv<bf16, VEC>(dx + base + (int64_t)(first + (niter - 1 - it) * ty) * C, o);
                if (it + D < niter) issue(slot, it + D);
                cp_async_commit();
                slot = (slot + 1 == D) ? 0 : slot + 1;
            }
            cp_async_wait<0>();
            __syncthreads();                                      // ring / sh are reused by the next reduce phase
        }
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
                                    const float* __restrict__ coef, const TO* __restrict__ add, TO* __restrict__ dx, int HW,
                                    int C, int G, int ppb, int act, const double* __restrict__ part, float* __restrict__ dgamma,
                                    float* __restrict__ dbeta, int acc, int N, double nel) {
    const int b = blockIdx.y;
    const int c0 = threadIdx.x * VEC;
    const int cg = C / G;
    float sc[VEC], sf[VEC], ca[VEC], cb[VEC], cc[VEC];       // dx = ds*ca + x*cb + cc
    if (part && dgamma && blockIdx.x == 0 && blockIdx.y == 0)
        gn_param_grads(part, dgamma, dbeta, acc, N, C, threadIdx.y * blockDim.x + threadIdx.x, blockDim.x * blockDim.y);
    float k1 = 0.f, k2 = 0.f;
    int gprev = -1;
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
        int g = (c0 + j) / cg;
        const float mean = stats[((int64_t)b * G + g) * 2];
        const float rstd = stats[((int64_t)b * G + g) * 2 + 1];
        if (g != gprev) { gn_group_coef(coef, part, gamma, b, g, cg, C, G, nel, k1, k2); gprev = g; }
        sc[j] = rstd * gamma[c0 + j];
        sf[j] = beta[c0 + j] - mean * sc[j];
        ca[j] = sc[j];                                    // rstd * gamma
        cb[j] = -rstd * rstd * k2;                        // -(x-mean)*rstd^2*k2
        cc[j] = -rstd * k1 - mean * cb[j];
    }
    constexpr bool APPROX = (sizeof(TI) == 2 && sizeof(TG) == 2);
    if (APPROX && act == VQB_ACT_SILU) {                  // ds arrives doubled (silu_grad2_of_half of u = t/2)
#pragma unroll
        for (int j = 0; j < VEC; ++j) { sc[j] *= 0.5f; sf[j] *= 0.5f; ca[j] *= 0.5f; }
    }
    const int p0 = blockIdx.x * ppb;
    int p1 = p0 + ppb; if (p1 > HW) p1 = HW;
    const int64_t base = (int64_t)b * HW * C + c0;
#pragma unroll 2
    for (int p = p0 + threadIdx.y; p < p1; p += blockDim.y) {
        float v[VEC], g[VEC], o[VEC];
        ldv<TI, VEC>(x + base + (int64_t)p * C, v);
        ldv<TG, VEC>(dy + base + (int64_t)p * C, g);
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            float ds = g[j];
            if constexpr (APPROX) { if (act == VQB_ACT_SILU) ds *= silu_grad2_of_half(fmaf(v[j], sc[j], sf[j])); }
            else { if (act == VQB_ACT_SILU) ds *= fast_silu_grad<false>(fmaf(v[j], sc[j], sf[j])); }
            o[j] = fmaf(ds, ca[j], fmaf(v[j], cb[j], cc[j]));
        }
        if (add) {                                    // fused accumulation of the skip-connection gradient
            float r[VEC];
            ldv<TO, VEC>(add + base + (int64_t)p * C, r);
#pragma unroll
            for (int j = 0; j < VEC; ++j) o[j] += r[j];
        }
        stv<TO, VEC>(dx + base + (int64_t)p * C, o);
    }
}

inline int gn_check(const char* name, int N, int HW, int C, int G) {
    if (!(N > 0 && HW > 0 && C > 0 && G > 0 && C % G == 0)) {
        vqb_set_error("%s: bad shape N=%d HW=%d C=%d G=%d", name, N, HW, C, G);
        return VQB_ERR_ARG;
    }
    if (C / gn_vec(C, G) > 1024) { vqb_set_error("%s: C=%d too large", name, C); return VQB_ERR_UNSUPPORTED; }
    if ((double)(C / G) * HW < 2.0) { vqb_set_error("%s: unbiased variance needs >= 2 elements per group", name); return VQB_ERR_ARG; }
    return VQB_OK;
}

}  // namespace

#define GN_VEC_DISPATCH(L, ...)                                     \
    if ((L).vec == 8) { constexpr int VEC = 8; __VA_ARGS__ }        \
    else if ((L).vec == 4) { constexpr int VEC = 4; __VA_ARGS__ }   \
    else { constexpr int VEC = 1; __VA_ARGS__ }

extern "C" int vqb_gn_stats(const void* x, int x_dtype, double* sums, int N, int HW, int C, int G, void* stream) {
    int rc = gn_check("gn_stats", N, HW, C, G); if (rc) return rc;
    VQB_CHECK_ARG(x && sums, "gn_stats: null pointer");
    static int use_async = getenv("VQB_GN_ASYNC") ? atoi(getenv("VQB_GN_ASYNC")) : 1;
    if (use_async && x_dtype == VQB_BF16 && gn_vec(C, G) == 8) {
        const int ppt = 128;                          // pixels per thread (tools/gn_probe.py sweep)
        GnLaunch La = gn_launch(N, HW, C, G, ppt, 8);
        const size_t nthr = (size_t)La.block.x * La.block.y;
        size_t ring = (size_t)8 * nthr * 16, red = 2 * sizeof(double) * nthr * 2;
        gn_stats_async_kernel<8><<<La.grid, La.block, ring > red ? ring : red, as_stream(stream)>>>((const bf16*)x, sums, HW, C, G, La.ppb);
        VQB_CHECK_LAUNCH("gn_stats_async");
        return VQB_OK;
    }
    GnLaunch L = gn_launch(N, HW, C, G, 32);
    const int nu = (L.vec >= 4) ? L.vec / 4 : 1;
    size_t sm = 2 * sizeof(double) * L.block.x * L.block.y * nu;
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

static int gn_apply_impl(const void* x, int x_dtype, const float* stats, const double* sums, float* stats_out, float eps,
                         const float* gamma, const float* beta, void* y, int y_dtype, int N, int HW, int C, int G, int act, void* stream) {
    int rc = gn_check("gn_apply", N, HW, C, G); if (rc) return rc;
    VQB_CHECK_ARG(x && (stats || sums) && gamma && beta && y, "gn_apply: null pointer");
    VQB_CHECK_ARG(act == VQB_ACT_NONE || act == VQB_ACT_SILU, "gn_apply: act must be NONE or SILU");
    const double nel = (double)(C / G) * HW;
    static int use_async = getenv("VQB_GN_ASYNC") ? atoi(getenv("VQB_GN_ASYNC")) : 1;
    if (use_async && x_dtype == VQB_BF16 && y_dtype == VQB_BF16 && gn_vec(C, G) == 8) {
        const int ppt = 32;
        GnLaunch La = gn_launch(N, HW, C, G, ppt, 8);
        const size_t nthr = (size_t)La.block.x * La.block.y;
        gn_apply_async_kernel<8><<<La.grid, La.block, (size_t)8 * nthr * 16, as_stream(stream)>>>((const bf16*)x, stats, gamma, beta, (bf16*)y, HW, C, G, La.ppb, act,
                                                                                                 sums, stats_out, nel, (double)eps);
        VQB_CHECK_LAUNCH("gn_apply_async");
        return VQB_OK;
    }
    GnLaunch L = gn_launch(N, HW, C, G);
    GN_VEC_DISPATCH(L, VQB_DISPATCH_1(x_dtype, TI, VQB_DISPATCH_1(y_dtype, TO,
        (gn_apply_kernel<TI, TO, VEC><<<L.grid, L.block, 0, as_stream(stream)>>>((const TI*)x, stats, gamma, beta, (TO*)y, HW, C, G, L.ppb, act,
                                                                                  sums, stats_out, nel, (double)eps));)))
    VQB_CHECK_LAUNCH("gn_apply");
    return VQB_OK;
}

extern "C" int vqb_gn_apply(const void* x, int x_dtype, const float* stats, const float* gamma, const float* beta, void* y,
                            int y_dtype, int N, int HW, int C, int G, int act, void* stream) {
    VQB_CHECK_ARG(stats, "gn_apply: null pointer");
    return gn_apply_impl(x, x_dtype, stats, nullptr, nullptr, 0.f, gamma, beta, y, y_dtype, N, HW, C, G, act, stream);
}

extern "C" int vqb_gn_apply_sums(const void* x, int x_dtype, const double* sums, const float* gamma, const float* beta, void* y, int y_dtype,
                                 float* stats_out, int N, int HW, int C, int G, float eps, int act, void* stream) {
    VQB_CHECK_ARG(sums, "gn_apply_sums: null pointer");
    return gn_apply_impl(x, x_dtype, nullptr, sums, stats_out, eps, gamma, beta, y, y_dtype, N, HW, C, G, act, stream);
}

extern "C" int vqb_gn_bwd_reduce(const void* x, int x_dtype, const void* dy, int dy_dtype, const float* stats,
                                 const float* gamma, const float* beta, double* part, int N, int HW, int C, int G, int act,
                                 void* stream) {
    int rc = gn_check("gn_bwd_reduce", N, HW, C, G); if (rc) return rc;
    VQB_CHECK_ARG(x && dy && stats && gamma && beta && part, "gn_bwd_reduce: null pointer");
    static int use_async = getenv("VQB_GN_ASYNC") ? atoi(getenv("VQB_GN_ASYNC")) : 1;
    if (use_async && x_dtype == VQB_BF16 && dy_dtype == VQB_BF16 && gn_vec(C, G) == 8 && (act == VQB_ACT_NONE || act == VQB_ACT_SILU)) {
        const int ppt = 128;
        GnLaunch L = gn_launch(N, HW, C, G, ppt, 8);
        const size_t nthr = (size_t)L.block.x * L.block.y;
        const int depth = GN_ASYNC_DEPTH;
        size_t ring = (size_t)depth * 2 * nthr * 16, red = 2 * sizeof(double) * nthr * 8;
        size_t sm = ring > red ? ring : red;
        static bool attr_set = false;
        if (!attr_set) {
            VQB_CUDA(cudaFuncSetAttribute(gn_bwd_reduce_async_kernel<GN_ASYNC_DEPTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
            attr_set = true;
        }
        gn_bwd_reduce_async_kernel<GN_ASYNC_DEPTH><<<L.grid, L.block, sm, as_stream(stream)>>>((const bf16*)x, (const bf16*)dy, stats, gamma, beta,
                                                                                             part, HW, C, G, L.ppb, act);
        VQB_CHECK_LAUNCH("gn_bwd_reduce_async");
        return VQB_OK;
    }
    GnLaunch L = gn_launch(N, HW, C, G, 128, 8);
    size_t sm = 2 * sizeof(double) * L.block.x * L.block.y * L.vec;
    GN_VEC_DISPATCH(L, VQB_DISPATCH_1(x_dtype, TI, VQB_DISPATCH_1(dy_dtype, TG,
        (gn_bwd_reduce_kernel<TI, TG, VEC, 4><<<L.grid, L.block, sm, as_stream(stream)>>>((const TI*)x, (const TG*)dy, stats, gamma, beta, part, HW, C, G, L.ppb, act));)))
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

static int gn_bwd_apply_impl(const void* x, int x_dtype, const void* dy, int dy_dtype, const float* stats, const float* gamma,
                             const float* beta, const float* coef, const double* part, const void* add, void* dx, int dx_dtype,
                             float* dgamma, float* dbeta, int acc, int N, int HW, int C, int G, int act, void* stream) {
    int rc = gn_check("gn_bwd_apply", N, HW, C, G); if (rc) return rc;
    VQB_CHECK_ARG(x && dy && stats && gamma && beta && (coef || part) && dx, "gn_bwd_apply: null pointer");
    VQB_CHECK_ARG(!part || (dgamma && dbeta), "gn_bwd_apply_part: null parameter-gradient pointer");
    const double nel = (double)(C / G) * HW;
    static int use_async = getenv("VQB_GN_ASYNC") ? atoi(getenv("VQB_GN_ASYNC")) : 1;
    if (use_async && x_dtype == VQB_BF16 && dy_dtype == VQB_BF16 && dx_dtype == VQB_BF16 && gn_vec(C, G) == 8 &&
        (act == VQB_ACT_NONE || act == VQB_ACT_SILU)) {
        const int ppt = 16;
        GnLaunch La = gn_launch(N, HW, C, G, ppt, 8);
        const size_t nthr = (size_t)La.block.x * La.block.y;
        constexpr int DB = 6;
        static bool attr_set = false;
        if (!attr_set) {
            VQB_CUDA(cudaFuncSetAttribute(gn_bwd_apply_async_kernel<DB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
            VQB_CUDA(cudaFuncSetAttribute(gn_bwd_apply_async_kernel<DB, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
            attr_set = true;
        }
        if (add)
            gn_bwd_apply_async_kernel<DB, true><<<La.grid, La.block, (size_t)DB * 3 * nthr * 16, as_stream(stream)>>>(
                (const bf16*)x, (const bf16*)dy, stats, gamma, beta, coef, (const bf16*)add, (bf16*)dx, HW, C, G, La.ppb, act, part, dgamma, dbeta, acc, N, nel);
        else
            gn_bwd_apply_async_kernel<DB, false><<<La.grid, La.block, (size_t)DB * 2 * nthr * 16, as_stream(stream)>>>(
                (const bf16*)x, (const bf16*)dy, stats, gamma, beta, coef, nullptr, (bf16*)dx, HW, C, G, La.ppb, act, part, dgamma, dbeta, acc, N, nel);
        VQB_CHECK_LAUNCH("gn_bwd_apply_async");
        return VQB_OK;
    }
    GnLaunch L = gn_launch(N, HW, C, G);
    GN_VEC_DISPATCH(L, VQB_DISPATCH_1(x_dtype, TI, VQB_DISPATCH_1(dy_dtype, TG, VQB_DISPATCH_1(dx_dtype, TO,
        (gn_bwd_apply_kernel<TI, TG, TO, VEC><<<L.grid, L.block, 0, as_stream(stream)>>>((const TI*)x, (const TG*)dy, stats, gamma, beta, coef, (const TO*)add, (TO*)dx, HW, C, G, L.ppb, act,
                                                                                          part, dgamma, dbeta, acc, N, nel));))))
    VQB_CHECK_LAUNCH("gn_bwd_apply");
    return VQB_OK;
}

extern "C" int vqb_gn_bwd_apply(const void* x, int x_dtype, const void* dy, int dy_dtype, const float* stats,
                                const float* gamma, const float* beta, const float* coef, const void* add, void* dx,
                                int dx_dtype, int N, int HW, int C, int G, int act, void* stream) {
    VQB_CHECK_ARG(coef, "gn_bwd_apply: null pointer");
    return gn_bwd_apply_impl(x, x_dtype, dy, dy_dtype, stats, gamma, beta, coef, nullptr, add, dx, dx_dtype, nullptr, nullptr, 0, N, HW, C, G, act, stream);
}

extern "C" int vqb_gn_bwd_apply_part(const void* x, int x_dtype, const void* dy, int dy_dtype, const float* stats, const float* gamma,
                                     const float* beta, const double* part, const void* add, void* dx, int dx_dtype, float* dgamma,
                                     float* dbeta, int accumulate_param_grads, int N, int HW, int C, int G, int act, void* stream) {
    VQB_CHECK_ARG(part, "gn_bwd_apply_part: null pointer");
    return gn_bwd_apply_impl(x, x_dtype, dy, dy_dtype, stats, gamma, beta, nullptr, part, add, dx, dx_dtype, dgamma, dbeta, accumulate_param_grads,
                             N, HW, C, G, act, stream);
}

// One cooperative launch for the whole backward (see gn_bwd_fused_kernel).  part [N][C][2] doubles and counters [N] ints are
// zero-filled by the caller.  Returns VQB_ERR_UNSUPPORTED (nothing launched) when the problem does not qualify -- the caller then
// uses vqb_gn_bwd_reduce + vqb_gn_bwd_apply_part.
extern "C" int vqb_gn_bwd_fused_supported(int x_dtype, int dy_dtype, int dx_dtype, int N, int HW, int C, int G, int act) {
    // OFF by default -- a measured negative result (profiles/r02_ncu_full_gn_bwd_fused_raw.csv): the L2 reuse works for 17 MB images
    // (DRAM reads = x + dy exactly once) but not for the 33.5 MB images that carry most of the bytes (3.8 T read: the effective L2
    // capacity for this streaming pattern is < 45 MB), and the per-image phase switch (ring drain, block reduction, counter) leaves
    // DRAM 17-23 % busy: 1.16 ms per launch against 0.35 + 0.45 ms for the two-launch form.  VQB_GN_BWD_FUSED=1 enables it.
    static const int enabled = getenv("VQB_GN_BWD_FUSED") ? atoi(getenv("VQB_GN_BWD_FUSED")) : 0;
    if (!enabled || x_dtype != VQB_BF16 || dy_dtype != VQB_BF16 || dx_dtype != VQB_BF16) return 0;
    if (!(N > 1 && HW > 0 && C > 0 && G > 0 && C % G == 0) || gn_vec(C, G) != 8 || C / 8 > 256 || 256 % (C / 8) != 0) return 0;
    if (!(act == VQB_ACT_NONE || act == VQB_ACT_SILU)) return 0;
    // worth it when one image's x + dy is a sizeable piece of L2 (>= 12 MB) yet two of them fit comfortably (<= 40 MB each)
    const double img = 4.0 * HW * C;
    return img >= 12e6 && img <= 40e6;
}

extern "C" int vqb_gn_bwd_fused(const void* x, const void* dy, const float* stats, const float* gamma, const float* beta, double* part,
                                int* counters, const void* add, void* dx, float* dgamma, float* dbeta, int accumulate_param_grads, int N,
                                int HW, int C, int G, int act, int max_ctas, void* stream) {
    int rc = gn_check("gn_bwd_fused", N, HW, C, G); if (rc) return rc;
    VQB_CHECK_ARG(x && dy && stats && gamma && beta && part && counters && dx && dgamma && dbeta, "gn_bwd_fused: null pointer");
    if (!(N > 1 && gn_vec(C, G) == 8 && C / 8 <= 256 && 256 % (C / 8) == 0 && (act == VQB_ACT_NONE || act == VQB_ACT_SILU))) {
        vqb_set_error("gn_bwd_fused: unsupported problem (N=%d HW=%d C=%d G=%d)", N, HW, C, G);
        return VQB_ERR_UNSUPPORTED;
    }
    constexpr int DB = 6;
    const int tx = C / 8, ty = 256 / tx;
    dim3 block(tx, ty);
    const size_t nthr = 256;
    size_t ring = (size_t)DB * (add ? 3 : 2) * nthr * 16, red = 2 * sizeof(double) * nthr * 8;
    size_t smem = ring > red ? ring : red;
    static bool attr_set = false;
    if (!attr_set) {
        VQB_CUDA(cudaFuncSetAttribute(gn_bwd_fused_kernel<DB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        VQB_CUDA(cudaFuncSetAttribute(gn_bwd_fused_kernel<DB, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        attr_set = true;
    }
    int dev = 0, sms = 0, occ = 0;
    VQB_CUDA(cudaGetDevice(&dev));
    VQB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (add) { VQB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, gn_bwd_fused_kernel<DB, true>, 256, smem)); }
    else { VQB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, gn_bwd_fused_kernel<DB, false>, 256, smem)); }
    if (occ < 1) { vqb_set_error("gn_bwd_fused: kernel does not fit an SM"); return VQB_ERR_UNSUPPORTED; }
    if (occ > 2) occ = 2;
    int grid = sms * occ;
    // max_ctas > 0: leave SMs free for kernels that run BESIDE this one (the overlapped NCCL gradient all-reduce of a data-parallel
    // step): a cooperative grid starts only when all its CTAs fit at once, so a full-machine grid would wait for those to drain
    if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;
    const int max_grid = (HW + ty - 1) / ty;                      // at least one pixel row of threads per CTA
    if (grid > max_grid) grid = max_grid;
    const double nel = (double)(C / G) * HW;
    const bf16* xp = (const bf16*)x; const bf16* dyp = (const bf16*)dy; const bf16* addp = (const bf16*)add; bf16* dxp = (bf16*)dx;
    void* args[] = {&xp, &dyp, &stats, &gamma, &beta, &part, &counters, &addp, &dxp, &dgamma, &dbeta, &accumulate_param_grads,
                    &N, &HW, &C, &G, &act, (void*)&nel};
    const void* fn = add ? (const void*)gn_bwd_fused_kernel<DB, true> : (const void*)gn_bwd_fused_kernel<DB, false>;
    VQB_CUDA(cudaLaunchCooperativeKernel(fn, dim3(grid), block, args, smem, as_stream(stream)));
    VQB_CHECK_LAUNCH("gn_bwd_fused");
    return VQB_OK;
}
