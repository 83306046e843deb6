// Entropy-regularised and Gumbel-softmax quantizer row kernels (reference: vqvae/modules/vector_quantizers.py:206-356).
//
// Both quantizers need a per-row softmax over the K codes and are therefore organised around an N x K fp32 matrix
// (affinity / logits) that is produced and consumed by the implicit-GEMM kernels (1x1 convolutions) and transformed IN
// PLACE here: one warp per row, coalesced strided passes over K, warp-shuffle reductions, double-precision atomics for
// the scalar loss partials.  All kernels are HBM-bound (one or two passes over N x K x 4 bytes).
#include "common.cuh"

namespace {

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__global__ void sqnorm_kernel(const float* __restrict__ a, float* __restrict__ out, int64_t R, int D) {
    int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (row >= R) return;
    float s = 0.f;
    for (int d = lane; d < D; d += 32) { float v = a[row * D + d]; s = fmaf(v, v, s); }
    s = warp_sum(s);
    if (lane == 0) out[row] = s;
}

// ---- entropy quantizer, forward rows ---------------------------------------------------------------------
// in : m[row][k] = z_row . e_k ; out: m[row][k] = log softmax_k(-d/T), d = (|z|^2 - 2 dot) + |e_k|^2  (:337-340)
__global__ void entropy_rows_kernel(float* __restrict__ m, const float* __restrict__ z, const float* __restrict__ cb_sq,
                                    float inv_t, int64_t* __restrict__ idx_out, double* __restrict__ ent_sum, int64_t N,
                                    int K, int D, int argmax_target) {
    int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (row >= N) return;
    float a2 = 0.f;
    for (int d = lane; d < D; d += 32) { float v = z[row * D + d]; a2 = fmaf(v, v, a2); }
    a2 = warp_sum(a2);
    float* mr = m + row * K;
    // pass 1: distances (stored), argmin with first-index ties
    float best = INFINITY; int besti = 0x7fffffff;
    for (int k = lane; k < K; k += 32) {
        float dist = __fadd_rn(__fsub_rn(a2, 2.0f * mr[k]), cb_sq[k]);
        mr[k] = dist;
        if (dist < best) { best = dist; besti = k; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        float ov = __shfl_xor_sync(0xffffffffu, best, o);
        int oi = __shfl_xor_sync(0xffffffffu, besti, o);
        if (ov < best || (ov == best && oi < besti)) { best = ov; besti = oi; }
    }
    if (lane == 0) idx_out[row] = (besti >= 0 && besti < K) ? besti : 0;
    // pass 2: softmax statistics of a = -d/T (max a = -min d / T)
    const float amax = -best * inv_t;
    float s = 0.f, sa = 0.f;
    for (int k = lane; k < K; k += 32) {
        float a = -mr[k] * inv_t;
        float e = expf(a - amax);
        s += e;
        sa = fmaf(e, a - amax, sa);
    }
    s = warp_sum(s); sa = warp_sum(sa);
    const float lse = amax + logf(s);
    // softmax targets: -sum_k p_k log p_k = -(sa/s - log(s));  argmax targets (:311-315): -log p at the arg-max = log(s)
    if (lane == 0) atomicAdd(ent_sum, argmax_target ? (double)logf(s) : (double)(-(sa / s - logf(s))));
    // pass 3: log-probabilities in place
    for (int k = lane; k < K; k += 32) mr[k] = -mr[k] * inv_t - lse;
}

__global__ void colsum_exp_kernel(const float* __restrict__ logp, float* __restrict__ out, int64_t N, int K, int rows_per_block) {
    int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
    int64_t r1 = r0 + rows_per_block; if (r1 > N) r1 = N;
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        float s = 0.f;
        for (int64_t r = r0 + threadIdx.y; r < r1; r += blockDim.y) s += expf(logp[r * K + k]);
        atomicAdd(out + k, s);
    }
}

// out[0] = ratio * (mean sample entropy - avg entropy), out[1] = avg entropy   (single block)
__global__ void entropy_finalize_kernel(const float* __restrict__ colsum_p, const double* __restrict__ ent_sum, float ratio,
                                        float* __restrict__ out, double n, int K) {
    __shared__ double sh[32];
    double acc = 0.0;
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        double mk = (double)(colsum_p[k] / (float)n);
        acc -= mk * (double)logf((float)mk + 1e-5f);
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += sh[i];
        out[0] = (float)((double)ratio * (ent_sum[0] / n - t));
        out[1] = (float)t;
    }
}

// in: m = logp; out: m = G = dLoss/dd = -(1/T) p (g - sum_k p g),  g = c (-(logp+1) + log(mbar+eps) + mbar/(mbar+eps))
// argmax targets (idx != null; colsum_p holds the code histogram): the straight-through one-hot t carries the softmax
// Jacobian, and log p is differentiated with t as its weights:  G = -(1/T) [ p (g - sum_k p g) + c (p - t) ]
__global__ void entropy_bwd_rows_kernel(float* __restrict__ m, const float* __restrict__ colsum_p, const float* __restrict__ g_loss,
                                        float ratio, float inv_t, int64_t N, int K, const int64_t* __restrict__ idx) {
    int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (row >= N) return;
    const float c = (g_loss ? g_loss[0] : 1.0f) * ratio / (float)N;
    const float invn = 1.0f / (float)N;
    float* mr = m + row * K;
    float r = 0.f;
    for (int k = lane; k < K; k += 32) {
        float lp = mr[k], p = expf(lp);
        float mb = colsum_p[k] * invn;
        float g = c * (-(lp + 1.0f) + logf(mb + 1e-5f) + mb / (mb + 1e-5f));
        r = fmaf(p, g, r);
    }
    r = warp_sum(r);
    for (int k = lane; k < K; k += 32) {
        float lp = mr[k], p = expf(lp);
        float mb = colsum_p[k] * invn;
        float g = c * (-(lp + 1.0f) + logf(mb + 1e-5f) + mb / (mb + 1e-5f));
        float v = p * (g - r);
        if (idx) v += c * (p - (idx[row] == k ? 1.0f : 0.0f));
        mr[k] = -inv_t * v;
    }
}

// dcb (in: MSE part) += 2 E * colsum(G)[:,None] - 2 GtZ
__global__ void entropy_combine_dcb_kernel(float* __restrict__ dcb, const float* __restrict__ cb, const float* __restrict__ colsum_g,
                                           const float* __restrict__ gtz, int K, int D) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= (int64_t)K * D) return;
    int k = (int)(i / D);
    dcb[i] += 2.0f * cb[i] * colsum_g[k] - 2.0f * gtz[i];
}

// ---- Gumbel-softmax rows ----------------------------------------------------------------------------------
// y = softmax((logits - log(E)) / tau)  [hard: one-hot(argmax)], idx = argmax y, kl_sum += sum_n qy log(qy K + 1e-10), qy = softmax(logits)
__global__ void gumbel_rows_fwd_kernel(const float* __restrict__ logits, const float* __restrict__ expo, float inv_tau, int hard,
                                       float* __restrict__ y, int64_t* __restrict__ idx_out, double* __restrict__ kl_sum, int64_t N, int K,
                                       const float* __restrict__ tau_dev) {
    if (tau_dev) inv_tau = 1.0f / tau_dev[0];           // temperature from device memory (CUDA-graph replay with a schedule)
    int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (row >= N) return;
    const float* lr = logits + row * K;
    const float* er = expo ? expo + row * K : nullptr;
    float* yr = y + row * K;
    float mx1 = -INFINITY, mx2 = -INFINITY; int arg = 0x7fffffff;
    for (int k = lane; k < K; k += 32) {
        float l = lr[k];
        float t = (er ? (l - logf(er[k])) : l) * inv_tau;
        if (t > mx1) { mx1 = t; arg = k; }
        mx2 = fmaxf(mx2, l);
    }
    float m1 = mx1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        float ov = __shfl_xor_sync(0xffffffffu, m1, o);
        int oi = __shfl_xor_sync(0xffffffffu, arg, o);
        if (ov > m1 || (ov == m1 && oi < arg)) { m1 = ov; arg = oi; }
    }
    const float m2 = warp_max(mx2);
    float s1 = 0.f, s2 = 0.f;
    for (int k = lane; k < K; k += 32) {
        float l = lr[k];
        float t = (er ? (l - logf(er[k])) : l) * inv_tau;
        s1 += expf(t - m1);
        s2 += expf(l - m2);
    }
    s1 = warp_sum(s1); s2 = warp_sum(s2);
    const float inv1 = 1.0f / s1, inv2 = 1.0f / s2;
    float kl = 0.f;
    for (int k = lane; k < K; k += 32) {
        float l = lr[k];
        float t = (er ? (l - logf(er[k])) : l) * inv_tau;
        float soft = expf(t - m1) * inv1;
        yr[k] = hard ? ((k == arg) ? 1.0f : 0.0f) : soft;
        float qy = expf(l - m2) * inv2;
        kl = fmaf(qy, logf(qy * (float)K + 1e-10f), kl);
    }
    kl = warp_sum(kl);
    if (lane == 0) {
        idx_out[row] = arg;
        atomicAdd(kl_sum, (double)kl);
    }
}

// dlogits = (1/tau) soft (dy - sum soft dy) + ckl * qy (f - sum qy f),  f = log(qy K + eps) + qy K / (qy K + eps)
__global__ void gumbel_rows_bwd_kernel(const float* __restrict__ logits, const float* __restrict__ expo, float inv_tau,
                                       const float* __restrict__ dy, const float* __restrict__ g_kl, float kl_scale,
                                       float* __restrict__ dlogits, int64_t N, int K, const float* __restrict__ tau_dev) {
    if (tau_dev) inv_tau = 1.0f / tau_dev[0];
    int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (row >= N) return;
    const float* lr = logits + row * K;
    const float* er = expo ? expo + row * K : nullptr;
    const float* gr = dy ? dy + row * K : nullptr;
    float* dr = dlogits + row * K;
    const float ckl = (g_kl ? g_kl[0] : 0.0f) * kl_scale;
    float mx1 = -INFINITY, mx2 = -INFINITY;
    for (int k = lane; k < K; k += 32) {
        float l = lr[k];
        mx1 = fmaxf(mx1, (er ? (l - logf(er[k])) : l) * inv_tau);
        mx2 = fmaxf(mx2, l);
    }
    const float m1 = warp_max(mx1), m2 = warp_max(mx2);
    float s1 = 0.f, s2 = 0.f;
    for (int k = lane; k < K; k += 32) {
        float l = lr[k];
        s1 += expf((er ? (l - logf(er[k])) : l) * inv_tau - m1);
        s2 += expf(l - m2);
    }
    s1 = warp_sum(s1); s2 = warp_sum(s2);
    const float inv1 = 1.0f / s1, inv2 = 1.0f / s2;
    float a = 0.f, b = 0.f;
    for (int k = lane; k < K; k += 32) {
        float l = lr[k];
        float soft = expf((er ? (l - logf(er[k])) : l) * inv_tau - m1) * inv1;
        float qy = expf(l - m2) * inv2;
        float qk = qy * (float)K;
        float f = logf(qk + 1e-10f) + qk / (qk + 1e-10f);
        if (gr) a = fmaf(soft, gr[k], a);
        b = fmaf(qy, f, b);
    }
    a = warp_sum(a); b = warp_sum(b);
    for (int k = lane; k < K; k += 32) {
        float l = lr[k];
        float soft = expf((er ? (l - logf(er[k])) : l) * inv_tau - m1) * inv1;
        float qy = expf(l - m2) * inv2;
        float qk = qy * (float)K;
        float f = logf(qk + 1e-10f) + qk / (qk + 1e-10f);
        float g = gr ? inv_tau * soft * (gr[k] - a) : 0.0f;
        dr[k] = g + ckl * qy * (f - b);
    }
}

}  // namespace

extern "C" int vqb_row_sqnorm(const float* a, float* out, int64_t R, int D, void* stream) {
    VQB_CHECK_ARG(a && out && R > 0 && D > 0, "row_sqnorm: bad arguments");
    sqnorm_kernel<<<(unsigned)ceil_div64(R * 32, 256), 256, 0, as_stream(stream)>>>(a, out, R, D);
    VQB_CHECK_LAUNCH("row_sqnorm");
    return VQB_OK;
}

extern "C" int vqb_vq_entropy_rows(float* dot_to_logp, const float* z, const float* codebook_sq, float temperature,
                                   int64_t* idx_out, double* sample_entropy_sum, int64_t N, int K, int D, int argmax_target,
                                   void* stream) {
    VQB_CHECK_ARG(dot_to_logp && z && codebook_sq && idx_out && sample_entropy_sum && N > 0 && K > 0 && D > 0 && temperature > 0.f,
                  "vq_entropy_rows: bad arguments");
    entropy_rows_kernel<<<(unsigned)ceil_div64(N * 32, 256), 256, 0, as_stream(stream)>>>(dot_to_logp, z, codebook_sq, 1.0f / temperature,
                                                                                           idx_out, sample_entropy_sum, N, K, D, argmax_target);
    VQB_CHECK_LAUNCH("vq_entropy_rows");
    return VQB_OK;
}

extern "C" int vqb_vq_colsum_exp(const float* logp, float* out, int64_t N, int K, void* stream) {
    VQB_CHECK_ARG(logp && out && N > 0 && K > 0, "vq_colsum_exp: bad arguments");
    int tx = K >= 128 ? 128 : 32, ty = 256 / tx;
    int rows = (int)ceil_div64(N, 148 * 4); if (rows < ty * 4) rows = ty * 4;
    dim3 block(tx, ty);
    colsum_exp_kernel<<<(unsigned)ceil_div64(N, rows), block, 0, as_stream(stream)>>>(logp, out, N, K, rows);
    VQB_CHECK_LAUNCH("vq_colsum_exp");
    return VQB_OK;
}

extern "C" int vqb_vq_entropy_finalize(const float* colsum_p, const double* sample_entropy_sum, float ratio, float* out, int64_t N,
                                       int K, void* stream) {
    VQB_CHECK_ARG(colsum_p && sample_entropy_sum && out && N > 0 && K > 0, "vq_entropy_finalize: bad arguments");
    entropy_finalize_kernel<<<1, 256, 0, as_stream(stream)>>>(colsum_p, sample_entropy_sum, ratio, out, (double)N, K);
    VQB_CHECK_LAUNCH("vq_entropy_finalize");
    return VQB_OK;
}

extern "C" int vqb_vq_entropy_bwd_rows(float* logp_to_g, const float* colsum_p, const float* g_loss, float ratio, float temperature,
                                       int64_t N, int K, const int64_t* argmax_idx, void* stream) {
    VQB_CHECK_ARG(logp_to_g && colsum_p && N > 0 && K > 0 && temperature > 0.f, "vq_entropy_bwd_rows: bad arguments");
    entropy_bwd_rows_kernel<<<(unsigned)ceil_div64(N * 32, 256), 256, 0, as_stream(stream)>>>(logp_to_g, colsum_p, g_loss, ratio,
                                                                                               1.0f / temperature, N, K, argmax_idx);
    VQB_CHECK_LAUNCH("vq_entropy_bwd_rows");
    return VQB_OK;
}

extern "C" int vqb_vq_entropy_combine_dcb(float* dcb, const float* codebook, const float* colsum_g, const float* gtz, int K, int D,
                                          void* stream) {
    VQB_CHECK_ARG(dcb && codebook && colsum_g && gtz && K > 0 && D > 0, "vq_entropy_combine_dcb: bad arguments");
    entropy_combine_dcb_kernel<<<(unsigned)ceil_div64((int64_t)K * D, 256), 256, 0, as_stream(stream)>>>(dcb, codebook, colsum_g, gtz, K, D);
    VQB_CHECK_LAUNCH("vq_entropy_combine_dcb");
    return VQB_OK;
}

extern "C" int vqb_gumbel_rows_fwd(const float* logits, const float* exp_noise, float tau, int hard, float* y, int64_t* idx_out,
                                   double* kl_sum, int64_t N, int K, void* stream) {
    VQB_CHECK_ARG(logits && y && idx_out && kl_sum && N > 0 && K > 0 && tau > 0.f, "gumbel_rows_fwd: bad arguments");
    gumbel_rows_fwd_kernel<<<(unsigned)ceil_div64(N * 32, 256), 256, 0, as_stream(stream)>>>(logits, exp_noise, 1.0f / tau, hard, y,
                                                                                              idx_out, kl_sum, N, K, nullptr);
    VQB_CHECK_LAUNCH("gumbel_rows_fwd");
    return VQB_OK;
}

extern "C" int vqb_gumbel_rows_fwd_dev(const float* logits, const float* exp_noise, const float* tau_dev, int hard, float* y,
                                       int64_t* idx_out, double* kl_sum, int64_t N, int K, void* stream) {
    VQB_CHECK_ARG(logits && y && idx_out && kl_sum && tau_dev && N > 0 && K > 0, "gumbel_rows_fwd_dev: bad arguments");
    gumbel_rows_fwd_kernel<<<(unsigned)ceil_div64(N * 32, 256), 256, 0, as_stream(stream)>>>(logits, exp_noise, 1.0f, hard, y, idx_out,
                                                                                              kl_sum, N, K, tau_dev);
    VQB_CHECK_LAUNCH("gumbel_rows_fwd_dev");
    return VQB_OK;
}

extern "C" int vqb_gumbel_rows_bwd(const float* logits, const float* exp_noise, float tau, const float* dy, const float* g_kl,
                                   float kl_scale, float* dlogits, int64_t N, int K, void* stream) {
    VQB_CHECK_ARG(logits && dlogits && N > 0 && K > 0 && tau > 0.f, "gumbel_rows_bwd: bad arguments");
    gumbel_rows_bwd_kernel<<<(unsigned)ceil_div64(N * 32, 256), 256, 0, as_stream(stream)>>>(logits, exp_noise, 1.0f / tau, dy, g_kl,
                                                                                              kl_scale, dlogits, N, K, nullptr);
    VQB_CHECK_LAUNCH("gumbel_rows_bwd");
    return VQB_OK;
}

extern "C" int vqb_gumbel_rows_bwd_dev(const float* logits, const float* exp_noise, const float* tau_dev, const float* dy,
                                       const float* g_kl, float kl_scale, float* dlogits, int64_t N, int K, void* stream) {
    VQB_CHECK_ARG(logits && dlogits && tau_dev && N > 0 && K > 0, "gumbel_rows_bwd_dev: bad arguments");
    gumbel_rows_bwd_kernel<<<(unsigned)ceil_div64(N * 32, 256), 256, 0, as_stream(stream)>>>(logits, exp_noise, 1.0f, dy, g_kl, kl_scale,
                                                                                              dlogits, N, K, tau_dev);
    VQB_CHECK_LAUNCH("gumbel_rows_bwd_dev");
    return VQB_OK;
}
