// HBM-bound operators of the VQGAN loss heads (NHWC):
//   * 4x4 FIR low-pass [1,3,3,1] x [1,3,3,1] / 64 with padding and decimation -- the only upfirdn2d configurations the
//     StyleGAN2 discriminator uses (reference: .../stylegan2_discriminator/utils/ops/upfirdn2d.{py,cu}, specialisations
//     <1,1,1,1,4,4,..> and <1,1,2,2,4,4,..>, upfirdn2d.cu:217,298; call sites conv2d_resample.py:107-122), forward and
//     backward (the adjoint: scatter of the same taps)
//   * 2x2/stride-2 max-pool forward/backward (torchvision VGG16 features, lpips_pytorch/modules/networks.py:89-97)
//   * per-channel affine (z-score of BaseNet.z_score, networks.py:48-49)
//   * LPIPS tap: channel-unit-normalise both feature maps, squared difference, 1x1 "lin" weights, spatial+batch mean
//     (lpips.py:31-38, utils.py:6-8) fused into one pass, and its backward w.r.t. the second feature map
//   * minibatch standard deviation (discriminator.py:277-293) forward/backward
#include "common.cuh"

namespace {

__device__ __forceinline__ float fir_tap(int i) { return (i == 0 || i == 3) ? 0.125f : 0.375f; }   // [1,3,3,1]/8 per axis

template <typename T>
__global__ void fir4_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int N, int H, int W, int C, int OH, int OW, int pad,
                                int down) {
    int64_t total = (int64_t)N * OH * OW * C;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int c = (int)(i % C); int64_t r = i / C; int ow = (int)(r % OW); r /= OW; int oh = (int)(r % OH); int n = (int)(r / OH);
        float acc = 0.f;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            int ih = oh * down - pad + a;
            if (ih < 0 || ih >= H) continue;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                int iw = ow * down - pad + b;
                if (iw < 0 || iw >= W) continue;
                acc = fmaf(fir_tap(a) * fir_tap(b), ld1(x + (((int64_t)n * H + ih) * W + iw) * C + c), acc);
            }
        }
        st1(y + i, acc);
    }
}

// dx[h,w] = sum_{a,b} f[a] f[b] dy[(h+pad-a)/down, (w+pad-b)/down] over taps where the division is exact and in range
template <typename T>
__global__ void fir4_bwd_kernel(const T* __restrict__ dy, T* __restrict__ dx, int N, int H, int W, int C, int OH, int OW, int pad,
                                int down) {
    int64_t total = (int64_t)N * H * W * C;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int c = (int)(i % C); int64_t r = i / C; int w = (int)(r % W); r /= W; int h = (int)(r % H); int n = (int)(r / H);
        float acc = 0.f;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            int th = h + pad - a;
            if (th < 0 || th % down) continue;
            int oh = th / down;
            if (oh >= OH) continue;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                int tw = w + pad - b;
                if (tw < 0 || tw % down) continue;
                int ow = tw / down;
                if (ow >= OW) continue;
                acc = fmaf(fir_tap(a) * fir_tap(b), ld1(dy + (((int64_t)n * OH + oh) * OW + ow) * C + c), acc);
            }
        }
        st1(dx + i, acc);
    }
}

template <typename T>
__global__ void maxpool2_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int N, int H, int W, int C) {
    // y [N,H,W,C], x [N,2H,2W,C]
    int64_t total = (int64_t)N * H * W * C;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int c = (int)(i % C); int64_t r = i / C; int w = (int)(r % W); r /= W; int h = (int)(r % H); int n = (int)(r / H);
        const T* p = x + (((int64_t)n * 2 * H + 2 * h) * 2 * W + 2 * w) * C + c;
        int64_t rowx = (int64_t)2 * W * C;
        float m = fmaxf(fmaxf(ld1(p), ld1(p + C)), fmaxf(ld1(p + rowx), ld1(p + rowx + C)));
        st1(y + i, m);
    }
}

// gradient goes to the FIRST maximal element in window scan order (torch max_pool2d semantics)
template <typename T, typename TG>
__global__ void maxpool2_bwd_kernel(const T* __restrict__ x, const TG* __restrict__ dy, TG* __restrict__ dx, int N, int H, int W, int C) {
    int64_t total = (int64_t)N * H * W * C;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int c = (int)(i % C); int64_t r = i / C; int w = (int)(r % W); r /= W; int h = (int)(r % H); int n = (int)(r / H);
        int64_t base = (((int64_t)n * 2 * H + 2 * h) * 2 * W + 2 * w) * C + c;
        int64_t rowx = (int64_t)2 * W * C;
        float v0 = ld1(x + base), v1 = ld1(x + base + C), v2 = ld1(x + base + rowx), v3 = ld1(x + base + rowx + C);
        int arg = 0; float m = v0;
        if (v1 > m) { m = v1; arg = 1; }
        if (v2 > m) { m = v2; arg = 2; }
        if (v3 > m) { m = v3; arg = 3; }
        float g = ld1(dy + i);
        st1(dx + base, arg == 0 ? g : 0.f);
        st1(dx + base + C, arg == 1 ? g : 0.f);
        st1(dx + base + rowx, arg == 2 ? g : 0.f);
        st1(dx + base + rowx + C, arg == 3 ? g : 0.f);
    }
}

// 3x3 / stride-2 max-pool without padding (torchvision AlexNet features): y [N,OH,OW,C], OH = (H-3)/2+1
template <typename T>
__global__ void maxpool3s2_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int N, int H, int W, int C, int OH, int OW) {
    int64_t total = (int64_t)N * OH * OW * C;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int c = (int)(i % C); int64_t r = i / C; int ow = (int)(r % OW); r /= OW; int oh = (int)(r % OH); int n = (int)(r / OH);
        const T* p = x + (((int64_t)n * H + 2 * oh) * W + 2 * ow) * C + c;
        float m = -INFINITY;
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b) m = fmaxf(m, ld1(p + ((int64_t)a * W + b) * C));
        st1(y + i, m);
    }
}

// overlapping windows: gather form.  Each input element belongs to <= 4 windows; it receives a window's gradient iff it is
// the FIRST maximal element of that window in scan order (torch max_pool2d argmax semantics).
template <typename T, typename TG>
__global__ void maxpool3s2_bwd_kernel(const T* __restrict__ x, const TG* __restrict__ dy, TG* __restrict__ dx, int N, int H, int W,
                                      int C, int OH, int OW) {
    int64_t total = (int64_t)N * H * W * C;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int c = (int)(i % C); int64_t r = i / C; int w = (int)(r % W); r /= W; int h = (int)(r % H); int n = (int)(r / H);
        const float xv = ld1(x + i);
        float acc = 0.f;
        // windows covering row h: 2*oh <= h <= 2*oh + 2  ->  oh in [ceil((h-2)/2), floor(h/2)], clipped to the output grid
        const int oh_lo = h > 0 ? (h - 1) / 2 : 0, oh_hi = min(h / 2, OH - 1);
        const int ow_lo = w > 0 ? (w - 1) / 2 : 0, ow_hi = min(w / 2, OW - 1);
        for (int oh = oh_lo; oh <= oh_hi; ++oh) {
            for (int ow = ow_lo; ow <= ow_hi; ++ow) {
                const T* p = x + (((int64_t)n * H + 2 * oh) * W + 2 * ow) * C + c;
                const int mine = (h - 2 * oh) * 3 + (w - 2 * ow);
                float m = -INFINITY; int arg = 0;
#pragma unroll
                for (int a = 0; a < 3; ++a)
#pragma unroll
                    for (int b = 0; b < 3; ++b) {
                        float v = ld1(p + ((int64_t)a * W + b) * C);
                        if (v > m) { m = v; arg = a * 3 + b; }
                    }
                if (arg == mine && m == xv) acc += ld1(dy + (((int64_t)n * OH + oh) * OW + ow) * C + c);
            }
        }
        st1(dx + i, acc);
    }
}

template <typename TI, typename TO>
__global__ void channel_affine_kernel(const TI* __restrict__ x, TO* __restrict__ y, const float* __restrict__ scale,
                                      const float* __restrict__ shift, int64_t P, int C) {
    int64_t total = P * C;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int c = (int)(i % C);
        st1(y + i, ld1(x + i) * scale[c] + shift[c]);
    }
}

// one warp per pixel: L += sum_c w_c (fx_c/(|fx|+eps) - fy_c/(|fy|+eps))^2
template <typename T>
__global__ void lpips_tap_fwd_kernel(const T* __restrict__ fx, const T* __restrict__ fy, const float* __restrict__ w,
                                     double* __restrict__ out, int64_t P, int C) {
    int64_t pix = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    float acc = 0.f;
    if (pix < P) {
        const T* a = fx + pix * C; const T* b = fy + pix * C;
        float na = 0.f, nb = 0.f;
        for (int c = lane; c < C; c += 32) { float u = ld1(a + c), v = ld1(b + c); na = fmaf(u, u, na); nb = fmaf(v, v, nb); }
        na = warp_sum(na); nb = warp_sum(nb);
        const float ia = 1.0f / (sqrtf(na) + 1e-10f), ib = 1.0f / (sqrtf(nb) + 1e-10f);
        for (int c = lane; c < C; c += 32) { float d = ld1(a + c) * ia - ld1(b + c) * ib; acc = fmaf(w[c] * d, d, acc); }
        acc = warp_sum(acc);
    }
    __shared__ float sh[8];
    if (lane == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += sh[i];
        atomicAdd(out, t);
    }
}

// d/dfy of scale * sum_c w_c (u_c - v_c)^2, v = fy/(|fy|+eps):  g_c = -2 w_c (u_c - v_c);
// dfy_j = s * ( g_j/(n+eps) - (sum_c g_c fy_c) fy_j / (n (n+eps)^2) )
template <typename T, typename TG>
__global__ void lpips_tap_bwd_kernel(const T* __restrict__ fx, const T* __restrict__ fy, const float* __restrict__ w,
                                     const float* __restrict__ upstream, float scale, TG* __restrict__ dfy, int64_t P, int C) {
    int64_t pix = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (pix >= P) return;
    const float s = scale * (upstream ? upstream[0] : 1.0f);
    const T* a = fx + pix * C; const T* b = fy + pix * C;
    float na = 0.f, nb = 0.f;
    for (int c = lane; c < C; c += 32) { float u = ld1(a + c), v = ld1(b + c); na = fmaf(u, u, na); nb = fmaf(v, v, nb); }
    na = warp_sum(na); nb = warp_sum(nb);
    const float nrm = sqrtf(nb);
    const float ia = 1.0f / (sqrtf(na) + 1e-10f), ib = 1.0f / (nrm + 1e-10f);
    float dot = 0.f;
    for (int c = lane; c < C; c += 32) {
        float v = ld1(b + c);
        float g = -2.0f * w[c] * (ld1(a + c) * ia - v * ib);
        dot = fmaf(g, v, dot);
    }
    dot = warp_sum(dot);
    const float k = (nrm > 0.f) ? dot * ib * ib / nrm : 0.f;
    for (int c = lane; c < C; c += 32) {
        float v = ld1(b + c);
        float g = -2.0f * w[c] * (ld1(a + c) * ia - v * ib);
        st1(dfy + pix * C + c, s * (g * ib - k * v));
    }
}

// minibatch stddev: x [N][HW][C] (NHWC), groups of G consecutive... the reference reshapes N -> (G, n): sample s belongs to
// group (s % n), n = N / G.  y[s][p][C] = stat[s % n]; out has C+1 channels.
template <typename T>
__global__ void mbstd_stat_kernel(const T* __restrict__ x, float* __restrict__ stat, int N, int G, int64_t E) {
    // one block per group index j in [0, n): stat[j] = mean_e sqrt(var_g(x[g*n + j][e]) + 1e-8)
    const int n = N / G, j = blockIdx.x;
    double acc = 0.0;
    for (int64_t e = threadIdx.x; e < E; e += blockDim.x) {
        float mean = 0.f;
        for (int g = 0; g < G; ++g) mean += ld1(x + ((int64_t)(g * n + j)) * E + e);
        mean /= (float)G;
        float var = 0.f;
        for (int g = 0; g < G; ++g) { float d = ld1(x + ((int64_t)(g * n + j)) * E + e) - mean; var = fmaf(d, d, var); }
        var /= (float)G;
        acc += (double)sqrtf(var + 1e-8f);
    }
    acc = warp_sum(acc);
    __shared__ double sh[8];
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += sh[i];
        stat[j] = (float)(t / (double)E);
    }
}

template <typename T>
__global__ void mbstd_concat_kernel(const T* __restrict__ x, const float* __restrict__ stat, T* __restrict__ y, int N, int G, int HW, int C) {
    const int n = N / G;
    int64_t total = (int64_t)N * HW * (C + 1);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int c = (int)(i % (C + 1)); int64_t r = i / (C + 1); int s = (int)(r / HW);
        st1(y + i, c < C ? ld1(x + r * C + c) : stat[s % n]);
    }
}

// dx[s][e] = dy_main[s][e] + gstat[j] * (1/E) * (x - mean_g) / (G * sqrt(var+1e-8)),  gstat[j] = sum over the extra channel of group j
template <typename T, typename TG>
__global__ void mbstd_bwd_kernel(const T* __restrict__ x, const TG* __restrict__ dy, TG* __restrict__ dx, int N, int G, int HW, int C) {
    const int n = N / G, j = blockIdx.x;
    const int64_t E = (int64_t)HW * C;
    __shared__ float gs;
    if (threadIdx.x < 32) {
        float a = 0.f;
        for (int t = threadIdx.x; t < G * HW; t += 32) {
            int g = t / HW, p = t - g * HW;
            a += ld1(dy + (((int64_t)(g * n + j)) * HW + p) * (C + 1) + C);
        }
        a = warp_sum(a);
        if (threadIdx.x == 0) gs = a;
    }
    __syncthreads();
    const float gstat = gs / (float)E;
    for (int64_t e = threadIdx.x; e < E; e += blockDim.x) {
        int64_t p = e / C; int c = (int)(e - p * C);
        float mean = 0.f;
        for (int g = 0; g < G; ++g) mean += ld1(x + ((int64_t)(g * n + j)) * E + e);
        mean /= (float)G;
        float var = 0.f;
        for (int g = 0; g < G; ++g) { float d = ld1(x + ((int64_t)(g * n + j)) * E + e) - mean; var = fmaf(d, d, var); }
        var /= (float)G;
        const float inv = gstat / ((float)G * sqrtf(var + 1e-8f));
        for (int g = 0; g < G; ++g) {
            int64_t s = (int64_t)(g * n + j);
            float d = ld1(x + s * E + e) - mean;
            float main = ld1(dy + (s * HW + p) * (C + 1) + c);
            st1(dx + s * E + e, main + inv * d);
        }
    }
}

inline int ew_grid(int64_t n) { int64_t b = (n + 255) / 256; if (b > 148 * 16) b = 148 * 16; if (b < 1) b = 1; return (int)b; }

}  // namespace

int vqb_fir4_fwd_vec(const void* x, void* y, int dtype, int N, int H, int W, int C, int OH, int OW, int pad, int down, cudaStream_t st);
int vqb_fir4_bwd_vec(const void* dy, void* dx, int dtype, int N, int H, int W, int C, int OH, int OW, int pad, int down, cudaStream_t st);

extern "C" int vqb_fir4_fwd(const void* x, void* y, int dtype, int N, int H, int W, int C, int pad, int down, void* stream) {
    VQB_CHECK_ARG(x && y && N > 0 && H > 0 && W > 0 && C > 0 && pad >= 0 && down >= 1, "fir4_fwd: bad arguments");
    int OH = (H + 2 * pad - 4) / down + 1, OW = (W + 2 * pad - 4) / down + 1;
    VQB_CHECK_ARG(OH > 0 && OW > 0, "fir4_fwd: empty output");
    if (C % ((dtype == VQB_BF16) ? 8 : 4) == 0 && (dtype == VQB_BF16 || dtype == VQB_F32) && down <= 2)
        return vqb_fir4_fwd_vec(x, y, dtype, N, H, W, C, OH, OW, pad, down, as_stream(stream));
    VQB_DISPATCH_1(dtype, T, (fir4_fwd_kernel<T><<<ew_grid((int64_t)N * OH * OW * C), 256, 0, as_stream(stream)>>>(
                                 (const T*)x, (T*)y, N, H, W, C, OH, OW, pad, down));)
    VQB_CHECK_LAUNCH("fir4_fwd");
    return VQB_OK;
}

extern "C" int vqb_fir4_bwd(const void* dy, void* dx, int dtype, int N, int H, int W, int C, int pad, int down, void* stream) {
    VQB_CHECK_ARG(dy && dx && N > 0 && H > 0 && W > 0 && C > 0 && pad >= 0 && down >= 1, "fir4_bwd: bad arguments");
    int OH = (H + 2 * pad - 4) / down + 1, OW = (W + 2 * pad - 4) / down + 1;
    if (C % ((dtype == VQB_BF16) ? 8 : 4) == 0 && (dtype == VQB_BF16 || dtype == VQB_F32) && down <= 2)
        return vqb_fir4_bwd_vec(dy, dx, dtype, N, H, W, C, OH, OW, pad, down, as_stream(stream));
    VQB_DISPATCH_1(dtype, T, (fir4_bwd_kernel<T><<<ew_grid((int64_t)N * H * W * C), 256, 0, as_stream(stream)>>>(
                                 (const T*)dy, (T*)dx, N, H, W, C, OH, OW, pad, down));)
    VQB_CHECK_LAUNCH("fir4_bwd");
    return VQB_OK;
}

extern "C" int vqb_maxpool2_fwd(const void* x, void* y, int dtype, int N, int H, int W, int C, void* stream) {
    VQB_CHECK_ARG(x && y && N > 0 && H > 0 && W > 0 && C > 0, "maxpool2_fwd: bad arguments");
    VQB_DISPATCH_1(dtype, T, (maxpool2_fwd_kernel<T><<<ew_grid((int64_t)N * H * W * C), 256, 0, as_stream(stream)>>>((const T*)x, (T*)y, N, H, W, C));)
    VQB_CHECK_LAUNCH("maxpool2_fwd");
    return VQB_OK;
}

extern "C" int vqb_maxpool2_bwd(const void* x, int x_dtype, const void* dy, void* dx, int g_dtype, int N, int H, int W, int C, void* stream) {
    VQB_CHECK_ARG(x && dy && dx && N > 0 && H > 0 && W > 0 && C > 0, "maxpool2_bwd: bad arguments");
    VQB_DISPATCH_1(x_dtype, T, VQB_DISPATCH_1(g_dtype, TG, (maxpool2_bwd_kernel<T, TG><<<ew_grid((int64_t)N * H * W * C), 256, 0, as_stream(stream)>>>(
                                                               (const T*)x, (const TG*)dy, (TG*)dx, N, H, W, C));))
    VQB_CHECK_LAUNCH("maxpool2_bwd");
    return VQB_OK;
}

extern "C" int vqb_maxpool3s2_fwd(const void* x, void* y, int dtype, int N, int H, int W, int C, void* stream) {
    VQB_CHECK_ARG(x && y && N > 0 && H >= 3 && W >= 3 && C > 0, "maxpool3s2_fwd: bad arguments (needs H, W >= 3)");
    const int OH = (H - 3) / 2 + 1, OW = (W - 3) / 2 + 1;
    VQB_DISPATCH_1(dtype, T, (maxpool3s2_fwd_kernel<T><<<ew_grid((int64_t)N * OH * OW * C), 256, 0, as_stream(stream)>>>((const T*)x, (T*)y, N, H, W, C, OH, OW));)
    VQB_CHECK_LAUNCH("maxpool3s2_fwd");
    return VQB_OK;
}

extern "C" int vqb_maxpool3s2_bwd(const void* x, int x_dtype, const void* dy, void* dx, int g_dtype, int N, int H, int W, int C, void* stream) {
    VQB_CHECK_ARG(x && dy && dx && N > 0 && H >= 3 && W >= 3 && C > 0, "maxpool3s2_bwd: bad arguments (needs H, W >= 3)");
    const int OH = (H - 3) / 2 + 1, OW = (W - 3) / 2 + 1;
    VQB_DISPATCH_1(x_dtype, T, VQB_DISPATCH_1(g_dtype, TG, (maxpool3s2_bwd_kernel<T, TG><<<ew_grid((int64_t)N * H * W * C), 256, 0, as_stream(stream)>>>(
        (const T*)x, (const TG*)dy, (TG*)dx, N, H, W, C, OH, OW));))
    VQB_CHECK_LAUNCH("maxpool3s2_bwd");
    return VQB_OK;
}

extern "C" int vqb_channel_affine(const void* x, int x_dtype, void* y, int y_dtype, const float* scale, const float* shift, int64_t P,
                                  int C, void* stream) {
    VQB_CHECK_ARG(x && y && scale && shift && P > 0 && C > 0, "channel_affine: bad arguments");
    VQB_DISPATCH_1(x_dtype, TI, VQB_DISPATCH_1(y_dtype, TO, (channel_affine_kernel<TI, TO><<<ew_grid(P * C), 256, 0, as_stream(stream)>>>(
                                                                (const TI*)x, (TO*)y, scale, shift, P, C));))
    VQB_CHECK_LAUNCH("channel_affine");
    return VQB_OK;
}

extern "C" int vqb_lpips_tap_fwd(const void* fx, const void* fy, int dtype, const float* w, double* out, int64_t P, int C, void* stream) {
    VQB_CHECK_ARG(fx && fy && w && out && P > 0 && C > 0, "lpips_tap_fwd: bad arguments");
    VQB_DISPATCH_1(dtype, T, (lpips_tap_fwd_kernel<T><<<(unsigned)ceil_div64(P * 32, 256), 256, 0, as_stream(stream)>>>(
                                 (const T*)fx, (const T*)fy, w, out, P, C));)
    VQB_CHECK_LAUNCH("lpips_tap_fwd");
    return VQB_OK;
}

extern "C" int vqb_lpips_tap_bwd(const void* fx, const void* fy, int dtype, const float* w, const float* upstream, float scale,
                                 void* dfy, int g_dtype, int64_t P, int C, void* stream) {
    VQB_CHECK_ARG(fx && fy && w && dfy && P > 0 && C > 0, "lpips_tap_bwd: bad arguments");
    VQB_DISPATCH_1(dtype, T, VQB_DISPATCH_1(g_dtype, TG, (lpips_tap_bwd_kernel<T, TG><<<(unsigned)ceil_div64(P * 32, 256), 256, 0, as_stream(stream)>>>(
                                                             (const T*)fx, (const T*)fy, w, upstream, scale, (TG*)dfy, P, C));))
    VQB_CHECK_LAUNCH("lpips_tap_bwd");
    return VQB_OK;
}

extern "C" int vqb_mbstd_fwd(const void* x, void* y, float* stat, int dtype, int N, int G, int HW, int C, void* stream) {
    VQB_CHECK_ARG(x && y && stat && N > 0 && G > 0 && N % G == 0 && HW > 0 && C > 0, "mbstd_fwd: bad arguments (batch must be a multiple of the group size)");
    VQB_DISPATCH_1(dtype, T, (mbstd_stat_kernel<T><<<N / G, 256, 0, as_stream(stream)>>>((const T*)x, stat, N, G, (int64_t)HW * C));
                   (mbstd_concat_kernel<T><<<ew_grid((int64_t)N * HW * (C + 1)), 256, 0, as_stream(stream)>>>((const T*)x, stat, (T*)y, N, G, HW, C));)
    VQB_CHECK_LAUNCH("mbstd_fwd");
    return VQB_OK;
}

extern "C" int vqb_mbstd_bwd(const void* x, int x_dtype, const void* dy, void* dx, int g_dtype, int N, int G, int HW, int C, void* stream) {
    VQB_CHECK_ARG(x && dy && dx && N > 0 && G > 0 && N % G == 0 && HW > 0 && C > 0, "mbstd_bwd: bad arguments");
    VQB_DISPATCH_1(x_dtype, T, VQB_DISPATCH_1(g_dtype, TG, (mbstd_bwd_kernel<T, TG><<<N / G, 256, 0, as_stream(stream)>>>(
                                                               (const T*)x, (const TG*)dy, (TG*)dx, N, G, HW, C));))
    VQB_CHECK_LAUNCH("mbstd_bwd");
    return VQB_OK;
}
