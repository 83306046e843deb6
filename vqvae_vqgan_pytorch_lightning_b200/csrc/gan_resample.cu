// Vectorised (16-byte) resampling kernels for the discriminator's down-sampling layers in the bf16 fast mode:
//   * fir4 v2: the 4x4 [1,3,3,1]^2/64 FIR forward / adjoint with 8 (bf16) or 4 (fp32) channels per thread
//   * decimate2 / zero_upsample2: y[o] = x[2o + off] and its adjoint.  A stride-2 3x3 convolution is executed as the
//     stride-1 tcgen05 convolution at full resolution followed by decimate2 (4x the MACs, but ~30x the throughput of the
//     fp32 SIMT strided kernel); its backward is zero_upsample2 followed by the stride-1 dgrad / wgrad kernels.
#include "common.cuh"

namespace {

template <typename T> struct Vec;
template <> struct Vec<float> { static constexpr int N = 4; };
template <> struct Vec<bf16> { static constexpr int N = 8; };

template <typename T>
__device__ __forceinline__ void ldvec(const T* p, float* v) {
    if constexpr (sizeof(T) == 2) {
        uint4 u = *reinterpret_cast<const uint4*>(p);
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
        for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
    } else {
        float4 a = *reinterpret_cast<const float4*>(p);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    }
}
template <typename T>
__device__ __forceinline__ void stvec(T* p, const float* v) {
    if constexpr (sizeof(T) == 2) {
        uint4 u;
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        *reinterpret_cast<uint4*>(p) = u;
    } else {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    }
}

__device__ __forceinline__ float tap(int i) { return (i == 0 || i == 3) ? 0.125f : 0.375f; }

template <typename T>
__global__ void fir4_fwd_vec_kernel(const T* __restrict__ x, T* __restrict__ y, int N, int H, int W, int C, int OH, int OW, int pad, int down) {
    constexpr int V = Vec<T>::N;
    const int Cv = C / V;
    const int64_t total = (int64_t)N * OH * OW * Cv;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int cv = (int)(i % Cv); int64_t r = i / Cv; int ow = (int)(r % OW); r /= OW; int oh = (int)(r % OH); int n = (int)(r / OH);
        float acc[V];
#pragma unroll
        for (int u = 0; u < V; ++u) acc[u] = 0.f;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int ih = oh * down - pad + a;
            if (ih < 0 || ih >= H) continue;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int iw = ow * down - pad + b;
                if (iw < 0 || iw >= W) continue;
                float v[V];
                ldvec<T>(x + (((int64_t)n * H + ih) * W + iw) * C + cv * V, v);
                const float f = tap(a) * tap(b);
#pragma unroll
                for (int u = 0; u < V; ++u) acc[u] = fmaf(f, v[u], acc[u]);
            }
        }
        stvec<T>(y + i * V, acc);
    }
}

template <typename T>
__global__ void fir4_bwd_vec_kernel(const T* __restrict__ dy, T* __restrict__ dx, int N, int H, int W, int C, int OH, int OW, int pad, int down) {
    constexpr int V = Vec<T>::N;
    const int Cv = C / V;
    const int64_t total = (int64_t)N * H * W * Cv;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int cv = (int)(i % Cv); int64_t r = i / Cv; int w = (int)(r % W); r /= W; int h = (int)(r % H); int n = (int)(r / H);
        float acc[V];
#pragma unroll
        for (int u = 0; u < V; ++u) acc[u] = 0.f;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int th = h + pad - a;
            if (th < 0 || (down == 2 && (th & 1))) continue;
            const int oh = (down == 2) ? (th >> 1) : th;
            if (oh >= OH) continue;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int tw = w + pad - b;
                if (tw < 0 || (down == 2 && (tw & 1))) continue;
                const int ow = (down == 2) ? (tw >> 1) : tw;
                if (ow >= OW) continue;
                float v[V];
                ldvec<T>(dy + (((int64_t)n * OH + oh) * OW + ow) * C + cv * V, v);
                const float f = tap(a) * tap(b);
#pragma unroll
                for (int u = 0; u < V; ++u) acc[u] = fmaf(f, v[u], acc[u]);
            }
        }
        stvec<T>(dx + i * V, acc);
    }
}

// y[n,oh,ow,:] = x[n, 2*oh+off, 2*ow+off, :]
template <typename T>
__global__ void decimate2_kernel(const T* __restrict__ x, T* __restrict__ y, int N, int H, int W, int C, int OH, int OW, int off) {
    constexpr int V = Vec<T>::N;
    const int Cv = C / V;
    const int64_t total = (int64_t)N * OH * OW * Cv;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int cv = (int)(i % Cv); int64_t r = i / Cv; int ow = (int)(r % OW); r /= OW; int oh = (int)(r % OH); int n = (int)(r / OH);
        float v[V];
        ldvec<T>(x + (((int64_t)n * H + 2 * oh + off) * W + 2 * ow + off) * C + cv * V, v);
        stvec<T>(y + i * V, v);
    }
}

// adjoint: x[n,h,w,:] = y[n,(h-off)/2,(w-off)/2,:] where both are exact and in range, else 0 (writes all of x)
template <typename T>
__global__ void zero_upsample2_kernel(const T* __restrict__ y, T* __restrict__ x, int N, int H, int W, int C, int OH, int OW, int off) {
    constexpr int V = Vec<T>::N;
    const int Cv = C / V;
    const int64_t total = (int64_t)N * H * W * Cv;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int cv = (int)(i % Cv); int64_t r = i / Cv; int w = (int)(r % W); r /= W; int h = (int)(r % H); int n = (int)(r / H);
        float v[V];
#pragma unroll
        for (int u = 0; u < V; ++u) v[u] = 0.f;
        const int th = h - off, tw = w - off;
        if (th >= 0 && tw >= 0 && !(th & 1) && !(tw & 1) && (th >> 1) < OH && (tw >> 1) < OW)
            ldvec<T>(y + (((int64_t)n * OH + (th >> 1)) * OW + (tw >> 1)) * C + cv * V, v);
        stvec<T>(x + i * V, v);
    }
}

inline int ew_grid(int64_t n) { int64_t b = (n + 255) / 256; if (b > 148 * 32) b = 148 * 32; if (b < 1) b = 1; return (int)b; }

}  // namespace

// vectorised entry points used by vqb_fir4_fwd / vqb_fir4_bwd when C is a multiple of the vector width
int vqb_fir4_fwd_vec(const void* x, void* y, int dtype, int N, int H, int W, int C, int OH, int OW, int pad, int down, cudaStream_t st) {
    if (dtype == VQB_BF16) fir4_fwd_vec_kernel<bf16><<<ew_grid((int64_t)N * OH * OW * C / 8), 256, 0, st>>>((const bf16*)x, (bf16*)y, N, H, W, C, OH, OW, pad, down);
    else fir4_fwd_vec_kernel<float><<<ew_grid((int64_t)N * OH * OW * C / 4), 256, 0, st>>>((const float*)x, (float*)y, N, H, W, C, OH, OW, pad, down);
    VQB_CHECK_LAUNCH("fir4_fwd_vec");
    return VQB_OK;
}
int vqb_fir4_bwd_vec(const void* dy, void* dx, int dtype, int N, int H, int W, int C, int OH, int OW, int pad, int down, cudaStream_t st) {
    if (dtype == VQB_BF16) fir4_bwd_vec_kernel<bf16><<<ew_grid((int64_t)N * H * W * C / 8), 256, 0, st>>>((const bf16*)dy, (bf16*)dx, N, H, W, C, OH, OW, pad, down);
    else fir4_bwd_vec_kernel<float><<<ew_grid((int64_t)N * H * W * C / 4), 256, 0, st>>>((const float*)dy, (float*)dx, N, H, W, C, OH, OW, pad, down);
    VQB_CHECK_LAUNCH("fir4_bwd_vec");
    return VQB_OK;
}

extern "C" int vqb_decimate2(const void* x, void* y, int dtype, int N, int H, int W, int C, int OH, int OW, int off, void* stream) {
    VQB_CHECK_ARG(x && y && N > 0 && H > 0 && W > 0 && C > 0 && OH > 0 && OW > 0 && off >= 0, "decimate2: bad arguments");
    VQB_CHECK_ARG(2 * (OH - 1) + off < H && 2 * (OW - 1) + off < W, "decimate2: output grid exceeds the input");
    VQB_CHECK_ARG(C % ((dtype == VQB_BF16) ? 8 : 4) == 0, "decimate2: C must be a multiple of the 16-byte vector width");
    if (dtype == VQB_BF16) decimate2_kernel<bf16><<<ew_grid((int64_t)N * OH * OW * C / 8), 256, 0, as_stream(stream)>>>((const bf16*)x, (bf16*)y, N, H, W, C, OH, OW, off);
    else decimate2_kernel<float><<<ew_grid((int64_t)N * OH * OW * C / 4), 256, 0, as_stream(stream)>>>((const float*)x, (float*)y, N, H, W, C, OH, OW, off);
    VQB_CHECK_LAUNCH("decimate2");
    return VQB_OK;
}

extern "C" int vqb_zero_upsample2(const void* y, void* x, int dtype, int N, int H, int W, int C, int OH, int OW, int off, void* stream) {
    VQB_CHECK_ARG(x && y && N > 0 && H > 0 && W > 0 && C > 0 && OH > 0 && OW > 0 && off >= 0, "zero_upsample2: bad arguments");
    VQB_CHECK_ARG(C % ((dtype == VQB_BF16) ? 8 : 4) == 0, "zero_upsample2: C must be a multiple of the 16-byte vector width");
    if (dtype == VQB_BF16) zero_upsample2_kernel<bf16><<<ew_grid((int64_t)N * H * W * C / 8), 256, 0, as_stream(stream)>>>((const bf16*)y, (bf16*)x, N, H, W, C, OH, OW, off);
    else zero_upsample2_kernel<float><<<ew_grid((int64_t)N * H * W * C / 4), 256, 0, as_stream(stream)>>>((const float*)y, (float*)x, N, H, W, C, OH, OW, off);
    VQB_CHECK_LAUNCH("zero_upsample2");
    return VQB_OK;
}
