// HBM-bound plumbing kernels: layout changes, weight packing, 2x resampling, loss reductions, activation
// backward, column sums, fused AdamW.  All are grid-stride, vectorised where alignment allows, and sized as
// multiples of the SM count (148 on B200).
#include "common.cuh"
#include <stdarg.h>
#include <string.h>

static thread_local char g_err[512] = "";
void vqb_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
extern "C" const char* vqb_last_error(void) { return g_err; }
extern "C" const char* vqb_version(void) { return "vqgan_b200 0.1 sm_100a"; }
extern "C" int vqb_device_supports_tcgen05(void) {
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
    return major == 10;
}

static const int kSMs = 148;
static inline int grid_for(int64_t work_items, int threads, int max_waves = 16) {
    int64_t b = ceil_div64(work_items, threads);
    int64_t cap = (int64_t)kSMs * max_waves;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

// ---------------------------------------------------------------------------------------------------
// NCHW <-> NHWC
// ---------------------------------------------------------------------------------------------------
template <typename TO>
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, TO* __restrict__ y, int64_t N, int64_t C, int64_t HW,
                                    int do_clamp, float lo, float hi, float shift, float scale) {
    // one thread per (n, p): reads C strided planes (coalesced across p), writes C contiguous values
    int64_t total = N * HW;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t n = i / HW, p = i - n * HW;
        const float* src = x + n * C * HW + p;
        TO* dst = y + i * C;
        for (int64_t c = 0; c < C; ++c) {
            float v = src[c * HW];
            if (do_clamp) v = fminf(fmaxf(v, lo), hi);
            st1(dst + c, (v - shift) * scale);
        }
    }
}

template <typename TI>
__global__ void nhwc_to_nchw_kernel(const TI* __restrict__ x, float* __restrict__ y, int64_t N, int64_t C, int64_t HW,
                                    float scale, float shift, int do_clamp, float lo, float hi) {
    int64_t total = N * HW;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t n = i / HW, p = i - n * HW;
        const TI* src = x + i * C;
        float* dst = y + n * C * HW + p;
        for (int64_t c = 0; c < C; ++c) {
            float v = ld1(src + c) * scale + shift;
            if (do_clamp) v = fminf(fmaxf(v, lo), hi);
            dst[c * HW] = v;
        }
    }
}

// tiled transpose for wide channel counts: [N][C][HW] <-> [N][HW][C]
template <typename TI, typename TO>
__global__ void transpose_tiled_kernel(const TI* __restrict__ x, TO* __restrict__ y, int R, int Cc, float scale,
                                       float shift) {
    // x: [batch][R][Cc] -> y: [batch][Cc][R]
    __shared__ float tile[32][33];
    int b = blockIdx.z;
    const TI* xb = x + (int64_t)b * R * Cc;
    TO* yb = y + (int64_t)b * R * Cc;
    int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        int r = r0 + j, c = c0 + threadIdx.x;
        if (r < R && c < Cc) tile[j][threadIdx.x] = ld1(xb + (int64_t)r * Cc + c);
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        int c = c0 + j, r = r0 + threadIdx.x;
        if (r < R && c < Cc) st1(yb + (int64_t)c * R + r, tile[threadIdx.x][j] * scale + shift);
    }
}

extern "C" int vqb_nchw_to_nhwc(const float* x, void* y, int out_dtype, int64_t N, int64_t C, int64_t H, int64_t W,
                                int do_clamp, float lo, float hi, float shift, float scale, void* stream) {
    VQB_CHECK_ARG(x && y && N > 0 && C > 0 && H > 0 && W > 0, "nchw_to_nhwc: bad arguments");
    int64_t HW = H * W;
    if (C <= 8 || do_clamp) {
        int g = grid_for(N * HW, 256);
        VQB_DISPATCH_1(out_dtype, TO, (nchw_to_nhwc_kernel<TO><<<g, 256, 0, as_stream(stream)>>>(
                                          x, (TO*)y, N, C, HW, do_clamp, lo, hi, shift, scale));)
    } else {
        dim3 grid((unsigned)ceil_div64(HW, 32), (unsigned)ceil_div64(C, 32), (unsigned)N), block(32, 8);
        // (v - shift) * scale == v*scale + (-shift*scale)
        VQB_DISPATCH_1(out_dtype, TO, (transpose_tiled_kernel<float, TO><<<grid, block, 0, as_stream(stream)>>>(
                                          x, (TO*)y, (int)C, (int)HW, scale, -shift * scale));)
    }
    VQB_CHECK_LAUNCH("nchw_to_nhwc");
    return VQB_OK;
}

extern "C" int vqb_nhwc_to_nchw(const void* x, int in_dtype, float* y, int64_t N, int64_t C, int64_t H, int64_t W,
                                float scale, float shift, int do_clamp, float lo, float hi, void* stream) {
    VQB_CHECK_ARG(x && y && N > 0 && C > 0 && H > 0 && W > 0, "nhwc_to_nchw: bad arguments");
    int64_t HW = H * W;
    if (C <= 8 || do_clamp) {
        int g = grid_for(N * HW, 256);
        VQB_DISPATCH_1(in_dtype, TI, (nhwc_to_nchw_kernel<TI><<<g, 256, 0, as_stream(stream)>>>(
                                         (const TI*)x, y, N, C, HW, scale, shift, do_clamp, lo, hi));)
    } else {
        dim3 grid((unsigned)ceil_div64(C, 32), (unsigned)ceil_div64(HW, 32), (unsigned)N), block(32, 8);
        VQB_DISPATCH_1(in_dtype, TI, (transpose_tiled_kernel<TI, float><<<grid, block, 0, as_stream(stream)>>>(
                                         (const TI*)x, y, (int)HW, (int)C, scale, shift));)
    }
    VQB_CHECK_LAUNCH("nhwc_to_nchw");
    return VQB_OK;
}

template <typename TI, typename TO>
__global__ void convert_kernel(const TI* __restrict__ x, TO* __restrict__ y, int64_t n) {
    int64_t n4 = n >> 2;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x)
        st4(y + i * 4, ld4(x + i * 4));
    for (int64_t i = (n4 << 2) + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        st1(y + i, ld1(x + i));
}

// fp32 [rows][C] -> bf16 [rows][2C] = [hi | lo] with x ~ hi + lo (16 mantissa bits): the split-precision operand layout of the
// strict numeric mode's tensor-core convolutions (conv_tc.cu: the k loop wraps over the two halves)
__global__ void split_hi_lo_kernel(const float* __restrict__ x, bf16* __restrict__ y, int64_t rows, int C) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;          // one float4 per thread
    const int c4 = C >> 2;
    if (i >= rows * c4) return;
    const int64_t r = i / c4;
    const int c = (int)(i - r * c4) * 4;
    const float4 v = *reinterpret_cast<const float4*>(x + r * C + c);
    const float f[4] = {v.x, v.y, v.z, v.w};
    uint32_t hw[2], lw[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const bf16 h0 = __float2bfloat16_rn(f[2 * k]), h1 = __float2bfloat16_rn(f[2 * k + 1]);
        const bf16 l0 = __float2bfloat16_rn(f[2 * k] - __bfloat162float(h0)), l1 = __float2bfloat16_rn(f[2 * k + 1] - __bfloat162float(h1));
        hw[k] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
        lw[k] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
    }
    *reinterpret_cast<uint2*>(y + r * 2 * C + c) = make_uint2(hw[0], hw[1]);
    *reinterpret_cast<uint2*>(y + r * 2 * C + C + c) = make_uint2(lw[0], lw[1]);
}

extern "C" int vqb_split_hi_lo(const float* x, void* y, int64_t rows, int C, void* stream) {
    VQB_CHECK_ARG(x && y && rows > 0 && C > 0 && C % 4 == 0, "split_hi_lo: bad arguments");
    const int64_t n = rows * (C / 4);
    split_hi_lo_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, as_stream(stream)>>>(x, (bf16*)y, rows, C);
    VQB_CHECK_LAUNCH("split_hi_lo");
    return VQB_OK;
}

extern "C" int vqb_convert(const void* x, int in_dtype, void* y, int out_dtype, int64_t n, void* stream) {
    VQB_CHECK_ARG(x && y && n >= 0, "convert: bad arguments");
    if (n == 0) return VQB_OK;
    int g = grid_for(n / 4 + 1, 256);
    VQB_DISPATCH_1(in_dtype, TI, VQB_DISPATCH_1(out_dtype, TO, (convert_kernel<TI, TO><<<g, 256, 0, as_stream(stream)>>>(
                                                                   (const TI*)x, (TO*)y, n));))
    VQB_CHECK_LAUNCH("convert");
    return VQB_OK;
}

// ---------------------------------------------------------------------------------------------------
// weight packing
// ---------------------------------------------------------------------------------------------------
template <typename TO>
__global__ void pack_weight_kernel(const float* __restrict__ w, TO* __restrict__ wp, int mode, int Co, int Ci, int KH,
                                   int KW, float scale) {
    int64_t total = (int64_t)Co * Ci * KH * KW;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        // i enumerates the OUTPUT layout so that writes are coalesced
        int co, ci, kh, kw;
        if (mode == 0) {  // [(kh,kw,ci)][co]
            co = (int)(i % Co); int64_t r = i / Co; ci = (int)(r % Ci); r /= Ci; kw = (int)(r % KW); kh = (int)(r / KW);
        } else if (mode == 1) {  // [(kh',kw',co)][ci], kh' = KH-1-kh
            ci = (int)(i % Ci); int64_t r = i / Ci; co = (int)(r % Co); r /= Co; kw = KW - 1 - (int)(r % KW); kh = KH - 1 - (int)(r / KW);
        } else if (mode == 2) {  // [co][(kh,kw,ci)]
            ci = (int)(i % Ci); int64_t r = i / Ci; kw = (int)(r % KW); r /= KW; kh = (int)(r % KH); co = (int)(r / KH);
        } else {  // mode 3: [ci][(kh',kw',co)]
            co = (int)(i % Co); int64_t r = i / Co; kw = KW - 1 - (int)(r % KW); r /= KW; kh = KH - 1 - (int)(r % KH); ci = (int)(r / KH);
        }
        float v = w[(((int64_t)co * Ci + ci) * KH + kh) * KW + kw] * scale;
        st1(wp + i, v);
    }
}

// modes 4 / 5: the K-major layouts of modes 2 / 3 with the reduction dimension zero-padded to 64 (narrow 3x3 heads run as
// a 64-channel 1x1 implicit GEMM on the im2col tensor of vqb_im2col3x3_narrow)
template <typename TO>
__global__ void pack_weight_pad64_kernel(const float* __restrict__ w, TO* __restrict__ wp, int mode, int Co, int Ci, int KH,
                                         int KW, float scale) {
    const int rows = (mode == 4) ? Co : Ci, inner = (mode == 4) ? Ci : Co, kreal = KH * KW * inner;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < (int64_t)rows * 64; i += (int64_t)gridDim.x * blockDim.x) {
        int j = (int)(i % 64), r = (int)(i / 64);
        float v = 0.f;
        if (j < kreal) {
            int c = j % inner, tap = j / inner, kw = tap % KW, kh = tap / KW;
            if (mode == 4) v = w[(((int64_t)r * Ci + c) * KH + kh) * KW + kw];                                 // [co][(kh,kw,ci)]
            else v = w[(((int64_t)c * Ci + r) * KH + (KH - 1 - kh)) * KW + (KW - 1 - kw)];                     // [ci][(kh',kw',co)]
        }
        st1(wp + i, v * scale);
    }
}

extern "C" int vqb_pack_conv_weight(const float* w, void* wp, int out_dtype, int mode, int Co, int Ci, int KH, int KW,
                                    float scale, void* stream) {
    VQB_CHECK_ARG(w && wp && mode >= 0 && mode <= 5 && Co > 0 && Ci > 0 && KH > 0 && KW > 0, "pack_conv_weight: bad arguments");
    if (mode >= 4) {
        VQB_CHECK_ARG(KH * KW * ((mode == 4) ? Ci : Co) <= 64, "pack_conv_weight: padded modes need KH*KW*C <= 64");
        int gp = grid_for((int64_t)((mode == 4) ? Co : Ci) * 64, 256);
        VQB_DISPATCH_1(out_dtype, TO, (pack_weight_pad64_kernel<TO><<<gp, 256, 0, as_stream(stream)>>>(w, (TO*)wp, mode, Co, Ci, KH, KW, scale));)
        VQB_CHECK_LAUNCH("pack_conv_weight(pad64)");
        return VQB_OK;
    }
    int64_t total = (int64_t)Co * Ci * KH * KW;
    int g = grid_for(total, 256);
    VQB_DISPATCH_1(out_dtype, TO, (pack_weight_kernel<TO><<<g, 256, 0, as_stream(stream)>>>(w, (TO*)wp, mode, Co, Ci, KH, KW, scale));)
    VQB_CHECK_LAUNCH("pack_conv_weight");
    return VQB_OK;
}

// ---- all kernel-layout weight copies of a model in ONE launch -------------------------------------------------------------
// After every optimizer step each convolution weight is re-packed into the layouts its kernels read (K-major bf16 for the
// forward, flipped / transposed for dgrad, ...): ~100 launches of a few microseconds each per step.  The batched form walks a
// descriptor table; element i of the concatenated outputs finds its descriptor by binary search on the start offsets.
struct PackDesc {
    const float* w;      // source [Co][Ci][KH][KW] fp32
    void* wp;            // destination
    int mode, bf16, co, ci, kh, kw;
    float scale;
    int pad_;
    long long start;     // first element of this entry in the concatenated index space
};

constexpr int PACK_T = 32;           // tile edge (co and ci) of the tiled form
constexpr int PACK_CHUNK = 4096;     // output elements per CTA; a descriptor's `start` is a multiple of it (the host pads)

__global__ void pack_weights_batched_kernel(const PackDesc* __restrict__ d, int n, long long total) {
    // one descriptor per CTA: found once by thread 0 (binary search on the chunk-aligned start offsets)
    __shared__ int which;
    const long long base = (long long)blockIdx.x * PACK_CHUNK;
    if (threadIdx.x == 0) {
        int lo = 0, hi = n - 1;
        while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (d[mid].start <= base) lo = mid; else hi = mid - 1; }
        which = lo;
    }
    __syncthreads();
    const PackDesc e = d[which];
    const int Co = e.co, Ci = e.ci, KH = e.kh, KW = e.kw;
    if (e.mode < 4 && KH * KW <= 9) {
        // Tiled form: every kernel layout is a permutation of [Co][Ci][taps] that moves co or ci to the innermost position, i.e.
        // a gather at a stride of taps * 4 (or Ci * taps * 4) bytes when written element by element (0.40 / 1.25 ms per step for
        // the 42 M / 70 M weights of the VQ-VAE / VQGAN configurations).  Here a CTA owns 32 co x 32 ci x all taps: the source is
        // read as 32 contiguous runs of 32 * taps floats into shared memory and written out in the destination's order, 32
        // consecutive co (or ci) per warp.  The CTAs of a descriptor (one per 4096 output elements) stride over its tiles.
        __shared__ float tile[PACK_T][PACK_T * 9 + 1];
        const int T = KH * KW;
        const long long next = (which + 1 < n) ? d[which + 1].start : total;
        const int nch = (int)((next - e.start) / PACK_CHUNK), t0 = (int)((base - e.start) / PACK_CHUNK);
        const int tci = (Ci + PACK_T - 1) / PACK_T, ntiles = ((Co + PACK_T - 1) / PACK_T) * tci;
        const bool flip = (e.mode == 1 || e.mode == 3);
        for (int tl = t0; tl < ntiles; tl += nch) {
            const int co0 = (tl / tci) * PACK_T, ci0 = (tl % tci) * PACK_T;
            const int nco = min(PACK_T, Co - co0), nci = min(PACK_T, Ci - ci0);
            const int run = nci * T;
            __syncthreads();                                               // the previous tile has been written out
            // (loop nests instead of a flat index: no integer division by run-time extents anywhere -- the element-wise form spent
            // more time on its five div / mod pairs per element than on memory)
            const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
            for (int r = warp; r < nco; r += nw) {
                const float* src = e.w + ((long long)(co0 + r) * Ci + ci0) * T;
                for (int k = lane; k < run; k += 32) tile[r][k] = src[k];
            }
            __syncthreads();
            const bool inner_co = (e.mode == 0 || e.mode == 3);            // which of co / ci is innermost in the destination
            const int n_in = inner_co ? nco : nci, n_out = inner_co ? nci : nco;
            if (lane < n_in)
                for (int tp = 0; tp < T; ++tp) {                           // tp: tap position in the DESTINATION
                    const int ts = flip ? T - 1 - tp : tp;
                    for (int o = warp; o < n_out; o += nw) {
                        const int col = inner_co ? lane : o, cil = inner_co ? o : lane;
                        const float v = tile[col][cil * T + ts] * e.scale;
                        const int co = co0 + col, ci = ci0 + cil;
                        long long j;
                        if (e.mode == 0) j = ((long long)tp * Ci + ci) * Co + co;
                        else if (e.mode == 1) j = ((long long)tp * Co + co) * Ci + ci;
                        else if (e.mode == 2) j = ((long long)co * T + tp) * Ci + ci;
                        else j = ((long long)ci * T + tp) * Co + co;
                        if (e.bf16) reinterpret_cast<bf16*>(e.wp)[j] = __float2bfloat16_rn(v);
                        else reinterpret_cast<float*>(e.wp)[j] = v;
                    }
                }
        }
        return;
    }
    const long long count = (e.mode < 4) ? (long long)Co * Ci * KH * KW : 64ll * ((e.mode == 4) ? Co : Ci);
    for (int t = threadIdx.x; t < PACK_CHUNK; t += blockDim.x) {
        const long long j = base - e.start + t;
        if (j >= count) break;
        float v = 0.f;
        if (e.mode < 4) {
            int co, ci, kh, kw;
            if (e.mode == 0) { co = (int)(j % Co); long long r = j / Co; ci = (int)(r % Ci); r /= Ci; kw = (int)(r % KW); kh = (int)(r / KW); }
            else if (e.mode == 1) { ci = (int)(j % Ci); long long r = j / Ci; co = (int)(r % Co); r /= Co; kw = KW - 1 - (int)(r % KW); kh = KH - 1 - (int)(r / KW); }
            else if (e.mode == 2) { ci = (int)(j % Ci); long long r = j / Ci; kw = (int)(r % KW); r /= KW; kh = (int)(r % KH); co = (int)(r / KH); }
            else { co = (int)(j % Co); long long r = j / Co; kw = KW - 1 - (int)(r % KW); r /= KW; kh = KH - 1 - (int)(r % KH); ci = (int)(r / KH); }
            v = e.w[(((long long)co * Ci + ci) * KH + kh) * KW + kw];
        } else {
            const int inner = (e.mode == 4) ? Ci : Co, kreal = KH * KW * inner;
            const int jj = (int)(j % 64), r = (int)(j / 64);
            if (jj < kreal) {
                const int c = jj % inner, tap = jj / inner, kw = tap % KW, kh = tap / KW;
                v = (e.mode == 4) ? e.w[(((long long)r * Ci + c) * KH + kh) * KW + kw]
                                  : e.w[(((long long)c * Ci + r) * KH + (KH - 1 - kh)) * KW + (KW - 1 - kw)];
            }
        }
        v *= e.scale;
        if (e.bf16) reinterpret_cast<bf16*>(e.wp)[j] = __float2bfloat16_rn(v);
        else reinterpret_cast<float*>(e.wp)[j] = v;
    }
}

extern "C" size_t vqb_pack_desc_bytes(void) { return sizeof(PackDesc); }

extern "C" int vqb_pack_conv_weights_batched(const void* desc_table, int n_desc, int64_t total_elems, void* stream) {
    VQB_CHECK_ARG(desc_table && n_desc > 0 && total_elems > 0 && total_elems % PACK_CHUNK == 0,
                  "pack_conv_weights_batched: bad arguments (total_elems and every start must be multiples of %d)", PACK_CHUNK);
    pack_weights_batched_kernel<<<(unsigned)(total_elems / PACK_CHUNK), 256, 0, as_stream(stream)>>>((const PackDesc*)desc_table, n_desc,
                                                                                                    (long long)total_elems);
    VQB_CHECK_LAUNCH("pack_conv_weights_batched");
    return VQB_OK;
}

// 16-byte vector access: 8 bf16 or 4 fp32 per thread
template <typename T> struct V16 { static constexpr int N = 16 / sizeof(T); };
template <typename T>
__device__ __forceinline__ void ld16(const T* p, float* v) {
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
__device__ __forceinline__ void st16(T* p, const float* v) {
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

// P[n,h,w, j] (64 channels, TO) = x[n, h+kh-1, w+kw-1, c] for j = (kh*3+kw)*C + c < 9*C, else 0   (3x3, pad 1, C <= 7)
template <typename TI, typename TO, int CT>
__global__ void im2col3x3_narrow_kernel(const TI* __restrict__ x, TO* __restrict__ P, int N, int H, int W, int Crt, int G) {
    const int C = (CT > 0) ? CT : Crt;                        // compile-time channel count (3 = RGB) avoids runtime div/mod
    const int64_t total = (int64_t)N * H * W * G;            // one thread per (pixel, group of 8 output columns); G groups written
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int grp = (int)(i % G); int64_t pix = i / G;
        int w = (int)(pix % W); int64_t r = pix / W; int h = (int)(r % H); int n = (int)(r / H);
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            int j = grp * 8 + u;
            float val = 0.f;
            if (j < 9 * C) {
                int tap = j / C, c = j - tap * C, kh = tap / 3, kw = tap - kh * 3;
                int ih = h + kh - 1, iw = w + kw - 1;
                if (ih >= 0 && ih < H && iw >= 0 && iw < W) val = ld1(x + (((int64_t)n * H + ih) * W + iw) * C + c);
            }
            v[u] = val;
        }
        TO* dst = P + pix * 64 + grp * 8;
        if constexpr (sizeof(TO) == 2) {
            st16<TO>(dst, v);                                    // one 16-byte store per thread
        } else {
            st4(dst, make_float4(v[0], v[1], v[2], v[3]));
            st4(dst + 4, make_float4(v[4], v[5], v[6], v[7]));
        }
    }
}

extern "C" int vqb_im2col3x3_narrow(const void* x, int x_dtype, void* P, int p_dtype, int N, int H, int W, int C, int write_all,
                                    void* stream) {
    VQB_CHECK_ARG(x && P && N > 0 && H > 0 && W > 0 && C > 0 && 9 * C <= 64, "im2col3x3_narrow: bad arguments (need 9*C <= 64)");
    const int G = write_all ? 8 : (9 * C + 7) / 8;
    int g = grid_for((int64_t)N * H * W * G, 256);
    if (C == 3) {
        VQB_DISPATCH_1(x_dtype, TI, VQB_DISPATCH_1(p_dtype, TO, (im2col3x3_narrow_kernel<TI, TO, 3><<<g, 256, 0, as_stream(stream)>>>(
                                                                    (const TI*)x, (TO*)P, N, H, W, C, G));))
    } else {
        VQB_DISPATCH_1(x_dtype, TI, VQB_DISPATCH_1(p_dtype, TO, (im2col3x3_narrow_kernel<TI, TO, 0><<<g, 256, 0, as_stream(stream)>>>(
                                                                    (const TI*)x, (TO*)P, N, H, W, C, G));))
    }
    VQB_CHECK_LAUNCH("im2col3x3_narrow");
    return VQB_OK;
}

// ACC: dw += (the destination is the parameter's gradient view inside the optimizer's flat buffer -- no separate autograd
// accumulation pass); REZERO: the packed partial-sum buffer is cleared behind the read, so that the persistent per-weight buffer is
// ready for the next weight-gradient launch without a fill kernel.
template <bool ACC, bool REZERO>
__global__ void unpack_wgrad_kernel(float* __restrict__ dwp, float* __restrict__ dw, int Co, int Ci, int KH, int KW, float scale) {
    // tile-transpose between [(tap,ci)][co] and [co][ci][tap]: handle per tap a [Ci][Co] -> [Co][Ci] transpose
    __shared__ float tile[32][33];
    int tap = blockIdx.z;
    int T = KH * KW;
    int ci0 = blockIdx.y * 32, co0 = blockIdx.x * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        int ci = ci0 + j, co = co0 + threadIdx.x;
        if (ci < Ci && co < Co) {
            const int64_t o = ((int64_t)tap * Ci + ci) * Co + co;
            tile[j][threadIdx.x] = dwp[o];
            if constexpr (REZERO) dwp[o] = 0.f;
        }
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        int co = co0 + j, ci = ci0 + threadIdx.x;
        if (ci < Ci && co < Co) {
            const int64_t o = ((int64_t)co * Ci + ci) * T + tap;
            const float v = tile[threadIdx.x][j] * scale;
            if constexpr (ACC) dw[o] += v; else dw[o] = v;
        }
    }
}

extern "C" int vqb_unpack_conv_wgrad_acc(float* dwp, float* dw, int Co, int Ci, int KH, int KW, float scale, int accumulate, int rezero,
                                         void* stream);
extern "C" int vqb_unpack_conv_wgrad(const float* dwp, float* dw, int Co, int Ci, int KH, int KW, float scale, void* stream) {
    return vqb_unpack_conv_wgrad_acc(const_cast<float*>(dwp), dw, Co, Ci, KH, KW, scale, 0, 0, stream);
}

// All taps of a 32 co x 32 ci tile per CTA: the destination [co][ci][tap] is then written in runs of 32 * T contiguous floats per co
// (the per-tap kernel above writes single floats at a stride of T: ncu 1.2 ms per step for 0.1 ms worth of bytes).  T <= 9.
template <bool ACC, bool REZERO>
__global__ void unpack_wgrad_alltaps_kernel(float* __restrict__ dwp, float* __restrict__ dw, int Co, int Ci, int T, float scale) {
    __shared__ float tile[9][32][33];
    const int ci0 = blockIdx.y * 32, co0 = blockIdx.x * 32;
    // all 9 x 4 loads of a thread are independent and issued back to back (small weights launch only a few CTAs)
    float v[9][4];
#pragma unroll
    for (int tap = 0; tap < 9; ++tap)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            const int ci = ci0 + threadIdx.y + jj * 8, co = co0 + threadIdx.x;
            v[tap][jj] = (tap < T && ci < Ci && co < Co) ? dwp[((int64_t)tap * Ci + ci) * Co + co] : 0.f;
        }
#pragma unroll
    for (int tap = 0; tap < 9; ++tap)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            const int ci = ci0 + threadIdx.y + jj * 8, co = co0 + threadIdx.x;
            if (tap < T) tile[tap][threadIdx.y + jj * 8][threadIdx.x] = v[tap][jj];
            if constexpr (REZERO) { if (tap < T && ci < Ci && co < Co) dwp[((int64_t)tap * Ci + ci) * Co + co] = 0.f; }
        }
    __syncthreads();
    const int nci = (Ci - ci0 < 32) ? Ci - ci0 : 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int co = co0 + j;
        if (co >= Co) continue;
        float* row = dw + ((int64_t)co * Ci + ci0) * T;                 // nci * T contiguous floats
        for (int e = threadIdx.x; e < nci * T; e += 32) {
            const int cil = e / T, tap = e - cil * T;
            const float v = tile[tap][cil][j] * scale;
            if constexpr (ACC) row[e] += v; else row[e] = v;
        }
    }
}

extern "C" int vqb_unpack_conv_wgrad_acc(float* dwp, float* dw, int Co, int Ci, int KH, int KW, float scale, int accumulate, int rezero,
                                         void* stream) {
    VQB_CHECK_ARG(dwp && dw && Co > 0 && Ci > 0 && KH > 0 && KW > 0, "unpack_conv_wgrad_acc: bad arguments");
    cudaStream_t st = as_stream(stream);
    const int T = KH * KW;
    if (T > 1 && T <= 9) {
        dim3 grid((Co + 31) / 32, (Ci + 31) / 32), block(32, 8);
        if (accumulate && rezero) unpack_wgrad_alltaps_kernel<true, true><<<grid, block, 0, st>>>(dwp, dw, Co, Ci, T, scale);
        else if (accumulate) unpack_wgrad_alltaps_kernel<true, false><<<grid, block, 0, st>>>(dwp, dw, Co, Ci, T, scale);
        else if (rezero) unpack_wgrad_alltaps_kernel<false, true><<<grid, block, 0, st>>>(dwp, dw, Co, Ci, T, scale);
        else unpack_wgrad_alltaps_kernel<false, false><<<grid, block, 0, st>>>(dwp, dw, Co, Ci, T, scale);
        VQB_CHECK_LAUNCH("unpack_conv_wgrad_acc");
        return VQB_OK;
    }
    dim3 grid((Co + 31) / 32, (Ci + 31) / 32, T), block(32, 8);
    if (accumulate && rezero) unpack_wgrad_kernel<true, true><<<grid, block, 0, st>>>(dwp, dw, Co, Ci, KH, KW, scale);
    else if (accumulate) unpack_wgrad_kernel<true, false><<<grid, block, 0, st>>>(dwp, dw, Co, Ci, KH, KW, scale);
    else if (rezero) unpack_wgrad_kernel<false, true><<<grid, block, 0, st>>>(dwp, dw, Co, Ci, KH, KW, scale);
    else unpack_wgrad_kernel<false, false><<<grid, block, 0, st>>>(dwp, dw, Co, Ci, KH, KW, scale);
    VQB_CHECK_LAUNCH("unpack_conv_wgrad_acc");
    return VQB_OK;
}

// ---------------------------------------------------------------------------------------------------
// 2x resampling
// ---------------------------------------------------------------------------------------------------
// VEC = 0: 16-byte vectors (C % V16<T>::N == 0), iterating over the LOW-resolution grid: four 16-byte loads -> one store
// (down2) or one load -> four stores (up2); VEC = 1: scalar fallback for odd channel counts
template <typename T, int VEC>
__global__ void down2_kernel(const T* __restrict__ x, T* __restrict__ y, int N, int H, int W, int C, float scale) {
    // y [N,H,W,C], x [N,2H,2W,C]
    constexpr int V = (VEC == 0) ? V16<T>::N : 1;
    int Cv = C / V;
    int64_t total = (int64_t)N * H * W * Cv;
    int64_t rowx = (int64_t)2 * W * C;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int cv = (int)(i % Cv); int64_t r = i / Cv; int w = (int)(r % W); r /= W; int h = (int)(r % H); int n = (int)(r / H);
        const T* p = x + (((int64_t)n * 2 * H + 2 * h) * 2 * W + 2 * w) * C + cv * V;
        if constexpr (VEC == 0) {
            float a[V], b[V], c[V], d[V], o[V];
            ld16<T>(p, a); ld16<T>(p + C, b); ld16<T>(p + rowx, c); ld16<T>(p + rowx + C, d);
#pragma unroll
            for (int j = 0; j < V; ++j) o[j] = (a[j] + b[j] + c[j] + d[j]) * scale;
            st16<T>(y + i * V, o);
        } else {
            st1(y + i, (ld1(p) + ld1(p + C) + ld1(p + rowx) + ld1(p + rowx + C)) * scale);
        }
    }
}

template <typename T, int VEC>
__global__ void up2_kernel(const T* __restrict__ x, T* __restrict__ y, int N, int H, int W, int C, float scale) {
    // x [N,H,W,C], y [N,2H,2W,C]; iterate over the input
    constexpr int V = (VEC == 0) ? V16<T>::N : 1;
    int Cv = C / V;
    int64_t total = (int64_t)N * H * W * Cv;
    int64_t rowy = (int64_t)2 * W * C;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int cv = (int)(i % Cv); int64_t r = i / Cv; int w = (int)(r % W); r /= W; int h = (int)(r % H); int n = (int)(r / H);
        T* q = y + (((int64_t)n * 2 * H + 2 * h) * 2 * W + 2 * w) * C + cv * V;
        if constexpr (VEC == 0) {
            float a[V];
            ld16<T>(x + i * V, a);
#pragma unroll
            for (int j = 0; j < V; ++j) a[j] *= scale;
            st16<T>(q, a); st16<T>(q + C, a); st16<T>(q + rowy, a); st16<T>(q + rowy + C, a);
        } else {
            float a = ld1(x + i) * scale;
            st1(q, a); st1(q + C, a); st1(q + rowy, a); st1(q + rowy + C, a);
        }
    }
}

extern "C" int vqb_down2(const void* x, void* y, int dtype, int N, int H, int W, int C, float scale, void* stream) {
    VQB_CHECK_ARG(x && y && N > 0 && H > 0 && W > 0 && C > 0, "down2: bad arguments");
    int64_t total = (int64_t)N * H * W * C;
    const int vw = (dtype == VQB_BF16) ? 8 : 4;
    if (C % vw == 0) {
        int g = grid_for(total / vw, 256);
        VQB_DISPATCH_1(dtype, T, (down2_kernel<T, 0><<<g, 256, 0, as_stream(stream)>>>((const T*)x, (T*)y, N, H, W, C, scale));)
    } else {
        int g = grid_for(total, 256);
        VQB_DISPATCH_1(dtype, T, (down2_kernel<T, 1><<<g, 256, 0, as_stream(stream)>>>((const T*)x, (T*)y, N, H, W, C, scale));)
    }
    VQB_CHECK_LAUNCH("down2");
    return VQB_OK;
}

extern "C" int vqb_up2(const void* x, void* y, int dtype, int N, int H, int W, int C, float scale, void* stream) {
    VQB_CHECK_ARG(x && y && N > 0 && H > 0 && W > 0 && C > 0, "up2: bad arguments");
    int64_t total = (int64_t)N * H * W * C;                    // threads iterate over the input grid
    const int vw = (dtype == VQB_BF16) ? 8 : 4;
    if (C % vw == 0) {
        int g = grid_for(total / vw, 256);
        VQB_DISPATCH_1(dtype, T, (up2_kernel<T, 0><<<g, 256, 0, as_stream(stream)>>>((const T*)x, (T*)y, N, H, W, C, scale));)
    } else {
        int g = grid_for(total, 256);
        VQB_DISPATCH_1(dtype, T, (up2_kernel<T, 1><<<g, 256, 0, as_stream(stream)>>>((const T*)x, (T*)y, N, H, W, C, scale));)
    }
    VQB_CHECK_LAUNCH("up2");
    return VQB_OK;
}

// ---------------------------------------------------------------------------------------------------
// loss reductions and their gradients
// ---------------------------------------------------------------------------------------------------
template <typename TA, typename TB>
__global__ void diff_sums_kernel(const TA* __restrict__ a, const TB* __restrict__ b, double* __restrict__ out, int64_t n) {
    float s2 = 0.f, s1 = 0.f;
    double d2 = 0.0, d1 = 0.0;
    int cnt = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float d = ld1(a + i) - ld1(b + i);
        s2 += d * d;
        s1 += fabsf(d);
        if (++cnt == 64) { d2 += s2; d1 += s1; s2 = s1 = 0.f; cnt = 0; }
    }
    d2 += s2; d1 += s1;
    d2 = warp_sum(d2); d1 = warp_sum(d1);
    __shared__ double sh[2][8];
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) { sh[0][wid] = d2; sh[1][wid] = d1; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t2 = 0, t1 = 0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { t2 += sh[0][i]; t1 += sh[1][i]; }
        atomicAdd(out, t2);
        atomicAdd(out + 1, t1);
    }
}

// ---- SSIM (evaluation, vqvae/model.py:495, 529-530, 549: torchmetrics StructuralSimilarityIndexMeasure with its defaults) -----------
// Published definition (torchmetrics functional/image/ssim.py, gaussian_kernel = True, kernel_size 11, sigma 1.5, k1 0.01,
// k2 0.03): both images are reflect-padded by 5, filtered with the separable Gaussian window, and the border the padding
// touched is cropped again -- i.e. the windows are the VALID 11 x 11 windows of the un-padded image.  Per window:
//   c1 = (k1 R)^2, c2 = (k2 R)^2, R = data range;  mu = E[x], mu' = E[y], s = max(E[x^2] - mu^2, 0), s' likewise, sxy = E[xy] - mu mu'
//   ssim = ((2 mu mu' + c1)(2 sxy + c2)) / ((mu^2 + mu'^2 + c1)(s + s' + c2))
// and an image's value is the mean over (C, H-10, W-10).  One CTA = one 16 x 16 tile of windows of one (image, channel) plane:
// the 26 x 26 input patches in shared memory, a horizontal pass into shared memory (five moments), a vertical pass in registers,
// fp64 sum per image.  NCHW fp32 inputs (the test-time images).
constexpr int SSIM_T = 16, SSIM_K = 11, SSIM_P = SSIM_T + SSIM_K - 1;
__global__ void __launch_bounds__(256) ssim_sums_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ range,
                                                         double* __restrict__ out, int C, int H, int W, float k1, float k2) {
    __shared__ float pa[SSIM_P][SSIM_P + 1], pb[SSIM_P][SSIM_P + 1];
    __shared__ float hz[5][SSIM_P][SSIM_T + 1];
    __shared__ float g[SSIM_K];
    __shared__ double red[8];
    const int plane = blockIdx.z, n = plane / C;
    const int x0 = blockIdx.x * SSIM_T, y0 = blockIdx.y * SSIM_T;
    const int OW = W - SSIM_K + 1, OH = H - SSIM_K + 1;
    const float* pa_g = a + (int64_t)plane * H * W;
    const float* pb_g = b + (int64_t)plane * H * W;
    if (threadIdx.x == 0) {
        float w[SSIM_K], sum = 0.f;
        for (int i = 0; i < SSIM_K; ++i) { const float d = (float)(i - SSIM_K / 2) / 1.5f; w[i] = expf(-0.5f * d * d); sum += w[i]; }
        for (int i = 0; i < SSIM_K; ++i) g[i] = w[i] / sum;
    }
    for (int i = threadIdx.x; i < SSIM_P * SSIM_P; i += 256) {
        const int r = i / SSIM_P, c = i - r * SSIM_P;
        const int y = y0 + r, x = x0 + c;
        const bool in = y < H && x < W;
        pa[r][c] = in ? pa_g[(int64_t)y * W + x] : 0.f;
        pb[r][c] = in ? pb_g[(int64_t)y * W + x] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < SSIM_P * SSIM_T; i += 256) {                    // horizontal pass: rows 0..25, window columns 0..15
        const int r = i / SSIM_T, c = i - r * SSIM_T;
        float m0 = 0.f, m1 = 0.f, m2 = 0.f, m3 = 0.f, m4 = 0.f;
#pragma unroll
        for (int t = 0; t < SSIM_K; ++t) {
            const float u = pa[r][c + t], v = pb[r][c + t], w = g[t];
            m0 = fmaf(w, u, m0); m1 = fmaf(w, v, m1); m2 = fmaf(w, u * u, m2); m3 = fmaf(w, v * v, m3); m4 = fmaf(w, u * v, m4);
        }
        hz[0][r][c] = m0; hz[1][r][c] = m1; hz[2][r][c] = m2; hz[3][r][c] = m3; hz[4][r][c] = m4;
    }
    __syncthreads();
    const int ty = threadIdx.x / SSIM_T, tx = threadIdx.x % SSIM_T;
    double val = 0.0;
    if (y0 + ty < OH && x0 + tx < OW) {
        float m[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int t = 0; t < SSIM_K; ++t)
#pragma unroll
            for (int q = 0; q < 5; ++q) m[q] = fmaf(g[t], hz[q][ty + t][tx], m[q]);
        const float R = range[0], c1 = (k1 * R) * (k1 * R), c2 = (k2 * R) * (k2 * R);
        const float mua2 = m[0] * m[0], mub2 = m[1] * m[1], muab = m[0] * m[1];
        const float sa = fmaxf(m[2] - mua2, 0.f), sb = fmaxf(m[3] - mub2, 0.f), sab = m[4] - muab;
        const float upper = 2.f * sab + c2, lower = sa + sb + c2;
        val = (double)(((2.f * muab + c1) * upper) / ((mua2 + mub2 + c1) * lower));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = val;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < 8; ++i) t += red[i];
        atomicAdd(out + n, t);
    }
}

extern "C" int vqb_ssim_sums(const float* preds, const float* target, const float* data_range, double* per_image_sum, int N, int C, int H,
                             int W, float k1, float k2, void* stream) {
    VQB_CHECK_ARG(preds && target && data_range && per_image_sum && N > 0 && C > 0, "ssim_sums: bad arguments");
    VQB_CHECK_ARG(H >= SSIM_K && W >= SSIM_K && (int64_t)N * C <= 65535, "ssim_sums: needs H, W >= 11 and N * C <= 65535 (got %d x %d, %d planes)", H, W, N * C);
    dim3 grid((unsigned)ceil_div64(W - SSIM_K + 1, SSIM_T), (unsigned)ceil_div64(H - SSIM_K + 1, SSIM_T), (unsigned)(N * C));
    ssim_sums_kernel<<<grid, 256, 0, as_stream(stream)>>>(preds, target, data_range, per_image_sum, C, H, W, k1, k2);
    VQB_CHECK_LAUNCH("ssim_sums");
    return VQB_OK;
}

extern "C" int vqb_diff_sums(const void* a, int a_dtype, const void* b, int b_dtype, double* out, int64_t n, void* stream) {
    VQB_CHECK_ARG(a && b && out && n > 0, "diff_sums: bad arguments");
    int g = grid_for(n, 256, 4);
    VQB_DISPATCH_1(a_dtype, TA, VQB_DISPATCH_1(b_dtype, TB, (diff_sums_kernel<TA, TB><<<g, 256, 0, as_stream(stream)>>>(
                                                                (const TA*)a, (const TB*)b, out, n));))
    VQB_CHECK_LAUNCH("diff_sums");
    return VQB_OK;
}

template <typename TA, typename TB, typename TD>
__global__ void diff_grad_kernel(const TA* __restrict__ a, const TB* __restrict__ b, TD* __restrict__ da, float c1, float c2,
                                 const float* __restrict__ upstream, int y_tanh, int64_t n) {
    const float up2 = upstream ? upstream[0] : 1.0f, up1 = upstream ? upstream[1] : 1.0f;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float av = ld1(a + i);
        float d = av - ld1(b + i);
        float sg = (d > 0.f) ? 1.f : ((d < 0.f) ? -1.f : 0.f);
        float g = c2 * 2.0f * d * up2 + c1 * sg * up1;
        if (y_tanh) g *= (1.0f - av * av);
        st1(da + i, g);
    }
}

extern "C" int vqb_diff_grad(const void* a, int a_dtype, const void* b, int b_dtype, void* da, int da_dtype, float c1,
                             float c2, const float* upstream, int y_tanh, int64_t n, void* stream) {
    VQB_CHECK_ARG(a && b && da && n > 0, "diff_grad: bad arguments");
    int g = grid_for(n, 256);
    VQB_DISPATCH_1(a_dtype, TA, VQB_DISPATCH_1(b_dtype, TB, VQB_DISPATCH_1(da_dtype, TD,
        (diff_grad_kernel<TA, TB, TD><<<g, 256, 0, as_stream(stream)>>>((const TA*)a, (const TB*)b, (TD*)da, c1, c2, upstream, y_tanh, n));)))
    VQB_CHECK_LAUNCH("diff_grad");
    return VQB_OK;
}

template <typename TY, typename TG, typename TD>
__global__ void act_bwd_out_kernel(const TY* __restrict__ y, const TG* __restrict__ dy, TD* __restrict__ dx, int act,
                                   float alpha, float gain, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float yv = ld1(y + i), g = ld1(dy + i);
        float d;
        if (act == VQB_ACT_TANH) { float t = yv / gain; d = gain * (1.0f - t * t); }
        else if (act == VQB_ACT_LRELU) d = (yv > 0.f) ? gain : gain * alpha;     // sign(y) == sign(pre-activation)
        else if (act == VQB_ACT_RELU) d = (yv > 0.f) ? gain : 0.f;
        else d = gain;
        st1(dx + i, g * d);
    }
}

extern "C" int vqb_act_bwd_from_output(const void* y, int y_dtype, const void* dy, int dy_dtype, void* dx, int dx_dtype,
                                       int act, float alpha, float gain, int64_t n, void* stream) {
    VQB_CHECK_ARG(y && dy && dx && n > 0, "act_bwd_from_output: bad arguments");
    VQB_CHECK_ARG(act != VQB_ACT_SILU, "act_bwd_from_output: SiLU is not invertible from its output");
    int g = grid_for(n, 256);
    VQB_DISPATCH_1(y_dtype, TY, VQB_DISPATCH_1(dy_dtype, TG, VQB_DISPATCH_1(dx_dtype, TD,
        (act_bwd_out_kernel<TY, TG, TD><<<g, 256, 0, as_stream(stream)>>>((const TY*)y, (const TG*)dy, (TD*)dx, act, alpha, gain, n));)))
    VQB_CHECK_LAUNCH("act_bwd_from_output");
    return VQB_OK;
}

// ---------------------------------------------------------------------------------------------------
// column sums (bias gradients): out[c] += sum_p a[p][c]
// ---------------------------------------------------------------------------------------------------
template <typename T>
__global__ void colsum_kernel(const T* __restrict__ a, float* __restrict__ out, int64_t P, int C, int rows_per_block) {
    // blockDim.x threads cover channels (strided), blockIdx.x covers a row chunk
    int64_t p0 = (int64_t)blockIdx.x * rows_per_block;
    int64_t p1 = p0 + rows_per_block; if (p1 > P) p1 = P;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float s = 0.f;
        for (int64_t p = p0 + threadIdx.y; p < p1; p += blockDim.y) s += ld1(a + p * C + c);
        atomicAdd(out + c, s);
    }
}

// 16-byte vector form: threadIdx.x owns V consecutive channels, threadIdx.y strides over rows (8 independent loads in flight
// per thread), block partials combined in shared memory, one atomicAdd per channel per block
template <typename T>
__global__ void colsum_vec_kernel(const T* __restrict__ a, float* __restrict__ out, int64_t P, int C, int rows_per_block) {
    constexpr int V = V16<T>::N;
    extern __shared__ float shc[];                               // [ty][C]
    const int64_t p0 = (int64_t)blockIdx.x * rows_per_block;
    int64_t p1 = p0 + rows_per_block; if (p1 > P) p1 = P;
    const int c0 = threadIdx.x * V;
    float acc[V];
#pragma unroll
    for (int j = 0; j < V; ++j) acc[j] = 0.f;
#pragma unroll 8
    for (int64_t p = p0 + threadIdx.y; p < p1; p += blockDim.y) {
        float v[V];
        ld16<T>(a + p * C + c0, v);
#pragma unroll
        for (int j = 0; j < V; ++j) acc[j] += v[j];
    }
#pragma unroll
    for (int j = 0; j < V; ++j) shc[threadIdx.y * C + c0 + j] = acc[j];
    __syncthreads();
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    for (int c = tid; c < C; c += blockDim.x * blockDim.y) {
        float t = 0.f;
        for (int y = 0; y < (int)blockDim.y; ++y) t += shc[y * C + c];
        atomicAdd(out + c, t);
    }
}

extern "C" int vqb_colsum(const void* a, int a_dtype, float* out, int64_t P, int C, void* stream) {
    VQB_CHECK_ARG(a && out && P > 0 && C > 0, "colsum: bad arguments");
    {
        const int vw = (a_dtype == VQB_BF16) ? 8 : 4;
        if (C % vw == 0 && C / vw <= 256 && P >= 4096) {
            int tx = C / vw, ty = 256 / tx; if (ty < 1) ty = 1;
            int rows = (int)ceil_div64(P, (int64_t)kSMs * 8);
            if (rows < ty * 16) rows = ty * 16;
            int g = (int)ceil_div64(P, rows);
            dim3 block(tx, ty);
            size_t sm = sizeof(float) * ty * C;
            VQB_DISPATCH_1(a_dtype, T, (colsum_vec_kernel<T><<<g, block, sm, as_stream(stream)>>>((const T*)a, out, P, C, rows));)
            VQB_CHECK_LAUNCH("colsum_vec");
            return VQB_OK;
        }
    }
    int tx = C >= 128 ? 128 : (C >= 32 ? 32 : (C >= 8 ? 8 : 4));
    int ty = 256 / tx;
    int64_t want_blocks = (int64_t)kSMs * 8;
    int rows = (int)ceil_div64(P, want_blocks);
    if (rows < ty * 4) rows = ty * 4;
    int g = (int)ceil_div64(P, rows);
    dim3 block(tx, ty);
    VQB_DISPATCH_1(a_dtype, T, (colsum_kernel<T><<<g, block, 0, as_stream(stream)>>>((const T*)a, out, P, C, rows));)
    VQB_CHECK_LAUNCH("colsum");
    return VQB_OK;
}

// ---------------------------------------------------------------------------------------------------
// activation backward from the saved output + bias gradient in ONE pass (bias_act.py:143-210: dx = dy * act'(y) * gain,
// db = sum over pixels of dx).  16-byte vectors; threadIdx.x owns V consecutive channels, threadIdx.y strides over pixels.
// ---------------------------------------------------------------------------------------------------
template <typename T>
__global__ void act_bwd_bias_kernel(const T* __restrict__ y, const T* __restrict__ dy, T* __restrict__ dx, int act, float alpha,
                                    float gain, int64_t P, int C, int rows_per_block, float* __restrict__ db) {
    constexpr int V = V16<T>::N;
    extern __shared__ float shc[];                               // [ty][C] (only when db != null)
    const int64_t p0 = (int64_t)blockIdx.x * rows_per_block;
    int64_t p1 = p0 + rows_per_block; if (p1 > P) p1 = P;
    const int c0 = threadIdx.x * V;
    const float neg = (act == VQB_ACT_LRELU) ? gain * alpha : ((act == VQB_ACT_RELU) ? 0.f : gain);
    const float inv_gain = 1.0f / gain;
    float acc[V];
#pragma unroll
    for (int j = 0; j < V; ++j) acc[j] = 0.f;
#pragma unroll 4
    for (int64_t p = p0 + threadIdx.y; p < p1; p += blockDim.y) {
        float yv[V], g[V], o[V];
        ld16<T>(y + p * C + c0, yv);
        ld16<T>(dy + p * C + c0, g);
#pragma unroll
        for (int j = 0; j < V; ++j) {
            float d;
            if (act == VQB_ACT_TANH) { float t = yv[j] * inv_gain; d = gain * (1.0f - t * t); }
            else d = (yv[j] > 0.f) ? gain : neg;
            o[j] = g[j] * d;
            if constexpr (sizeof(T) == 2) o[j] = __bfloat162float(__float2bfloat16_rn(o[j]));   // db sums what is stored
            acc[j] += o[j];
        }
        st16<T>(dx + p * C + c0, o);
    }
    if (db == nullptr) return;
#pragma unroll
    for (int j = 0; j < V; ++j) shc[threadIdx.y * C + c0 + j] = acc[j];
    __syncthreads();
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    for (int c = tid; c < C; c += blockDim.x * blockDim.y) {
        float t = 0.f;
        for (int yy = 0; yy < (int)blockDim.y; ++yy) t += shc[yy * C + c];
        atomicAdd(db + c, t);
    }
}


// bf16 form on a per-thread cp.async ring (the scheme of the GroupNorm kernels): y and dy stream through a private DEPTH-deep
// ring in shared memory, no registers held in flight.  The plain-load kernel above ran at ~55 % of the HBM roof on the
// discriminator's activations (ncu launch list: 3.4 ms per VQGAN step for 13.5 GB).
constexpr int ACT_RING_DEPTH = 8;
__device__ __forceinline__ void act_cp_async16(void* smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__global__ void act_bwd_bias_async_kernel(const bf16* __restrict__ y, const bf16* __restrict__ dy, bf16* __restrict__ dx, int act, float alpha,
                                          float gain, int64_t P, int C, int rows_per_block, float* __restrict__ db) {
    constexpr int V = 8, D = ACT_RING_DEPTH;
    extern __shared__ __align__(16) unsigned char act_smem[];
    uint4* ring = reinterpret_cast<uint4*>(act_smem);            // [D][2][nthreads]
    float* shc = reinterpret_cast<float*>(act_smem);             // reused after the loop: [ty][C]
    const int tx = blockDim.x, ty = blockDim.y, nthr = tx * ty;
    const int tid = threadIdx.y * tx + threadIdx.x;
    const int64_t p0 = (int64_t)blockIdx.x * rows_per_block;
    int64_t p1 = p0 + rows_per_block; if (p1 > P) p1 = P;
    const int c0 = threadIdx.x * V;
    const float neg = (act == VQB_ACT_LRELU) ? gain * alpha : ((act == VQB_ACT_RELU) ? 0.f : gain);
    const float inv_gain = 1.0f / gain;
    const int64_t first = p0 + threadIdx.y;
    const int niter = first < p1 ? (int)((p1 - first + ty - 1) / ty) : 0;
    float acc[V];
#pragma unroll
    for (int j = 0; j < V; ++j) acc[j] = 0.f;
#pragma unroll
    for (int s_ = 0; s_ < D; ++s_) {
        if (s_ < niter) {
            const int64_t off = (first + (int64_t)s_ * ty) * C + c0;
            act_cp_async16(&ring[(s_ * 2 + 0) * nthr + tid], y + off);
            act_cp_async16(&ring[(s_ * 2 + 1) * nthr + tid], dy + off);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    int slot = 0;
    for (int it = 0; it < niter; ++it) {
        asm volatile("cp.async.wait_group %0;" ::"n"(D - 1) : "memory");
        const uint4 uy = ring[(slot * 2 + 0) * nthr + tid], ug = ring[(slot * 2 + 1) * nthr + tid];
        const __nv_bfloat162* hy = reinterpret_cast<const __nv_bfloat162*>(&uy);
        const __nv_bfloat162* hg = reinterpret_cast<const __nv_bfloat162*>(&ug);
        uint4 uo;
        __nv_bfloat162* ho = reinterpret_cast<__nv_bfloat162*>(&uo);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 yv = __bfloat1622float2(hy[j]), g = __bfloat1622float2(hg[j]);
            float d0, d1;
            if (act == VQB_ACT_TANH) { const float t0 = yv.x * inv_gain, t1 = yv.y * inv_gain; d0 = gain * (1.0f - t0 * t0); d1 = gain * (1.0f - t1 * t1); }
            else { d0 = (yv.x > 0.f) ? gain : neg; d1 = (yv.y > 0.f) ? gain : neg; }
            ho[j] = __floats2bfloat162_rn(g.x * d0, g.y * d1);
            const float2 o = __bfloat1622float2(ho[j]);              // db sums what is stored
            acc[2 * j] += o.x; acc[2 * j + 1] += o.y;
        }
        const int64_t off = (first + (int64_t)it * ty) * C + c0;
        *reinterpret_cast<uint4*>(dx + off) = uo;
        if (it + D < niter) {
            const int64_t offn = (first + (int64_t)(it + D) * ty) * C + c0;
            act_cp_async16(&ring[(slot * 2 + 0) * nthr + tid], y + offn);
            act_cp_async16(&ring[(slot * 2 + 1) * nthr + tid], dy + offn);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        slot = (slot + 1 == D) ? 0 : slot + 1;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (db == nullptr) return;
    __syncthreads();                                             // every thread is done with its ring before the reuse
#pragma unroll
    for (int j = 0; j < V; ++j) shc[threadIdx.y * C + c0 + j] = acc[j];
    __syncthreads();
    for (int c = tid; c < C; c += nthr) {
        float t = 0.f;
        for (int yy = 0; yy < ty; ++yy) t += shc[yy * C + c];
        atomicAdd(db + c, t);
    }
}

extern "C" int vqb_act_bwd_bias(const void* y, const void* dy, void* dx, int dtype, int act, float alpha, float gain, int64_t P,
                                int C, float* db, void* stream) {
    VQB_CHECK_ARG(y && dy && dx && P > 0 && C > 0 && gain != 0.f, "act_bwd_bias: bad arguments");
    VQB_CHECK_ARG(act == VQB_ACT_NONE || act == VQB_ACT_TANH || act == VQB_ACT_LRELU || act == VQB_ACT_RELU,
                  "act_bwd_bias: activation must be recoverable from its output");
    const int vw = (dtype == VQB_BF16) ? 8 : 4;
    VQB_CHECK_ARG(C % vw == 0 && C / vw <= 256, "act_bwd_bias: C must be a multiple of the 16-byte vector width and <= 256 vectors");
    int tx = C / vw, ty = 256 / tx; if (ty < 1) ty = 1;
    int rows = (int)ceil_div64(P, (int64_t)kSMs * 16);
    if (rows < ty * 8) rows = ty * 8;
    int g = (int)ceil_div64(P, rows);
    dim3 block(tx, ty);
    static const int use_async = getenv("VQB_ACT_ASYNC") ? atoi(getenv("VQB_ACT_ASYNC")) : 1;
    if (use_async && dtype == VQB_BF16 && tx * ty == 256) {
        size_t ring = (size_t)ACT_RING_DEPTH * 2 * 256 * 16, red = sizeof(float) * ty * C;
        size_t smem = ring > red ? ring : red;
        static bool attr_set = false;
        if (!attr_set) { VQB_CUDA(cudaFuncSetAttribute(act_bwd_bias_async_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024)); attr_set = true; }
        act_bwd_bias_async_kernel<<<g, block, smem, as_stream(stream)>>>((const bf16*)y, (const bf16*)dy, (bf16*)dx, act, alpha, gain, P, C, rows, db);
        VQB_CHECK_LAUNCH("act_bwd_bias_async");
        return VQB_OK;
    }
    size_t sm = db ? sizeof(float) * ty * C : 0;
    VQB_DISPATCH_1(dtype, T, (act_bwd_bias_kernel<T><<<g, block, sm, as_stream(stream)>>>((const T*)y, (const T*)dy, (T*)dx, act, alpha,
                                                                                         gain, P, C, rows, db));)
    VQB_CHECK_LAUNCH("act_bwd_bias");
    return VQB_OK;
}

// ---------------------------------------------------------------------------------------------------
// fused AdamW over a flat range (torch.optim.AdamW semantics)
// ---------------------------------------------------------------------------------------------------
__global__ void adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                             float* __restrict__ v, int64_t n, float lr, float beta1, float beta2, float eps,
                             float wd, float bc1, float bc2_sqrt, float grad_scale) {
    const float step_size = lr / bc1;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float gi = g[i] * grad_scale;
        float pi = p[i] * (1.0f - lr * wd);
        float mi = m[i] * beta1 + (1.0f - beta1) * gi;
        float vi = v[i] * beta2 + (1.0f - beta2) * gi * gi;
        float denom = sqrtf(vi) / bc2_sqrt + eps;
        p[i] = pi - step_size * (mi / denom);
        m[i] = mi;
        v[i] = vi;
    }
}

// the same update with the per-step scalars read from DEVICE memory (hyper = {lr, bc1, sqrt(bc2)}), so that the launch can be
// captured in a CUDA graph and replayed while the host only rewrites three floats in a pinned buffer (lr schedule, step count)
__global__ void adamw_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                 int64_t n, const float* __restrict__ hyper, float beta1, float beta2, float eps, float wd,
                                 float grad_scale) {
    const float lr = hyper[0], bc2_sqrt = hyper[2];
    const float step_size = lr / hyper[1];
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float gi = g[i] * grad_scale;
        float pi = p[i] * (1.0f - lr * wd);
        float mi = m[i] * beta1 + (1.0f - beta1) * gi;
        float vi = v[i] * beta2 + (1.0f - beta2) * gi * gi;
        float denom = sqrtf(vi) / bc2_sqrt + eps;
        p[i] = pi - step_size * (mi / denom);
        m[i] = mi;
        v[i] = vi;
    }
}

extern "C" int vqb_adamw_dev(float* p, const float* g, float* m, float* v, int64_t n, const float* hyper, float beta1, float beta2,
                             float eps, float weight_decay, float grad_scale, void* stream) {
    VQB_CHECK_ARG(p && g && m && v && hyper && n >= 0, "adamw_dev: bad arguments");
    if (n == 0) return VQB_OK;
    int gsz = grid_for(n, 256);
    adamw_dev_kernel<<<gsz, 256, 0, as_stream(stream)>>>(p, g, m, v, n, hyper, beta1, beta2, eps, weight_decay, grad_scale);
    VQB_CHECK_LAUNCH("adamw_dev");
    return VQB_OK;
}

extern "C" int vqb_adamw(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                         float eps, float weight_decay, int step, float grad_scale, void* stream) {
    VQB_CHECK_ARG(p && g && m && v && n >= 0 && step >= 1, "adamw: bad arguments");
    if (n == 0) return VQB_OK;
    double bc1 = 1.0 - pow((double)beta1, (double)step);
    double bc2 = 1.0 - pow((double)beta2, (double)step);
    int gsz = grid_for(n, 256);
    adamw_kernel<<<gsz, 256, 0, as_stream(stream)>>>(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, (float)bc1,
                                                      (float)sqrt(bc2), grad_scale);
    VQB_CHECK_LAUNCH("adamw");
    return VQB_OK;
}
