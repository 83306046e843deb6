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

// Separable, register-tiled form: one thread produces a 2 x 2 patch of outputs for V channels.  The DOWN + 4 input rows the
// patch needs are streamed one at a time: DOWN + 4 vector loads per row, two horizontal 4-tap sums, then the vertical taps
// accumulate them into the two output rows -- (DOWN+4)^2 loads per 4 outputs (6.25 / 9 per output instead of 16) and one
// pass of FMAs less.  y[o] = sum_a f[a] x[o*DOWN - pad + a] per axis, zero outside the input; OH / OW are free parameters, so
// the adjoint of the down = 1 filter is this kernel too (symmetric taps): dx = FIR(dy, pad' = 3 - pad) with OH = H.
template <typename T, int DOWN>
__global__ void fir4_patch_kernel(const T* __restrict__ x, T* __restrict__ y, int N, int H, int W, int C, int OH, int OW, int pad) {
    constexpr int V = Vec<T>::N, R = DOWN + 4;
    const int Cv = C / V, PH = (OH + 1) / 2, PW = (OW + 1) / 2;
    const int64_t total = (int64_t)N * PH * PW * Cv;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int cv = (int)(i % Cv); int64_t r_ = i / Cv; int pw = (int)(r_ % PW); r_ /= PW; int ph = (int)(r_ % PH); int n = (int)(r_ / PH);
        const int oh0 = 2 * ph, ow0 = 2 * pw;
        const int ih0 = oh0 * DOWN - pad, iw0 = ow0 * DOWN - pad;
        float acc[2][2][V];
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 2; ++b)
#pragma unroll
                for (int u = 0; u < V; ++u) acc[a][b][u] = 0.f;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int ih = ih0 + r;
            if (ih < 0 || ih >= H) continue;
            float v[R][V];
#pragma unroll
            for (int c = 0; c < R; ++c) {
                const int iw = iw0 + c;
                if (iw >= 0 && iw < W) ldvec<T>(x + (((int64_t)n * H + ih) * W + iw) * C + cv * V, v[c]);
                else {
#pragma unroll
                    for (int u = 0; u < V; ++u) v[c][u] = 0.f;
                }
            }
            float h0[V], h1[V];
#pragma unroll
            for (int u = 0; u < V; ++u) {
                h0[u] = fmaf(0.375f, v[1][u] + v[2][u], 0.125f * (v[0][u] + v[3][u]));
                h1[u] = fmaf(0.375f, v[DOWN + 1][u] + v[DOWN + 2][u], 0.125f * (v[DOWN][u] + v[DOWN + 3][u]));
            }
            if (r <= 3) {                                         // output row oh0: vertical tap a = r
                const float f = tap(r);
#pragma unroll
                for (int u = 0; u < V; ++u) { acc[0][0][u] = fmaf(f, h0[u], acc[0][0][u]); acc[0][1][u] = fmaf(f, h1[u], acc[0][1][u]); }
            }
            if (r >= DOWN) {                                      // output row oh0 + 1: vertical tap a = r - DOWN
                const float f = tap(r - DOWN);
#pragma unroll
                for (int u = 0; u < V; ++u) { acc[1][0][u] = fmaf(f, h0[u], acc[1][0][u]); acc[1][1][u] = fmaf(f, h1[u], acc[1][1][u]); }
            }
        }
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 2; ++b)
                if (oh0 + a < OH && ow0 + b < OW) stvec<T>(y + (((int64_t)n * OH + oh0 + a) * OW + ow0 + b) * C + cv * V, acc[a][b]);
    }
}

template <typename T>
int launch_fir4_patch(const void* x, void* y, int N, int H, int W, int C, int OH, int OW, int pad, int down, cudaStream_t st) {
    const int64_t total = (int64_t)N * ((OH + 1) / 2) * ((OW + 1) / 2) * (C / Vec<T>::N);
    int64_t blocks = (total + 255) / 256; if (blocks > 148 * 32) blocks = 148 * 32; if (blocks < 1) blocks = 1;
    const unsigned grid = (unsigned)blocks;
    if (down == 1) fir4_patch_kernel<T, 1><<<grid, 256, 0, st>>>((const T*)x, (T*)y, N, H, W, C, OH, OW, pad);
    else fir4_patch_kernel<T, 2><<<grid, 256, 0, st>>>((const T*)x, (T*)y, N, H, W, C, OH, OW, pad);
    return 0;
}


// Sliding-window form of the down = 1 filter: one thread owns V channels x CW = 4 output columns and walks DOWN a strip of RH
// output rows.  Per input row it needs CW + 3 = 7 vectors (1.75 16-byte loads per output instead of the patch kernel's 6.25, which
// ncu had LSU-bound at ~22 % of the HBM roof); the 4 horizontal sums feed the four output rows in flight (a ring of accumulators
// with static indices: the row loop is unrolled by 4).  The 128 accumulator registers leave ~8 warps per SM, too few to cover
// HBM latency with plain loads (first version: 1 TB/s) -- so every thread streams ITS OWN input rows through a private
// DEPTH-deep ring in shared memory with cp.async (no registers held in flight, no inter-thread synchronisation), the scheme of the
// GroupNorm kernels.
// Layout flags: IN_S2D / OUT_S2D address the tensor as its 2x2 space-to-depth form [N][ceil(H/2)][ceil(W/2)][(dy, dx, c)]
// (physical dims given by the *_h2 / *_w2 arguments) -- the discriminator's stride-2 3x3 convolution then runs as a 2x2-tap
// stride-1 convolution over 4C channels of the filtered tensor (vqb_conv2d_fwd_sub), with no decimation pass.  OUT_S2D also
// zero-fills the padding row / column of the physical tensor (logical index OH / OW when they are odd).
constexpr int FIR_DEPTH = 4;          // input rows in flight per thread
constexpr int FIR_THREADS = 128;

__device__ __forceinline__ void fir_cp_async16(void* smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void fir_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void fir_cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <typename T>
__device__ __forceinline__ void unpack_vec(const uint4& u, float* v) {
    if constexpr (sizeof(T) == 2) {
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
        for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
    } else {
        v[0] = __uint_as_float(u.x); v[1] = __uint_as_float(u.y); v[2] = __uint_as_float(u.z); v[3] = __uint_as_float(u.w);
    }
}

template <typename T, bool IN_S2D, bool OUT_S2D>
__global__ void __launch_bounds__(FIR_THREADS) fir4_strip_kernel(const T* __restrict__ x, T* __restrict__ y, int N, int H, int W, int C, int OH, int OW,
                                                                 int pad, int RH, int in_h2, int in_w2, int out_h2, int out_w2) {
    constexpr int V = Vec<T>::N, CW = 4, NC = CW + 3, D = FIR_DEPTH;
    extern __shared__ uint4 fir_ring_raw[];                         // 56 KB: slot [row mod D][column][thread]
    uint4 (*ring)[NC][FIR_THREADS] = reinterpret_cast<uint4 (*)[NC][FIR_THREADS]>(fir_ring_raw);
    const int tid = threadIdx.x;
    const int Cv = C / V;
    const int OHp = OUT_S2D ? 2 * out_h2 : OH, OWp = OUT_S2D ? 2 * out_w2 : OW;     // rows / columns that must be WRITTEN
    const int nstrips = (OHp + RH - 1) / RH, ncb = (OWp + CW - 1) / CW;
    const int64_t total = (int64_t)N * nstrips * ncb * Cv;
    auto in_off = [&](int n, int ih, int iw) -> int64_t {
        if constexpr (IN_S2D) return ((((int64_t)n * in_h2 + (ih >> 1)) * in_w2 + (iw >> 1)) * 4 + ((ih & 1) * 2 + (iw & 1))) * C;
        else return (((int64_t)n * H + ih) * W + iw) * C;
    };
    auto out_off = [&](int n, int oh, int ow) -> int64_t {
        if constexpr (OUT_S2D) return ((((int64_t)n * out_h2 + (oh >> 1)) * out_w2 + (ow >> 1)) * 4 + ((oh & 1) * 2 + (ow & 1))) * C;
        else return (((int64_t)n * OH + oh) * OW + ow) * C;
    };
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + tid; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int cv = (int)(i % Cv); int64_t r_ = i / Cv;
        const int cb = (int)(r_ % ncb); r_ /= ncb;
        const int strip = (int)(r_ % nstrips); const int n = (int)(r_ / nstrips);
        const int oh0 = strip * RH, ow0 = cb * CW;
        int rows = OHp - oh0; if (rows > RH) rows = RH;
        const int nin = rows + 3;                                  // input rows oh0 - pad .. oh0 - pad + rows + 2
        // request input row r into ring slot r mod D (one commit group per row, empty when the row lies outside the image)
        auto request = [&](int r) {
            const int ih = oh0 - pad + r;
            if (r < nin && ih >= 0 && ih < H) {
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    const int iw = ow0 - pad + c;
                    if (iw >= 0 && iw < W) fir_cp_async16(&ring[r % D][c][tid], x + in_off(n, ih, iw) + cv * V);
                    else ring[r % D][c][tid] = make_uint4(0u, 0u, 0u, 0u);
                }
            }
            fir_cp_commit();
        };
#pragma unroll
        for (int r = 0; r < D; ++r) request(r);
        float acc[4][CW][V];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int j = 0; j < CW; ++j)
#pragma unroll
                for (int u = 0; u < V; ++u) acc[a][j][u] = 0.f;
        for (int r4 = 0; r4 < nin; r4 += 4) {
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) {
                const int r = r4 + rr;
                if (r < nin) {
                    fir_cp_wait<D - 1>();                           // the group of row r has landed (rows r+1 .. r+D-1 may be in flight)
                    const int ih = oh0 - pad + r;
                    if (ih >= 0 && ih < H) {
                        // horizontal sums with a 4-wide register window over the 7 columns
                        float w0[V], w1[V], w2[V], w3[V];
                        unpack_vec<T>(ring[r % D][0][tid], w0); unpack_vec<T>(ring[r % D][1][tid], w1); unpack_vec<T>(ring[r % D][2][tid], w2);
#pragma unroll
                        for (int j = 0; j < CW; ++j) {
                            unpack_vec<T>(ring[r % D][j + 3][tid], w3);
#pragma unroll
                            for (int u = 0; u < V; ++u) {
                                const float hs = fmaf(0.375f, w1[u] + w2[u], 0.125f * (w0[u] + w3[u]));
                                // vertical tap a of this input row belongs to output row r - a (ring slot (rr - a) & 3)
                                acc[(rr + 4 - 0) & 3][j][u] = fmaf(0.125f, hs, acc[(rr + 4 - 0) & 3][j][u]);
                                acc[(rr + 4 - 1) & 3][j][u] = fmaf(0.375f, hs, acc[(rr + 4 - 1) & 3][j][u]);
                                acc[(rr + 4 - 2) & 3][j][u] = fmaf(0.375f, hs, acc[(rr + 4 - 2) & 3][j][u]);
                                acc[(rr + 4 - 3) & 3][j][u] = fmaf(0.125f, hs, acc[(rr + 4 - 3) & 3][j][u]);
                                w0[u] = w1[u]; w1[u] = w2[u]; w2[u] = w3[u];
                            }
                        }
                    }
                    request(r + D);                                 // refill the slot just consumed (D == 4: slot rr)
                    {                                               // output row r - 3 has received its last tap (rows < 0: only the reset)
                        const int oh = oh0 + r - 3;
#pragma unroll
                        for (int j = 0; j < CW; ++j) {
                            const int ow = ow0 + j;
                            if (r >= 3 && ow < OWp) {
                                if (OUT_S2D && (oh >= OH || ow >= OW)) {
#pragma unroll
                                    for (int u = 0; u < V; ++u) acc[(rr + 1) & 3][j][u] = 0.f;
                                }
                                stvec<T>(y + out_off(n, oh, ow) + cv * V, acc[(rr + 1) & 3][j]);
                            }
#pragma unroll
                            for (int u = 0; u < V; ++u) acc[(rr + 1) & 3][j][u] = 0.f;
                        }
                    }
                }
            }
        }
        fir_cp_wait<0>();                                           // nothing of this item is in flight when the ring is reused
    }
}

template <typename T>
int launch_fir4_strip(const void* x, void* y, int N, int H, int W, int C, int OH, int OW, int pad, int in_s2d, int out_s2d, cudaStream_t st) {
    const int in_h2 = (H + 1) / 2, in_w2 = (W + 1) / 2, out_h2 = (OH + 1) / 2, out_w2 = (OW + 1) / 2;
    const int OHp = out_s2d ? 2 * out_h2 : OH, OWp = out_s2d ? 2 * out_w2 : OW;
    // strip height: enough threads to fill the machine (>= ~8 warps per SM scheduler), but long strips amortise the 3-row prologue
    const int64_t per_row_strip = (int64_t)N * ((OWp + 3) / 4) * (C / Vec<T>::N);
    int RH = 32;
    while (RH > 8 && per_row_strip * ((OHp + RH - 1) / RH) < (int64_t)148 * 2048) RH >>= 1;
    const int64_t total = per_row_strip * ((OHp + RH - 1) / RH);
    int64_t blocks = (total + FIR_THREADS - 1) / FIR_THREADS; if (blocks > 148 * 64) blocks = 148 * 64; if (blocks < 1) blocks = 1;
    const unsigned grid = (unsigned)blocks;
    if (in_s2d && out_s2d) return -1;
    const size_t smem = (size_t)FIR_DEPTH * 7 * FIR_THREADS * sizeof(uint4);
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(fir4_strip_kernel<T, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(fir4_strip_kernel<T, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(fir4_strip_kernel<T, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr_set = true;
    }
    if (in_s2d) fir4_strip_kernel<T, true, false><<<grid, FIR_THREADS, smem, st>>>((const T*)x, (T*)y, N, H, W, C, OH, OW, pad, RH, in_h2, in_w2, out_h2, out_w2);
    else if (out_s2d) fir4_strip_kernel<T, false, true><<<grid, FIR_THREADS, smem, st>>>((const T*)x, (T*)y, N, H, W, C, OH, OW, pad, RH, in_h2, in_w2, out_h2, out_w2);
    else fir4_strip_kernel<T, false, false><<<grid, FIR_THREADS, smem, st>>>((const T*)x, (T*)y, N, H, W, C, OH, OW, pad, RH, in_h2, in_w2, out_h2, out_w2);
    return 0;
}


// ---- down = 2 (the discriminator's skip path: upfirdn2d(x, f, down=2, padding=1)), the same strip / cp.async-ring scheme ---------
// Forward: y[oh][ow] = sum_{a,b} f[a] f[b] x[2 oh - pad + a][2 ow - pad + b].  One thread = V channels x 2 output columns (6 input
// columns) walking down a strip; input row r feeds output row r >> 1 (tap r & 1) and the one before it (tap (r & 1) + 2): two
// output rows in flight.
template <typename T>
__global__ void __launch_bounds__(FIR_THREADS) fir4_down2_strip_kernel(const T* __restrict__ x, T* __restrict__ y, int N, int H, int W, int C,
                                                                       int OH, int OW, int pad, int RH) {
    constexpr int V = Vec<T>::N, CW = 2, NC = 2 * CW + 2, D = FIR_DEPTH;
    extern __shared__ uint4 fir_ring_raw[];
    uint4 (*ring)[NC][FIR_THREADS] = reinterpret_cast<uint4 (*)[NC][FIR_THREADS]>(fir_ring_raw);
    const int tid = threadIdx.x;
    const int Cv = C / V;
    const int nstrips = (OH + RH - 1) / RH, ncb = (OW + CW - 1) / CW;
    const int64_t total = (int64_t)N * nstrips * ncb * Cv;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + tid; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int cv = (int)(i % Cv); int64_t r_ = i / Cv;
        const int cb = (int)(r_ % ncb); r_ /= ncb;
        const int strip = (int)(r_ % nstrips); const int n = (int)(r_ / nstrips);
        const int oh0 = strip * RH, ow0 = cb * CW;
        int rows = OH - oh0; if (rows > RH) rows = RH;
        const int nin = 2 * rows + 2;                               // input rows 2 oh0 - pad .. + 2 rows + 1
        auto request = [&](int r) {
            const int ih = 2 * oh0 - pad + r;
            if (r < nin && ih >= 0 && ih < H) {
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    const int iw = 2 * ow0 - pad + c;
                    if (iw >= 0 && iw < W) fir_cp_async16(&ring[r % D][c][tid], x + (((int64_t)n * H + ih) * W + iw) * C + cv * V);
                    else ring[r % D][c][tid] = make_uint4(0u, 0u, 0u, 0u);
                }
            }
            fir_cp_commit();
        };
#pragma unroll
        for (int r = 0; r < D; ++r) request(r);
        float acc[2][CW][V];
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int j = 0; j < CW; ++j)
#pragma unroll
                for (int u = 0; u < V; ++u) acc[a][j][u] = 0.f;
        for (int r4 = 0; r4 < nin; r4 += 4) {
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) {
                const int r = r4 + rr;
                if (r < nin) {
                    fir_cp_wait<D - 1>();
                    const int ih = 2 * oh0 - pad + r;
                    // taps: output (r >> 1) gets f[r & 1] (slot (rr >> 1) & 1), output (r >> 1) - 1 gets f[(r & 1) + 2] (the other slot)
                    constexpr float F[4] = {0.125f, 0.375f, 0.375f, 0.125f};
                    const float fa = F[rr & 1], fb = F[(rr & 1) + 2];
                    if (ih >= 0 && ih < H) {
                        float v[NC][V];
#pragma unroll
                        for (int c = 0; c < NC; ++c) unpack_vec<T>(ring[r % D][c][tid], v[c]);
#pragma unroll
                        for (int j = 0; j < CW; ++j)
#pragma unroll
                            for (int u = 0; u < V; ++u) {
                                const float hs = fmaf(0.375f, v[2 * j + 1][u] + v[2 * j + 2][u], 0.125f * (v[2 * j][u] + v[2 * j + 3][u]));
                                acc[(rr >> 1) & 1][j][u] = fmaf(fa, hs, acc[(rr >> 1) & 1][j][u]);
                                acc[((rr >> 1) + 1) & 1][j][u] = fmaf(fb, hs, acc[((rr >> 1) + 1) & 1][j][u]);
                            }
                    }
                    request(r + D);
                    if (rr & 1) {                                   // odd input row r = 2 o + 3 completes output o = (r - 3) / 2 (other slot)
                        const int oh = oh0 + ((r - 3) >> 1);
#pragma unroll
                        for (int j = 0; j < CW; ++j) {
                            const int ow = ow0 + j;
                            if (r >= 3 && ow < OW) stvec<T>(y + (((int64_t)n * OH + oh) * OW + ow) * C + cv * V, acc[((rr >> 1) + 1) & 1][j]);
#pragma unroll
                            for (int u = 0; u < V; ++u) acc[((rr >> 1) + 1) & 1][j][u] = 0.f;
                        }
                    }
                }
            }
        }
        fir_cp_wait<0>();
    }
}

// Adjoint for pad = 1: dx[h][w] = sum over the taps of matching parity of f[a] f[b] dy[(h + 1 - a) / 2][(w + 1 - b) / 2].  Per axis:
//   D[2k] = f1 dy[k] + f3 dy[k-1],   D[2k+1] = f0 dy[k+1] + f2 dy[k].
// One thread = V channels x 4 output columns 4q .. 4q+3 (dy columns 2q-1 .. 2q+2) walking down the dy rows of a strip: dy row m
// feeds output rows 2m-1 (f0), 2m (f1), 2m+1 (f2), 2m+2 (f3) -- four output rows in flight -- and completes 2m-1 and 2m.
template <typename T>
__global__ void __launch_bounds__(FIR_THREADS) fir4_up2_adjoint_strip_kernel(const T* __restrict__ dy, T* __restrict__ dx, int N, int H, int W, int C,
                                                                             int OH, int OW, int RM) {
    constexpr int V = Vec<T>::N, NC = 4, D = FIR_DEPTH;
    extern __shared__ uint4 fir_ring_raw[];
    uint4 (*ring)[NC][FIR_THREADS] = reinterpret_cast<uint4 (*)[NC][FIR_THREADS]>(fir_ring_raw);
    const int tid = threadIdx.x;
    const int Cv = C / V;
    const int MH = (H + 1) / 2;                                      // dy rows m = 0 .. MH cover the output rows (row m completes 2m-1, 2m)
    const int nstrips = (MH + RM) / RM, ncb = (W + 3) / 4;           // strips of RM input rows over m in [0, MH]
    const int64_t total = (int64_t)N * nstrips * ncb * Cv;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + tid; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int cv = (int)(i % Cv); int64_t r_ = i / Cv;
        const int cb = (int)(r_ % ncb); r_ /= ncb;
        const int strip = (int)(r_ % nstrips); const int n = (int)(r_ / nstrips);
        const int m_lo = strip * RM;                                 // this strip OWNS output rows 2 m_lo - 1 .. 2 (m_lo + cnt - 1)
        int cnt = MH + 1 - m_lo; if (cnt > RM) cnt = RM;
        const int w0 = cb * 4, q2 = cb * 2;
        // input rows m_lo - 1 .. m_lo + cnt - 1 (the first one only contributes its f2 / f3 taps to the first owned rows)
        const int nin = cnt + 1;
        auto request = [&](int r) {
            const int m = m_lo - 1 + r;
            if (r < nin && m >= 0 && m < OH) {
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    const int k = q2 - 1 + c;
                    if (k >= 0 && k < OW) fir_cp_async16(&ring[r % D][c][tid], dy + (((int64_t)n * OH + m) * OW + k) * C + cv * V);
                    else ring[r % D][c][tid] = make_uint4(0u, 0u, 0u, 0u);
                }
            }
            fir_cp_commit();
        };
#pragma unroll
        for (int r = 0; r < D; ++r) request(r);
        // slot s holds output row h with (h - (2 m_lo - 1)) & 3 == s ; relative row t = h - (2 m_lo - 1)
        float acc[4][4][V];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int u = 0; u < V; ++u) acc[a][j][u] = 0.f;
        for (int r2 = 0; r2 < nin; r2 += 2) {
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
                const int r = r2 + rr;
                if (r < nin) {
                    fir_cp_wait<D - 1>();
                    const int m = m_lo - 1 + r;
                    // relative output rows fed by input row r: 2r-2 (f0), 2r-1 (f1), 2r (f2), 2r+1 (f3); with r = r2 + rr and r2 even the
                    // slots are static: (2rr-2)&3, (2rr-1)&3, (2rr)&3, (2rr+1)&3
                    if (m >= 0 && m < OH) {
                        float v[NC][V];
#pragma unroll
                        for (int c = 0; c < NC; ++c) unpack_vec<T>(ring[r % D][c][tid], v[c]);
#pragma unroll
                        for (int u = 0; u < V; ++u) {
                            float hsv[4];
                            hsv[0] = fmaf(0.375f, v[1][u], 0.125f * v[0][u]);        // D[4q]   = f1 dy[2q]   + f3 dy[2q-1]
                            hsv[1] = fmaf(0.125f, v[2][u], 0.375f * v[1][u]);        // D[4q+1] = f0 dy[2q+1] + f2 dy[2q]
                            hsv[2] = fmaf(0.375f, v[2][u], 0.125f * v[1][u]);        // D[4q+2] = f1 dy[2q+1] + f3 dy[2q]
                            hsv[3] = fmaf(0.125f, v[3][u], 0.375f * v[2][u]);        // D[4q+3] = f0 dy[2q+2] + f2 dy[2q+1]
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                acc[(2 * rr + 2) & 3][j][u] = fmaf(0.125f, hsv[j], acc[(2 * rr + 2) & 3][j][u]);   // row 2r-2: f0
                                acc[(2 * rr + 3) & 3][j][u] = fmaf(0.375f, hsv[j], acc[(2 * rr + 3) & 3][j][u]);   // row 2r-1: f1
                                acc[(2 * rr + 0) & 3][j][u] = fmaf(0.375f, hsv[j], acc[(2 * rr + 0) & 3][j][u]);   // row 2r  : f2
                                acc[(2 * rr + 1) & 3][j][u] = fmaf(0.125f, hsv[j], acc[(2 * rr + 1) & 3][j][u]);   // row 2r+1: f3
                            }
                        }
                    }
                    request(r + D);
                    // rows 2r-2 and 2r-1 (relative) are complete; absolute h = 2 m_lo - 1 + t.  Owned rows: t in [0, 2 cnt - 1]
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int t = 2 * r - 2 + e;
                        const int h = 2 * m_lo - 1 + t;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int w = w0 + j;
                            if (t >= 0 && h >= 0 && h < H && w < W) stvec<T>(dx + (((int64_t)n * H + h) * W + w) * C + cv * V, acc[(2 * rr + 2 + e) & 3][j]);
#pragma unroll
                            for (int u = 0; u < V; ++u) acc[(2 * rr + 2 + e) & 3][j][u] = 0.f;
                        }
                    }
                }
            }
        }
        fir_cp_wait<0>();
    }
}

template <typename T>
int launch_fir4_down2_strip(const void* x, void* y, int N, int H, int W, int C, int OH, int OW, int pad, cudaStream_t st) {
    const int64_t per_row_strip = (int64_t)N * ((OW + 1) / 2) * (C / Vec<T>::N);
    int RH = 16;
    while (RH > 4 && per_row_strip * ((OH + RH - 1) / RH) < (int64_t)148 * 2048) RH >>= 1;
    const int64_t total = per_row_strip * ((OH + RH - 1) / RH);
    int64_t blocks = (total + FIR_THREADS - 1) / FIR_THREADS; if (blocks > 148 * 64) blocks = 148 * 64; if (blocks < 1) blocks = 1;
    const size_t smem = (size_t)FIR_DEPTH * 6 * FIR_THREADS * sizeof(uint4);
    static bool attr_set = false;
    if (!attr_set) { cudaFuncSetAttribute(fir4_down2_strip_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr_set = true; }
    fir4_down2_strip_kernel<T><<<(unsigned)blocks, FIR_THREADS, smem, st>>>((const T*)x, (T*)y, N, H, W, C, OH, OW, pad, RH);
    return 0;
}

template <typename T>
int launch_fir4_up2_adjoint_strip(const void* dy, void* dx, int N, int H, int W, int C, int OH, int OW, cudaStream_t st) {
    const int MH = (H + 1) / 2;
    const int64_t per_row_strip = (int64_t)N * ((W + 3) / 4) * (C / Vec<T>::N);
    int RM = 16;
    while (RM > 4 && per_row_strip * ((MH + RM) / RM) < (int64_t)148 * 2048) RM >>= 1;
    const int64_t total = per_row_strip * ((MH + RM) / RM);
    int64_t blocks = (total + FIR_THREADS - 1) / FIR_THREADS; if (blocks > 148 * 64) blocks = 148 * 64; if (blocks < 1) blocks = 1;
    const size_t smem = (size_t)FIR_DEPTH * 4 * FIR_THREADS * sizeof(uint4);
    fir4_up2_adjoint_strip_kernel<T><<<(unsigned)blocks, FIR_THREADS, smem, st>>>((const T*)dy, (T*)dx, N, H, W, C, OH, OW, RM);
    return 0;
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
    static const int use_strip = getenv("VQB_FIR_STRIP") ? atoi(getenv("VQB_FIR_STRIP")) : 1;
    if (use_strip && down == 1) {
        if (dtype == VQB_BF16) launch_fir4_strip<bf16>(x, y, N, H, W, C, OH, OW, pad, 0, 0, st);
        else launch_fir4_strip<float>(x, y, N, H, W, C, OH, OW, pad, 0, 0, st);
        VQB_CHECK_LAUNCH("fir4_strip");
        return VQB_OK;
    }
    if (use_strip && down == 2) {
        if (dtype == VQB_BF16) launch_fir4_down2_strip<bf16>(x, y, N, H, W, C, OH, OW, pad, st);
        else launch_fir4_down2_strip<float>(x, y, N, H, W, C, OH, OW, pad, st);
        VQB_CHECK_LAUNCH("fir4_down2_strip");
        return VQB_OK;
    }
    static const int use_patch = getenv("VQB_FIR_PATCH") ? atoi(getenv("VQB_FIR_PATCH")) : 1;
    if (use_patch && (down == 1 || down == 2)) {
        if (dtype == VQB_BF16) launch_fir4_patch<bf16>(x, y, N, H, W, C, OH, OW, pad, down, st);
        else launch_fir4_patch<float>(x, y, N, H, W, C, OH, OW, pad, down, st);
        VQB_CHECK_LAUNCH("fir4_patch");
        return VQB_OK;
    }
    if (dtype == VQB_BF16) fir4_fwd_vec_kernel<bf16><<<ew_grid((int64_t)N * OH * OW * C / 8), 256, 0, st>>>((const bf16*)x, (bf16*)y, N, H, W, C, OH, OW, pad, down);
    else fir4_fwd_vec_kernel<float><<<ew_grid((int64_t)N * OH * OW * C / 4), 256, 0, st>>>((const float*)x, (float*)y, N, H, W, C, OH, OW, pad, down);
    VQB_CHECK_LAUNCH("fir4_fwd_vec");
    return VQB_OK;
}
int vqb_fir4_bwd_vec(const void* dy, void* dx, int dtype, int N, int H, int W, int C, int OH, int OW, int pad, int down, cudaStream_t st) {
    static const int use_strip = getenv("VQB_FIR_STRIP") ? atoi(getenv("VQB_FIR_STRIP")) : 1;
    if (use_strip && down == 1 && pad <= 3) {
        // adjoint of the symmetric down = 1 filter = the same filter with pad' = 3 - pad, applied to dy [OH, OW] -> dx [H, W]
        if (dtype == VQB_BF16) launch_fir4_strip<bf16>(dy, dx, N, OH, OW, C, H, W, 3 - pad, 0, 0, st);
        else launch_fir4_strip<float>(dy, dx, N, OH, OW, C, H, W, 3 - pad, 0, 0, st);
        VQB_CHECK_LAUNCH("fir4_strip(adjoint)");
        return VQB_OK;
    }
    if (use_strip && down == 2 && pad == 1) {
        if (dtype == VQB_BF16) launch_fir4_up2_adjoint_strip<bf16>(dy, dx, N, H, W, C, OH, OW, st);
        else launch_fir4_up2_adjoint_strip<float>(dy, dx, N, H, W, C, OH, OW, st);
        VQB_CHECK_LAUNCH("fir4_up2_adjoint_strip");
        return VQB_OK;
    }
    static const int use_patch = getenv("VQB_FIR_PATCH") ? atoi(getenv("VQB_FIR_PATCH")) : 1;
    if (use_patch && down == 1 && pad <= 3) {
        // adjoint of the symmetric down = 1 filter = the same filter with pad' = 3 - pad, applied to dy [OH, OW] -> dx [H, W]
        if (dtype == VQB_BF16) launch_fir4_patch<bf16>(dy, dx, N, OH, OW, C, H, W, 3 - pad, 1, st);
        else launch_fir4_patch<float>(dy, dx, N, OH, OW, C, H, W, 3 - pad, 1, st);
        VQB_CHECK_LAUNCH("fir4_patch(adjoint)");
        return VQB_OK;
    }
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

// FIR [1,3,3,1]^2/64 (down = 1) between a plain NHWC tensor and a 2x2 space-to-depth tensor (see fir4_strip_kernel):
//   out_s2d: y = s2d(FIR(x, pad)),   x [N,H,W,C] -> y physical [N, ceil(OH/2), ceil(OW/2), 4C], OH = H + 2 pad - 3 (padding slots zeroed)
//   in_s2d : y = FIR(unS2D(x), pad), x physical [N, ceil(H/2), ceil(W/2), 4C] holding a logical [N,H,W,C] -> y [N,OH,OW,C] (OH, OW given:
//            the adjoint of the first form is this one with pad' = 3 - pad and OH = the original input height)
extern "C" int vqb_fir4_s2d(const void* x, void* y, int dtype, int N, int H, int W, int C, int OH, int OW, int pad, int in_s2d, int out_s2d,
                            void* stream) {
    VQB_CHECK_ARG(x && y && N > 0 && H > 0 && W > 0 && C > 0 && OH > 0 && OW > 0 && pad >= 0 && pad <= 3, "fir4_s2d: bad arguments");
    VQB_CHECK_ARG(!(in_s2d && out_s2d), "fir4_s2d: only one side may be in space-to-depth layout");
    VQB_CHECK_ARG(dtype == VQB_BF16 || dtype == VQB_F32, "fir4_s2d: dtype");
    VQB_CHECK_ARG(C % ((dtype == VQB_BF16) ? 8 : 4) == 0, "fir4_s2d: C must be a multiple of the 16-byte vector width");
    if (dtype == VQB_BF16) launch_fir4_strip<bf16>(x, y, N, H, W, C, OH, OW, pad, in_s2d, out_s2d, as_stream(stream));
    else launch_fir4_strip<float>(x, y, N, H, W, C, OH, OW, pad, in_s2d, out_s2d, as_stream(stream));
    VQB_CHECK_LAUNCH("fir4_s2d");
    return VQB_OK;
}
