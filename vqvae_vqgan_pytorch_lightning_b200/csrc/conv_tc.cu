// tcgen05 / TMA implicit-GEMM convolution for sm_100a (bf16 operands, fp32 accumulation in TMEM).
//
// Forward / dgrad, four kernels sharing one epilogue (dispatch in vqb_conv2d_fwd_tc):
//   conv_fwd_tc_kernel         generic (any KHxKW 'same' conv, any spatial size): per (tap, 64-channel chunk) one 4-D TMA box
//                              {64 ch, tw, th, nb} of the NHWC activation at offset (kh-pad, kw-pad) is a K-major
//                              SWIZZLE_128B A tile; out-of-image elements are zero-filled by TMA (= the conv padding).
//   conv_fwd_tc_halo_kernel    3x3 convs on >= 16x8 images: ONE halo box {64 ch, 10, 18} per channel chunk serves all nine
//                              taps -- the A operand of tap (kh,kw) is the same shared-memory tile addressed through a UMMA
//                              descriptor whose start is shifted by (kh*pitch + kw) rows and whose 8-row-group stride (SBO)
//                              is the halo pitch.  L2->SMEM traffic for A drops ~6x.  Now only the narrow heads (Co <= 16) and
//                              64-channel tiles take this kernel.
//   conv_fwd_tc_halo_t_kernel  Co tiles of 128: operand roles swapped (M = 128 co, N = 256 pixels), smem-transposed epilogue.
//   conv_fwd_tc_halo2_kernel   Co tiles of 256: CTA pairs (tcgen05 cta_group::2), M = 256 pixels, half the weight tile per SM.
//   All: B = packed weight [Co][(kh,kw,ci)] via 2-D TMA; persistent CTAs; warp 0 = TMA producer, warp 1 = TMEM allocator +
//   single-thread tcgen05.mma issuer, warps 2..9 = epilogue (tcgen05.ld -> bias / activation / prefetched residual -> NHWC
//   stores); fp32 accumulators double-buffered in TMEM so the epilogue of tile i overlaps the MMAs of tile i+1.
//
// Weight gradient (conv_wgrad_tc_kernel; conv_wgrad_tc_halo_kernel = 3 taps per CTA on one x halo, ONE wave of CTAs):
//   dW[tap][ci][co] = sum_pixels x_shift[pix][ci] * dy[pix][co]: the reduction runs over pixels, so both operands are
//   MN-major SWIZZLE_128B tiles (rows = pixels, 128 B = 64 channels), again straight from 4-D TMA boxes.  M = 128 co
//   (TMEM lanes), N = up to 256 ci (TMEM columns); the pixel range is split across CTAs and partial tiles are combined
//   with fp32 red.global.add (coalesced across lanes).
//
// Roofline: tensor pipe (bf16 dense); see DESIGN.md for the per-layer FLOP table.
#include "common.cuh"
#include "ptx.cuh"
#include <mutex>
#include <stdlib.h>

namespace {

constexpr int BM = 128;          // UMMA M (pixels for fwd, co for wgrad)
constexpr int BK = 64;           // K elements per pipeline stage (one 128-byte swizzle row of bf16)
constexpr int UMMA_K = 16;
constexpr int NTHREADS = 192;      // wgrad kernels: 6 warps (TMA, MMA, 4 x epilogue)
constexpr int NTHREADS_FWD = 320;  // forward kernels: 10 warps (TMA, MMA, 8 x epilogue -- two warps per TMEM lane quarter,
                                   // alternating 32-column chunks; the epilogue paces the 128-channel layers)
constexpr int SMEM_LIMIT = 227 * 1024;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

// NHWC bf16 activation as a 4-D tensor {C, W, H, N} with box {64, bw, bh, bn}
int make_act_map(CUtensorMap* m, const void* base, int N, int H, int W, int C, int bw, int bh, int bn) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { vqb_set_error("cuTensorMapEncodeTiled unavailable"); return VQB_ERR_CUDA; }
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bn};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { vqb_set_error("cuTensorMapEncodeTiled(activation) failed: %d", (int)r); return VQB_ERR_CUDA; }
    return VQB_OK;
}

// packed weight [rows][K] bf16 as a 2-D tensor {K, rows} with box {64, box_rows}
int make_weight_map(CUtensorMap* m, const void* base, int rows, int K, int box_rows) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { vqb_set_error("cuTensorMapEncodeTiled unavailable"); return VQB_ERR_CUDA; }
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { vqb_set_error("cuTensorMapEncodeTiled(weight) failed: %d", (int)r); return VQB_ERR_CUDA; }
    return VQB_OK;
}

inline int pow2_floor(int v) { int p = 1; while (p * 2 <= v) p *= 2; return p; }

// spatial tile of `pixels` (power of two) pixels: tw x th x nb
inline void pick_tile(int pixels, int H, int W, int& tw, int& th, int& nb) {
    tw = pow2_floor(W); if (tw > 16) tw = 16; if (tw > pixels) tw = pixels;
    th = pow2_floor(H); if (th > pixels / tw) th = pixels / tw;
    nb = pixels / (tw * th);
}

__device__ __forceinline__ float act_f(float v, int act, float alpha) {
    if (act == VQB_ACT_TANH) return tanhf(v);
    if (act == VQB_ACT_SILU) return silu_f(v);
    if (act == VQB_ACT_LRELU) return v > 0.f ? v : v * alpha;
    if (act == VQB_ACT_RELU) return fmaxf(v, 0.f);
    return v;
}

struct FwdParams {
    int N, H, W, Ci, Co, KH, KW, pad;
    int Cx;              // channels of the x TENSOR (== Ci, or 2*Ci/terms for split-precision operands: the k loop wraps over [hi | lo])
    int tw, th, nb, tiles_w, tiles_h, tiles_n, co_tiles, BN, stages, MT;
    int num_tiles, ksteps, cchunks;
    // halo kernel only
    int pitch, bo_mode, a_tile_bytes, a_stages, b_stages;
    const float* bias;
    const void* residual;
    void* y;
    int y_f32, act, narrow, res_prefetch;
    float alpha, gain;
    // GroupNorm statistics of the OUTPUT fused into the epilogue (the consumer is a GroupNorm with Co / gn_cpg groups): per (image,
    // group) sum and sum of squares of the final values, fp32 partials per tile, double atomics -- replaces the gn_stats pass
    double* gn_sums;     // [N][Co / gn_cpg][2] or NULL
    int gn_cpg;          // channels per group: 4, 8 or 16
    // halo kernels: the taps actually multiplied, as a sub-rectangle of the 3x3 frame the halo box serves -- tap t reads the frame
    // position (kh0 + t / ktw, kw0 + t % ktw) and the weight block t of the packed weight.  3x3: ntaps 9, ktw 3, kh0 = kw0 = 0.
    // A 2x2-tap convolution (the space-to-depth form of the discriminator's stride-2 3x3 convolution and its dgrad) uses 4.
    int ntaps, ktw, kh0, kw0;
    // one-CTA halo kernel: the WHOLE packed weight (ntaps x cchunks tiles of BN x 64) stays resident in shared memory -- loaded once
    // per CTA instead of once per pixel tile (narrow output heads: Co <= 16, one channel tile, 2 KB weight tiles whose per-tap TMA
    // round trips, not the tensor pipe, paced the kernel)
    int b_resident;
};

// ---- shared epilogue: 32 accumulator columns of one pixel row -> bias / act / residual -> NHWC store ---------------
// GroupNorm statistics in the pixel-major epilogue (thread = pixel, 32 channels per chunk): gs[2g], gs[2g+1] = (sum, sum of
// squares) of group g of the chunk for THIS thread's pixel.  The 32 lanes (pixels of one image) are reduced by xor butterflies
// -- every lane ends with every total -- and lane i keeps total i in `acc`, which lives across tiles: global double atomics are
// issued only when the CTA moves on to another image.  (One atomic per chunk and tile from every warp of every CTA hammered the
// 64 addresses of the image in flight: measured +30 % on the convolution, more than the statistics pass it replaced.)
__device__ __forceinline__ void gn_reduce_chunk(const FwdParams& p, float (&gs)[16], float& acc, int lane) {
    const int nv = 64 / p.gn_cpg;                        // 2 * groups per chunk: 16 / 8 / 4
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        if (i < nv) {
            float v = gs[i];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == i) acc += v;
        }
    }
}

template <int SHIFT>
__device__ __forceinline__ void gn_accumulate32_t(const float (&v)[32], float (&gs)[16]) {
#pragma unroll
    for (int j = 0; j < 32; ++j) {                       // static register indices: g = j >> SHIFT
        gs[2 * (j >> SHIFT)] += v[j];
        gs[2 * (j >> SHIFT) + 1] = fmaf(v[j], v[j], gs[2 * (j >> SHIFT) + 1]);
    }
}
__device__ __forceinline__ void gn_accumulate32(const FwdParams& p, const float (&v)[32], float (&gs)[16]) {
    // cpg = 4 / 8 / 16 channels per group -> 8 / 4 / 2 groups in a 32-channel chunk
    if (p.gn_cpg == 4) gn_accumulate32_t<2>(v, gs);
    else if (p.gn_cpg == 8) gn_accumulate32_t<3>(v, gs);
    else gn_accumulate32_t<4>(v, gs);
}

__device__ __forceinline__ void epilogue_chunk(const FwdParams& p, const uint32_t (&r)[32], bool valid, int64_t pix, int co0, int c,
                                               const uint4* qpre = nullptr, float* gsums = nullptr) {
    if (!valid) return;
    if (p.narrow) {
        // Co < BN (e.g. the 3-channel image head): scalar, masked stores; static register indices
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            if (co0 + c + j < p.Co) {
                float t = __uint_as_float(r[j]);
                if (p.bias) t += __ldg(p.bias + co0 + c + j);
                t = act_f(t, p.act, p.alpha) * p.gain;
                const int64_t off = pix * p.Co + co0 + c + j;
                if (p.y_f32) {
                    if (p.residual) t += reinterpret_cast<const float*>(p.residual)[off];
                    reinterpret_cast<float*>(p.y)[off] = t;
                } else {
                    if (p.residual) t += __bfloat162float(reinterpret_cast<const bf16*>(p.residual)[off]);
                    reinterpret_cast<bf16*>(p.y)[off] = __float2bfloat16_rn(t);
                }
            }
        }
        return;
    }
    float v[32];
    if (p.bias || p.act != VQB_ACT_NONE || p.gain != 1.0f) {
        // the 32 bias values of the chunk are the same for every pixel (thread): eight broadcast 16-byte loads, and the activation
        // is selected ONCE per chunk -- a per-element __ldg + act_f() chain made the bias + lrelu / relu layers of the loss heads
        // epilogue-bound (discriminator 128->256 @257^2: 422 TFLOP/s against 1.6 PFLOP/s for the same kernel without bias)
        float bb[32];
        if (p.bias) {
            const float4* bp = reinterpret_cast<const float4*>(p.bias + co0 + c);
#pragma unroll
            for (int j = 0; j < 8; ++j) { const float4 q = __ldg(bp + j); bb[4 * j] = q.x; bb[4 * j + 1] = q.y; bb[4 * j + 2] = q.z; bb[4 * j + 3] = q.w; }
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) bb[j] = 0.f;
        }
        const float gain = p.gain, alpha = p.alpha;
        if (p.act == VQB_ACT_LRELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) { const float t = __uint_as_float(r[j]) + bb[j]; v[j] = (t > 0.f ? t : t * alpha) * gain; }
        } else if (p.act == VQB_ACT_RELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(__uint_as_float(r[j]) + bb[j], 0.f) * gain;
        } else if (p.act == VQB_ACT_NONE) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = (__uint_as_float(r[j]) + bb[j]) * gain;
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = act_f(__uint_as_float(r[j]) + bb[j], p.act, alpha) * gain;
        }
    } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
    }
    const int64_t off = pix * p.Co + co0 + c;
    if (p.y_f32) {
        float* yo = reinterpret_cast<float*>(p.y) + off;
        const float* ro = p.residual ? reinterpret_cast<const float*>(p.residual) + off : nullptr;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
            if (ro) { float4 q = *reinterpret_cast<const float4*>(ro + j); v[j] += q.x; v[j + 1] += q.y; v[j + 2] += q.z; v[j + 3] += q.w; }
            *reinterpret_cast<float4*>(yo + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
    } else {
        bf16* yo = reinterpret_cast<bf16*>(p.y) + off;
        const bf16* ro = p.residual ? reinterpret_cast<const bf16*>(p.residual) + off : nullptr;
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
            if (ro) {
                // qpre: the residual of this chunk was requested before the accumulator was awaited (its latency is hidden)
                uint4 q = qpre ? qpre[j >> 3] : *reinterpret_cast<const uint4*>(ro + j);
                const __nv_bfloat162* qb = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
                for (int u = 0; u < 4; ++u) { float2 f = __bfloat1622float2(qb[u]); v[j + 2 * u] += f.x; v[j + 2 * u + 1] += f.y; }
            }
            uint4 o;
            __nv_bfloat162* ob = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
            for (int u = 0; u < 4; ++u) ob[u] = __floats2bfloat162_rn(v[j + 2 * u], v[j + 2 * u + 1]);
            *reinterpret_cast<uint4*>(yo + j) = o;
        }
    }
    if (gsums) gn_accumulate32(p, v, *reinterpret_cast<float (*)[16]>(gsums));
}

// epilogue warps of both forward kernels: drain the MT sub-tile accumulators of every tile this CTA owns.
// A fused bf16 residual is PREFETCHED one 32-column chunk ahead (the first chunk of a tile before the accumulator barrier is
// awaited): issued at the point of use, these 16-byte loads were latency-exposed and made residual convolutions 1.6x slower
// than plain ones (ncu: long-scoreboard stalls 9.6 vs 1.7 per issue, tensor pipe 43 % vs 91 % busy).
__device__ __forceinline__ void epilogue_loop(const FwdParams& p, uint32_t tmem_base, uint64_t* tfull, uint64_t* tempty, int warp, int lane) {
    const int quarter = warp & 3;                    // TMEM lane quarter this warp may read
    const int half = (warp - 2) >> 2;                // two warps per quarter take alternating 32-column chunks
    const int row = quarter * 32 + lane;
    const int wi = row % p.tw, r2 = row / p.tw, hi = r2 % p.th, ni = r2 / p.th;
    const bool pre = p.residual && !p.y_f32 && !p.narrow && p.res_prefetch;
    const int nch = (p.BN - half * 32 + 63) / 64;    // chunks of one sub-tile handled by this warp
    int as = 0; uint32_t aphase = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int ct = tile % p.co_tiles, pt = tile / p.co_tiles;
        const int co0 = ct * p.BN;
        auto locate = [&](int m, bool& valid, int64_t& pix) {
            const int q = pt * p.MT + m;
            const int twi = q % p.tiles_w, t2 = q / p.tiles_w, thi = t2 % p.tiles_h, tni = t2 / p.tiles_h;
            const int w = twi * p.tw + wi, h = thi * p.th + hi, n = tni * p.nb + ni;
            valid = (w < p.W) && (h < p.H) && (n < p.N);
            pix = ((int64_t)n * p.H + h) * p.W + w;
        };
        auto fetch = [&](int s, uint4 (&q)[4]) {       // s = m * nch + chunk index
            bool valid; int64_t pix;
            locate(s / nch, valid, pix);
            const int c = half * 32 + (s % nch) * 64;
            if (valid) {
                const uint4* ro = reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(p.residual) + pix * p.Co + co0 + c);
#pragma unroll
                for (int i = 0; i < 4; ++i) q[i] = ro[i];
            }
        };
        const int total = p.MT * nch;
        uint4 qn[4] = {};
        if (pre && total > 0) fetch(0, qn);
        ptx::mbar_wait(&tfull[as], aphase);
        ptx::tc_fence_after();
        for (int s = 0; s < total; ++s) {
            const int m = s / nch, c = half * 32 + (s % nch) * 64;
            bool valid; int64_t pix;
            locate(m, valid, pix);
            uint4 qc[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) qc[i] = qn[i];
            if (pre && s + 1 < total) fetch(s + 1, qn);
            const uint32_t t_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)((as * p.MT + m) * p.BN);
            uint32_t r[32];
            ptx::tmem_ld32(t_addr + (uint32_t)c, r);
            ptx::tmem_ld_wait();
            epilogue_chunk(p, r, valid, pix, co0, c, pre ? qc : nullptr);
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&tempty[as]);
        if (++as == 2) { as = 0; aphase ^= 1; }
    }
}

// ---------------------------------------------------------------------------------------------------
// generic forward kernel
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NTHREADS_FWD, 1)
conv_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const FwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int a_tile = BM * BK * 2;                  // 16 KB per 128-pixel sub-tile
    const int a_bytes = p.MT * a_tile;               // MT sub-tiles share one B tile (doubles FLOP per byte staged)
    const int b_bytes = p.BN * BK * 2;
    const int stage_bytes = a_bytes + b_bytes;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
    uint64_t* empty = full + p.stages;
    uint64_t* tfull = empty + p.stages;              // [2]
    uint64_t* tempty = tfull + 2;                    // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmA);
        ptx::prefetch_tmap(&tmB);
        for (int i = 0; i < p.stages; ++i) { ptx::mbar_init(&full[i], 1); ptx::mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { ptx::mbar_init(&tfull[i], 1); ptx::mbar_init(&tempty[i], 8); }
        ptx::fence_barrier_init();
    }
    if (warp == 1) ptx::tmem_alloc(tmem_slot, 512);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                int ct = tile % p.co_tiles; int pt = tile / p.co_tiles;
                int w0[2], h0[2], n0[2];
                for (int m = 0; m < p.MT; ++m) {          // sub-tiles past the end decode to n0 >= N: TMA zero-fills them
                    int q = pt * p.MT + m;
                    int twi = q % p.tiles_w; int t2 = q / p.tiles_w; int thi = t2 % p.tiles_h; int tni = t2 / p.tiles_h;
                    w0[m] = twi * p.tw; h0[m] = thi * p.th; n0[m] = tni * p.nb;
                }
                const int co0 = ct * p.BN;
                for (int tap = 0; tap < p.KH * p.KW; ++tap) {
                    int kh = tap / p.KW, kw = tap - kh * p.KW;
                    for (int cc = 0; cc < p.cchunks; ++cc) {
                        ptx::mbar_wait(&empty[stage], phase ^ 1);
                        uint8_t* sa = smem + (size_t)stage * stage_bytes;
                        ptx::mbar_expect_tx(&full[stage], (uint32_t)stage_bytes);
                        for (int m = 0; m < p.MT; ++m)
                            ptx::tma_load_4d(sa + m * a_tile, &tmA, &full[stage], (cc * BK) % p.Cx, w0[m] + kw - p.pad, h0[m] + kh - p.pad, n0[m]);
                        ptx::tma_load_2d(sa + a_bytes, &tmB, &full[stage], tap * p.Ci + cc * BK, co0);
                        if (++stage == p.stages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            const uint32_t idesc = ptx::umma_idesc_bf16(BM, p.BN, 0, 0);
            int stage = 0; uint32_t phase = 0;
            int as = 0; uint32_t aphase = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                ptx::mbar_wait(&tempty[as], aphase ^ 1);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(as * p.MT * p.BN);
                for (int ks = 0; ks < p.ksteps; ++ks) {
                    ptx::mbar_wait(&full[stage], phase);
                    ptx::tc_fence_after();
                    const uint32_t sa = ptx::smem_u32(smem + (size_t)stage * stage_bytes);
                    const uint64_t bdesc = ptx::umma_smem_desc(sa + a_bytes, 0, 1024);
                    for (int m = 0; m < p.MT; ++m) {
                        const uint64_t adesc = ptx::umma_smem_desc(sa + m * a_tile, 0, 1024);
#pragma unroll
                        for (int k = 0; k < BK / UMMA_K; ++k)
                            ptx::umma_bf16(d_tmem + (uint32_t)(m * p.BN), adesc + (uint64_t)(k * UMMA_K * 2 / 16),
                                           bdesc + (uint64_t)(k * UMMA_K * 2 / 16), idesc, (ks | k) != 0 ? 1u : 0u);
                    }
                    ptx::umma_commit(&empty[stage]);
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
                ptx::umma_commit(&tfull[as]);
                if (++as == 2) { as = 0; aphase ^= 1; }
            }
        }
    } else {
        epilogue_loop(p, tmem_base, tfull, tempty, warp, lane);
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, 512);
    }
}

// ---------------------------------------------------------------------------------------------------
// 3x3 forward kernel with halo reuse (tw = 8, th = 16): one A load per channel chunk serves nine taps
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NTHREADS_FWD, 1)
conv_fwd_tc_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const FwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int a_tile = p.a_tile_bytes;                              // (th+2) x pitch rows of 128 B, rounded up to 1024
    const int a_stage = p.MT * a_tile;
    const int b_stage = p.BN * BK * 2;
    uint8_t* smemA = smem;
    uint8_t* smemB = smem + (size_t)p.a_stages * a_stage;
    uint64_t* fullA = reinterpret_cast<uint64_t*>(smemB + (size_t)p.b_stages * b_stage);
    uint64_t* emptyA = fullA + p.a_stages;
    uint64_t* fullB = emptyA + p.a_stages;
    uint64_t* emptyB = fullB + p.b_stages;
    uint64_t* tfull = emptyB + p.b_stages;           // [2]
    uint64_t* tempty = tfull + 2;                    // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t a_box_bytes = (uint32_t)((p.th + 2) * p.pitch * 128);

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmA);
        ptx::prefetch_tmap(&tmB);
        for (int i = 0; i < p.a_stages; ++i) { ptx::mbar_init(&fullA[i], 1); ptx::mbar_init(&emptyA[i], 1); }
        for (int i = 0; i < p.b_stages; ++i) { ptx::mbar_init(&fullB[i], 1); ptx::mbar_init(&emptyB[i], 1); }
        for (int i = 0; i < 2; ++i) { ptx::mbar_init(&tfull[i], 1); ptx::mbar_init(&tempty[i], 8); }
        ptx::fence_barrier_init();
    }
    if (warp == 1) ptx::tmem_alloc(tmem_slot, 512);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int sa = 0, sb = 0; uint32_t pa = 0, pb = 0;
            if (p.b_resident && blockIdx.x < p.num_tiles) {
                // all weight tiles once, on ONE barrier: slot cc * ntaps + tap
                ptx::mbar_expect_tx(&fullB[0], (uint32_t)(p.cchunks * p.ntaps * b_stage));
                for (int cc = 0; cc < p.cchunks; ++cc)
                    for (int tap = 0; tap < p.ntaps; ++tap)
                        ptx::tma_load_2d(smemB + (size_t)(cc * p.ntaps + tap) * b_stage, &tmB, &fullB[0], tap * p.Ci + cc * BK, 0);
            }
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                int ct = tile % p.co_tiles; int pt = tile / p.co_tiles;
                int w0[2], h0[2], n0[2];
                for (int m = 0; m < p.MT; ++m) {
                    int q = pt * p.MT + m;
                    int twi = q % p.tiles_w; int t2 = q / p.tiles_w; int thi = t2 % p.tiles_h; int tni = t2 / p.tiles_h;
                    w0[m] = twi * p.tw; h0[m] = thi * p.th; n0[m] = tni;
                }
                const int co0 = ct * p.BN;
                for (int cc = 0; cc < p.cchunks; ++cc) {
                    ptx::mbar_wait(&emptyA[sa], pa ^ 1);
                    ptx::mbar_expect_tx(&fullA[sa], a_box_bytes * (uint32_t)p.MT);
                    for (int m = 0; m < p.MT; ++m)
                        ptx::tma_load_4d(smemA + (size_t)sa * a_stage + m * a_tile, &tmA, &fullA[sa], (cc * BK) % p.Cx, w0[m] - 1, h0[m] - 1, n0[m]);
                    if (++sa == p.a_stages) { sa = 0; pa ^= 1; }
                    if (p.b_resident) continue;
                    for (int tap = 0; tap < p.ntaps; ++tap) {
                        ptx::mbar_wait(&emptyB[sb], pb ^ 1);
                        ptx::mbar_expect_tx(&fullB[sb], (uint32_t)b_stage);
                        ptx::tma_load_2d(smemB + (size_t)sb * b_stage, &tmB, &fullB[sb], tap * p.Ci + cc * BK, co0);
                        if (++sb == p.b_stages) { sb = 0; pb ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = ptx::umma_idesc_bf16(BM, p.BN, 0, 0);
            const uint32_t sbo = (uint32_t)p.pitch * 128u;           // stride between 8-pixel row groups = one halo row
            int sa = 0, sb = 0; uint32_t pa = 0, pb = 0;
            int as = 0; uint32_t aphase = 0;
            if (p.b_resident && blockIdx.x < p.num_tiles) ptx::mbar_wait(&fullB[0], 0);
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                ptx::mbar_wait(&tempty[as], aphase ^ 1);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(as * p.MT * p.BN);
                for (int cc = 0; cc < p.cchunks; ++cc) {
                    ptx::mbar_wait(&fullA[sa], pa);
                    const uint32_t a_addr = ptx::smem_u32(smemA + (size_t)sa * a_stage);
                    for (int tap = 0; tap < p.ntaps; ++tap) {
                        const int kh = p.kh0 + tap / p.ktw, kw = p.kw0 + tap % p.ktw;
                        if (p.b_resident) sb = cc * p.ntaps + tap;
                        else ptx::mbar_wait(&fullB[sb], pb);
                        ptx::tc_fence_after();
                        const uint64_t bdesc = ptx::umma_smem_desc(ptx::smem_u32(smemB + (size_t)sb * b_stage), 0, 1024);
                        const uint32_t row_off = (uint32_t)(kh * p.pitch + kw);
                        for (int m = 0; m < p.MT; ++m) {
                            uint64_t adesc = ptx::umma_smem_desc(a_addr + (uint32_t)(m * a_tile) + row_off * 128u, 0, sbo);
                            if (p.bo_mode) adesc |= (uint64_t)(row_off & 7u) << 49;      // matrix base offset field
#pragma unroll
                            for (int k = 0; k < BK / UMMA_K; ++k)
                                ptx::umma_bf16(d_tmem + (uint32_t)(m * p.BN), adesc + (uint64_t)(k * UMMA_K * 2 / 16),
                                               bdesc + (uint64_t)(k * UMMA_K * 2 / 16), idesc, (cc | tap | k) != 0 ? 1u : 0u);
                        }
                        if (!p.b_resident) {
                            ptx::umma_commit(&emptyB[sb]);
                            if (++sb == p.b_stages) { sb = 0; pb ^= 1; }
                        }
                    }
                    ptx::umma_commit(&emptyA[sa]);
                    if (++sa == p.a_stages) { sa = 0; pa ^= 1; }
                }
                ptx::umma_commit(&tfull[as]);
                if (++as == 2) { as = 0; aphase ^= 1; }
            }
        }
    } else {
        epilogue_loop(p, tmem_base, tfull, tempty, warp, lane);
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, 512);
    }
}

// ---------------------------------------------------------------------------------------------------
// CTA-pair (tcgen05 cta_group::2) halo kernel: M = 256 pixels (128 per CTA of a 2-CTA cluster), N = BN output channels.
// Each CTA loads the halo tile of ITS pixel tile and HALF of the weight tile (BN/2 rows); the leader's single MMA thread issues
// M = 256 UMMAs that read both CTAs' shared memory, so each SM feeds 128 x 16 of A and only BN/2 x 16 of B per instruction
// (ncu on the one-CTA kernels: the SMEM -> tensor operand path, sm__mem_tensor_cycles_active, is 91 % busy at 72 % tensor-pipe
// activity).  Barrier protocol (tools/probes/umma2_probe.cu): operand bytes of both CTAs complete on the LEADER's full
// barriers; empty / accumulator-full barriers are arrived by multicast commits in both CTAs; the epilogue warps of both CTAs
// arrive on the leader's accumulator-empty barrier.
// ---------------------------------------------------------------------------------------------------
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NTHREADS_FWD, 1)
conv_fwd_tc_halo2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const FwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int a_stage = p.a_tile_bytes;                             // one halo tile per CTA and stage
    const int b_stage = (p.BN / 2) * BK * 2;                        // this CTA's half of the weight tile
    uint8_t* smemA = smem;
    uint8_t* smemB = smem + (size_t)p.a_stages * a_stage;
    uint64_t* fullA = reinterpret_cast<uint64_t*>(smemB + (size_t)p.b_stages * b_stage);
    uint64_t* emptyA = fullA + p.a_stages;
    uint64_t* fullB = emptyA + p.a_stages;
    uint64_t* emptyB = fullB + p.b_stages;
    uint64_t* tfull = emptyB + p.b_stages;           // [2]
    uint64_t* tempty = tfull + 2;                    // [2] (the leader's are used)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = ptx::cluster_ctarank();
    const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
    const uint32_t a_box_bytes = (uint32_t)((p.th + 2) * p.pitch * 128);
    const int ptiles = p.tiles_w * p.tiles_h * p.tiles_n;
    // a CONTIGUOUS range of cluster tiles per cluster (neighbouring tiles share halo rows in L2, and the fused GroupNorm statistics
    // are flushed only when the image changes)
    const int tile_begin = (int)((int64_t)cluster_id * p.num_tiles / num_clusters);
    const int tile_end = (int)((int64_t)(cluster_id + 1) * p.num_tiles / num_clusters);

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmA);
        ptx::prefetch_tmap(&tmB);
        for (int i = 0; i < p.a_stages; ++i) { ptx::mbar_init(&fullA[i], 1); ptx::mbar_init(&emptyA[i], 1); }
        for (int i = 0; i < p.b_stages; ++i) { ptx::mbar_init(&fullB[i], 1); ptx::mbar_init(&emptyB[i], 1); }
        for (int i = 0; i < 2; ++i) { ptx::mbar_init(&tfull[i], 1); ptx::mbar_init(&tempty[i], 16); }   // 8 epilogue warps x 2 CTAs
        ptx::fence_barrier_init();
    }
    if (warp == 1) ptx::tmem_alloc2(tmem_slot, 512);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::cluster_sync();                             // barriers and TMEM of BOTH CTAs exist before any remote signal
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int sa = 0, sb = 0; uint32_t pa = 0, pb = 0;
            for (int tile = tile_begin; tile < tile_end; ++tile) {
                const int ct = tile % p.co_tiles, pt = tile / p.co_tiles;
                const int q = pt * 2 + (int)rank;                   // this CTA's pixel tile (may be one past the end)
                const int twi = q % p.tiles_w, t2 = q / p.tiles_w, thi = t2 % p.tiles_h;
                const int w0 = twi * p.tw, h0 = thi * p.th, n0 = (q < ptiles) ? t2 / p.tiles_h : p.N;   // n = N: TMA zero-fill
                const int co0 = ct * p.BN + (int)rank * (p.BN / 2);
                for (int cc = 0; cc < p.cchunks; ++cc) {
                    ptx::mbar_wait(&emptyA[sa], pa ^ 1);
                    if (rank == 0) ptx::mbar_expect_tx(&fullA[sa], 2u * a_box_bytes);
                    ptx::tma_load_4d_2sm(smemA + (size_t)sa * a_stage, &tmA, ptx::mapa_rank(ptx::smem_u32(&fullA[sa]), 0), (cc * BK) % p.Cx, w0 - 1,
                                         h0 - 1, n0);
                    if (++sa == p.a_stages) { sa = 0; pa ^= 1; }
                    for (int tap = 0; tap < p.ntaps; ++tap) {
                        ptx::mbar_wait(&emptyB[sb], pb ^ 1);
                        if (rank == 0) ptx::mbar_expect_tx(&fullB[sb], 2u * (uint32_t)b_stage);
                        ptx::tma_load_2d_2sm(smemB + (size_t)sb * b_stage, &tmB, ptx::mapa_rank(ptx::smem_u32(&fullB[sb]), 0),
                                             tap * p.Ci + cc * BK, co0);
                        if (++sb == p.b_stages) { sb = 0; pb ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && rank == 0) {
            const uint32_t idesc = ptx::umma_idesc_bf16(2 * BM, p.BN, 0, 0);
            const uint32_t sbo = (uint32_t)p.pitch * 128u;
            int sa = 0, sb = 0; uint32_t pa = 0, pb = 0;
            int as = 0; uint32_t aphase = 0;
            for (int tile = tile_begin; tile < tile_end; ++tile) {
                ptx::mbar_wait(&tempty[as], aphase ^ 1);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(as * p.BN);
                for (int cc = 0; cc < p.cchunks; ++cc) {
                    ptx::mbar_wait(&fullA[sa], pa);
                    const uint32_t a_addr = ptx::smem_u32(smemA + (size_t)sa * a_stage);
                    for (int tap = 0; tap < p.ntaps; ++tap) {
                        const int kh = p.kh0 + tap / p.ktw, kw = p.kw0 + tap % p.ktw;
                        ptx::mbar_wait(&fullB[sb], pb);
                        ptx::tc_fence_after();
                        const uint64_t bdesc = ptx::umma_smem_desc(ptx::smem_u32(smemB + (size_t)sb * b_stage), 0, 1024);
                        const uint64_t adesc = ptx::umma_smem_desc(a_addr + (uint32_t)(kh * p.pitch + kw) * 128u, 0, sbo);
#pragma unroll
                        for (int k = 0; k < BK / UMMA_K; ++k)
                            ptx::umma2_bf16(d_tmem, adesc + (uint64_t)(k * UMMA_K * 2 / 16), bdesc + (uint64_t)(k * UMMA_K * 2 / 16), idesc,
                                            (cc | tap | k) != 0 ? 1u : 0u);
                        ptx::umma2_commit(&emptyB[sb]);
                        if (++sb == p.b_stages) { sb = 0; pb ^= 1; }
                    }
                    ptx::umma2_commit(&emptyA[sa]);
                    if (++sa == p.a_stages) { sa = 0; pa ^= 1; }
                }
                ptx::umma2_commit(&tfull[as]);
                if (++as == 2) { as = 0; aphase ^= 1; }
            }
        }
    } else {
        // epilogue: this CTA's 128 pixel rows x BN channels
        const int quarter = warp & 3, half = (warp - 2) >> 2;
        const int row = quarter * 32 + lane;
        const int wi = row % p.tw, hi = row / p.tw;
        const bool pre = p.residual && !p.y_f32 && p.res_prefetch;
        const int nch = (p.BN - half * 32 + 63) / 64;
        // fused GroupNorm statistics: acc[ct][s] = this lane's total (see gn_reduce_chunk) of chunk s of channel tile ct, image gn_n
        float gacc[2][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) gacc[i >> 2][i & 3] = 0.f;
        int gn_n = -1;
        auto gn_flush = [&]() {
            if (gn_n >= 0 && lane < 64 / p.gn_cpg) {
                const int G = p.Co / p.gn_cpg;
#pragma unroll
                for (int ctl = 0; ctl < 2; ++ctl)
#pragma unroll
                    for (int s_ = 0; s_ < 4; ++s_)
                        if (ctl < p.co_tiles && s_ < nch) {
                            const int c_abs = ctl * p.BN + half * 32 + s_ * 64;
                            atomicAdd(p.gn_sums + ((int64_t)gn_n * G + c_abs / p.gn_cpg) * 2 + lane, (double)gacc[ctl][s_]);
                        }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) gacc[i >> 2][i & 3] = 0.f;
        };
        int as = 0; uint32_t aphase = 0;
        for (int tile = tile_begin; tile < tile_end; ++tile) {
            const int ct = tile % p.co_tiles, pt = tile / p.co_tiles;
            const int co0 = ct * p.BN;
            const int q = pt * 2 + (int)rank;
            const int twi = q % p.tiles_w, t2 = q / p.tiles_w, thi = t2 % p.tiles_h, n = t2 / p.tiles_h;
            const int w = twi * p.tw + wi, h = thi * p.th + hi;
            const bool valid = (q < ptiles) && (w < p.W) && (h < p.H) && (n < p.N);
            const int64_t pix = ((int64_t)n * p.H + h) * p.W + w;
            auto fetch = [&](int s_, uint4 (&qv)[4]) {
                if (valid) {
                    const int c = half * 32 + s_ * 64;
                    const uint4* ro = reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(p.residual) + pix * p.Co + co0 + c);
#pragma unroll
                    for (int i = 0; i < 4; ++i) qv[i] = ro[i];
                }
            };
            uint4 qn[4] = {};
            if (pre) fetch(0, qn);
            if (p.gn_sums && q < ptiles && n != gn_n) { gn_flush(); gn_n = n; }
            ptx::mbar_wait(&tfull[as], aphase);
            ptx::tc_fence_after();
            const uint32_t t_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * p.BN);
            for (int s_ = 0; s_ < nch; ++s_) {
                const int c = half * 32 + s_ * 64;
                uint4 qc[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) qc[i] = qn[i];
                if (pre && s_ + 1 < nch) fetch(s_ + 1, qn);
                uint32_t r[32];
                ptx::tmem_ld32(t_addr + (uint32_t)c, r);
                ptx::tmem_ld_wait();
                if (p.gn_sums) {
                    float gs[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) gs[i] = 0.f;
                    epilogue_chunk(p, r, valid, pix, co0, c, pre ? qc : nullptr, gs);
                    // static accumulator indices: (ct, s_) enumerated
#pragma unroll
                    for (int ctl = 0; ctl < 2; ++ctl)
#pragma unroll
                        for (int sl = 0; sl < 4; ++sl)
                            if (ctl == ct && sl == s_) gn_reduce_chunk(p, gs, gacc[ctl][sl], lane);
                } else {
                    epilogue_chunk(p, r, valid, pix, co0, c, pre ? qc : nullptr);
                }
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive_cluster(ptx::mapa_rank(ptx::smem_u32(&tempty[as]), 0));
            if (++as == 2) { as = 0; aphase ^= 1; }
        }
        if (p.gn_sums) gn_flush();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::cluster_sync();                             // the peer may still be reading this CTA's shared memory / TMEM until here
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc2(tmem_base, 512);
    }
}

// ---------------------------------------------------------------------------------------------------
// 3x3 forward kernel for Co tiles of 128 with the operand roles swapped: M = 128 output channels (the packed weight tile is
// the A operand), N = 256 pixels (a 32 x 8 pixel tile; ONE halo box {64 ch, 10, 34} per channel chunk is the B operand of
// all nine taps).  With N = 128 the UMMA reads 8 KB of shared memory per 64 cycles -- exactly the 128 B/clk SMEM port, so
// the 128-channel layers were SMEM-bandwidth-bound at ~800 TFLOP/s; N = 256 needs 12 KB per 128 cycles.  The accumulator
// is D^T [co lanes][pixel columns]; the epilogue writes one pixel per instruction with the warp's 32 lanes covering 32
// consecutive channels (64-byte coalesced segments).
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NTHREADS_FWD, 1)
conv_fwd_tc_halo_t_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, const FwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    constexpr int TH = 32, TW = 8, PITCH = 10, NPIX = 256;
    const int x_stage = p.a_tile_bytes;               // (TH+2) x PITCH rows of 128 B, rounded up to 1024
    const int w_stage = BM * BK * 2;                  // 16 KB weight tile (128 co x 64 ci)
    uint8_t* smemX = smem;
    uint8_t* smemW = smem + (size_t)p.a_stages * x_stage;
    uint64_t* fullX = reinterpret_cast<uint64_t*>(smemW + (size_t)p.b_stages * w_stage);
    uint64_t* emptyX = fullX + p.a_stages;
    uint64_t* fullW = emptyX + p.a_stages;
    uint64_t* emptyW = fullW + p.b_stages;
    uint64_t* tfull = emptyW + p.b_stages;           // [2]
    uint64_t* tempty = tfull + 2;                    // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t x_box_bytes = (uint32_t)((TH + 2) * PITCH * 128);
    // a CONTIGUOUS range of tiles per CTA (see conv_fwd_tc_halo2_kernel)
    const int tile_begin = (int)((int64_t)blockIdx.x * p.num_tiles / gridDim.x);
    const int tile_end = (int)((int64_t)(blockIdx.x + 1) * p.num_tiles / gridDim.x);

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmX);
        ptx::prefetch_tmap(&tmW);
        for (int i = 0; i < p.a_stages; ++i) { ptx::mbar_init(&fullX[i], 1); ptx::mbar_init(&emptyX[i], 1); }
        for (int i = 0; i < p.b_stages; ++i) { ptx::mbar_init(&fullW[i], 1); ptx::mbar_init(&emptyW[i], 1); }
        for (int i = 0; i < 2; ++i) { ptx::mbar_init(&tfull[i], 1); ptx::mbar_init(&tempty[i], 8); }
        ptx::fence_barrier_init();
    }
    if (warp == 1) ptx::tmem_alloc(tmem_slot, 512);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int sx = 0, sw = 0; uint32_t px = 0, pw = 0;
            for (int tile = tile_begin; tile < tile_end; ++tile) {
                const int ct = tile % p.co_tiles, pt = tile / p.co_tiles;
                const int twi = pt % p.tiles_w, t2 = pt / p.tiles_w, thi = t2 % p.tiles_h, n = t2 / p.tiles_h;
                const int w0 = twi * TW, h0 = thi * TH, co0 = ct * BM;
                for (int cc = 0; cc < p.cchunks; ++cc) {
                    ptx::mbar_wait(&emptyX[sx], px ^ 1);
                    ptx::mbar_expect_tx(&fullX[sx], x_box_bytes);
                    ptx::tma_load_4d(smemX + (size_t)sx * x_stage, &tmX, &fullX[sx], (cc * BK) % p.Cx, w0 - 1, h0 - 1, n);
                    if (++sx == p.a_stages) { sx = 0; px ^= 1; }
                    for (int tap = 0; tap < p.ntaps; ++tap) {
                        ptx::mbar_wait(&emptyW[sw], pw ^ 1);
                        ptx::mbar_expect_tx(&fullW[sw], (uint32_t)w_stage);
                        ptx::tma_load_2d(smemW + (size_t)sw * w_stage, &tmW, &fullW[sw], tap * p.Ci + cc * BK, co0);
                        if (++sw == p.b_stages) { sw = 0; pw ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = ptx::umma_idesc_bf16(BM, NPIX, 0, 0);
            int sx = 0, sw = 0; uint32_t px = 0, pw = 0;
            int as = 0; uint32_t aphase = 0;
            for (int tile = tile_begin; tile < tile_end; ++tile) {
                ptx::mbar_wait(&tempty[as], aphase ^ 1);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(as * NPIX);
                for (int cc = 0; cc < p.cchunks; ++cc) {
                    ptx::mbar_wait(&fullX[sx], px);
                    const uint32_t x_addr = ptx::smem_u32(smemX + (size_t)sx * x_stage);
                    for (int tap = 0; tap < p.ntaps; ++tap) {
                        const int kh = p.kh0 + tap / p.ktw, kw = p.kw0 + tap % p.ktw;
                        ptx::mbar_wait(&fullW[sw], pw);
                        ptx::tc_fence_after();
                        const uint64_t adesc = ptx::umma_smem_desc(ptx::smem_u32(smemW + (size_t)sw * w_stage), 0, 1024);
                        const uint64_t bdesc = ptx::umma_smem_desc(x_addr + (uint32_t)((kh * PITCH + kw) * 128), 0, (uint32_t)(PITCH * 128));
#pragma unroll
                        for (int k = 0; k < BK / UMMA_K; ++k)
                            ptx::umma_bf16(d_tmem, adesc + (uint64_t)(k * UMMA_K * 2 / 16), bdesc + (uint64_t)(k * UMMA_K * 2 / 16), idesc,
                                           (cc | tap | k) != 0 ? 1u : 0u);
                        ptx::umma_commit(&emptyW[sw]);
                        if (++sw == p.b_stages) { sw = 0; pw ^= 1; }
                    }
                    ptx::umma_commit(&emptyX[sx]);
                    if (++sx == p.a_stages) { sx = 0; px ^= 1; }
                }
                ptx::umma_commit(&tfull[as]);
                if (++as == 2) { as = 0; aphase ^= 1; }
            }
        }
    } else {
        // ---- transposed epilogue: lane = output channel, TMEM column = pixel of the 32 x 8 tile.  Each 32x32 block is
        // transposed through a padded per-warp shared-memory tile so that global stores are 16-byte vectors along channels.
        const int quarter = warp & 3;
        const int half = (warp - 2) >> 2;                    // two warps per quarter take alternating 32-column chunks
        float* tsm = reinterpret_cast<float*>(tmem_slot + 4) + (size_t)(warp - 2) * (32 * 36);   // [32 pixels][36] fp32
        const int grp = lane & 3, prow = lane >> 2;          // phase 2: lane -> (pixel row within 8, group of 8 channels)
        // fused GroupNorm statistics: partial sums of channels co8..co8+3 / co8+4..co8+7 over this lane's pixels, kept ACROSS tiles
        // and flushed (shuffle reduction over the eight lanes that share `grp`, then double atomics) when the image or the
        // channel tile changes -- one atomic per tile from every warp of every CTA hammered the 64 addresses of the image in flight
        float gsa = 0.f, gqa = 0.f, gsb = 0.f, gqb = 0.f;
        int gn_n = -1, gn_co8 = 0;
        auto gn_flush = [&]() {
            if (gn_n >= 0) {
#pragma unroll
                for (int o = 4; o < 32; o <<= 1) {
                    gsa += __shfl_xor_sync(0xffffffffu, gsa, o); gqa += __shfl_xor_sync(0xffffffffu, gqa, o);
                    gsb += __shfl_xor_sync(0xffffffffu, gsb, o); gqb += __shfl_xor_sync(0xffffffffu, gqb, o);
                }
                if (prow == 0) {
                    const int G = p.Co / p.gn_cpg;
                    if (p.gn_cpg == 4) {
                        double* dst = p.gn_sums + ((int64_t)gn_n * G + gn_co8 / 4) * 2;
                        atomicAdd(dst, (double)gsa); atomicAdd(dst + 1, (double)gqa);
                        atomicAdd(dst + 2, (double)gsb); atomicAdd(dst + 3, (double)gqb);
                    } else {                                   // 8 channels per group; 16: two neighbouring lanes add to the same group
                        double* dst = p.gn_sums + ((int64_t)gn_n * G + gn_co8 / p.gn_cpg) * 2;
                        atomicAdd(dst, (double)(gsa + gsb)); atomicAdd(dst + 1, (double)(gqa + gqb));
                    }
                }
            }
            gsa = gqa = gsb = gqb = 0.f;
        };
        int as = 0; uint32_t aphase = 0;
        for (int tile = tile_begin; tile < tile_end; ++tile) {
            const int ct = tile % p.co_tiles, pt = tile / p.co_tiles;
            const int twi = pt % p.tiles_w, t2 = pt / p.tiles_w, thi = t2 % p.tiles_h, n = t2 / p.tiles_h;
            const int w0 = twi * TW, h0 = thi * TH;
            const int co8 = ct * BM + quarter * 32 + grp * 8;
            if (p.gn_sums && (n != gn_n || co8 != gn_co8)) { gn_flush(); gn_n = n; gn_co8 = co8; }
            float bv[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) bv[u] = p.bias ? __ldg(p.bias + co8 + u) : 0.f;
            // bf16 residual of the NEXT 32-column chunk is requested one chunk ahead (see epilogue_loop)
            const bool pre = p.residual && !p.y_f32 && p.res_prefetch;
            auto fetch = [&](int c, uint4 (&q)[4]) {
#pragma unroll
                for (int it = 0; it < 4; ++it) {
                    const int pl = it * 8 + prow;
                    const int h = h0 + (c >> 3) + (pl >> 3), w = w0 + (pl & 7);
                    if (h < p.H && w < p.W)
                        q[it] = *reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(p.residual) +
                                                                (((int64_t)n * p.H + h) * p.W + w) * p.Co + co8);
                }
            };
            uint4 qn[4] = {};
            if (pre) fetch(half * 32, qn);
            ptx::mbar_wait(&tfull[as], aphase);
            ptx::tc_fence_after();
            const uint32_t t_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * NPIX);
            for (int c = half * 32; c < NPIX; c += 64) {
                uint4 qc[4];
#pragma unroll
                for (int it = 0; it < 4; ++it) qc[it] = qn[it];
                if (pre && c + 64 < NPIX) fetch(c + 64, qn);
                uint32_t r[32];
                ptx::tmem_ld32(t_addr + (uint32_t)c, r);
                ptx::tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) tsm[j * 36 + lane] = __uint_as_float(r[j]);
                __syncwarp();
#pragma unroll
                for (int it = 0; it < 4; ++it) {
                    const int pl = it * 8 + prow;                       // pixel within this 32-column chunk
                    const int h = h0 + (c >> 3) + (pl >> 3), w = w0 + (pl & 7);
                    const float4 v0 = *reinterpret_cast<const float4*>(tsm + pl * 36 + grp * 8);
                    const float4 v1 = *reinterpret_cast<const float4*>(tsm + pl * 36 + grp * 8 + 4);
                    if (h < p.H && w < p.W) {
                        float v[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
                        if (p.bias || p.act != VQB_ACT_NONE || p.gain != 1.0f) {
#pragma unroll
                            for (int u = 0; u < 8; ++u) v[u] = act_f(v[u] + bv[u], p.act, p.alpha) * p.gain;
                        }
                        const int64_t off = (((int64_t)n * p.H + h) * p.W + w) * p.Co + co8;
                        if (p.y_f32) {
                            float* yo = reinterpret_cast<float*>(p.y) + off;
                            if (p.residual) {
                                const float* ro = reinterpret_cast<const float*>(p.residual) + off;
                                float4 q0 = *reinterpret_cast<const float4*>(ro), q1 = *reinterpret_cast<const float4*>(ro + 4);
                                v[0] += q0.x; v[1] += q0.y; v[2] += q0.z; v[3] += q0.w; v[4] += q1.x; v[5] += q1.y; v[6] += q1.z; v[7] += q1.w;
                            }
                            *reinterpret_cast<float4*>(yo) = make_float4(v[0], v[1], v[2], v[3]);
                            *reinterpret_cast<float4*>(yo + 4) = make_float4(v[4], v[5], v[6], v[7]);
                        } else {
                            bf16* yo = reinterpret_cast<bf16*>(p.y) + off;
                            if (p.residual) {
                                const uint4 q = pre ? qc[it] : *reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(p.residual) + off);
                                const __nv_bfloat162* qb = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
                                for (int u = 0; u < 4; ++u) { float2 f = __bfloat1622float2(qb[u]); v[2 * u] += f.x; v[2 * u + 1] += f.y; }
                            }
                            uint4 o;
                            __nv_bfloat162* ob = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
                            for (int u = 0; u < 4; ++u) ob[u] = __floats2bfloat162_rn(v[2 * u], v[2 * u + 1]);
                            *reinterpret_cast<uint4*>(yo) = o;
                        }
                        if (p.gn_sums) {
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                gsa += v[u]; gqa = fmaf(v[u], v[u], gqa);
                                gsb += v[4 + u]; gqb = fmaf(v[4 + u], v[4 + u], gqb);
                            }
                        }
                    }
                }
                __syncwarp();
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&tempty[as]);
            if (++as == 2) { as = 0; aphase ^= 1; }
        }
        if (p.gn_sums) gn_flush();
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, 512);
    }
}


// ---------------------------------------------------------------------------------------------------
// Narrow-OUTPUT 3x3 convolution (Co <= 3: the decoder's 128 -> 3 image head with tanh, autoencoder.py:170,178): per-tap partial
// products + shift-add.  With N = 16 the halo kernel issued 72 UMMAs per 128-pixel tile and a UMMA costs the same ~128 cycles at
// N = 16 as at N = 256 (1.2-1.4 ms for 29 GFLOP).  Here ONE small GEMM per tile computes, for every pixel q of a 16 x 8 HALO
// tile, P[q][(tap, co)] = sum_ci x[q][ci] w[co][ci][tap] (M = 128 pixels, N = 32 >= 9 Co, K = Ci: Ci / 16 UMMAs), and the
// epilogue forms y[h][w][co] = act(bias + sum_tap P[(h + kh - 1, w + kw - 1)][(tap, co)]) for the 14 x 6 interior from a
// shared-memory copy of P.  The packed weight ([32][Ci], rows (tap * Co + co), zero-padded) stays resident in shared memory.
// ---------------------------------------------------------------------------------------------------
struct NarrowOutParams {
    int N, H, W, Ci, Co;
    int tiles_w, tiles_h, num_tiles, cchunks, stages;
    const float* bias;
    void* y;
    int y_f32, act;
    float alpha, gain;
};

__global__ void __launch_bounds__(NTHREADS, 1)
conv_fwd_tc_narrowout_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const NarrowOutParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    constexpr int TWH = 16, THH = 8;                       // halo tile: 16 x 8 pixels = 128 rows of the A operand
    constexpr int OW_T = TWH - 2, OH_T = THH - 2;          // interior: 14 x 6 outputs
    constexpr int PN = 32, PP = 33;                        // P columns (UMMA N) and the padded shared-memory pitch
    const int a_tile = BM * BK * 2;                        // 16 KB per (tile, 64-channel chunk)
    const int w_tile = PN * BK * 2;                        // 4 KB per chunk of the resident weight
    uint8_t* smemW = smem;
    uint8_t* smemA = smem + (size_t)p.cchunks * 4096;      // (4 KB tiles keep the 1024-byte alignment)
    float* ptile = reinterpret_cast<float*>(smemA + (size_t)p.stages * a_tile);        // [2][128][PP]
    uint64_t* full = reinterpret_cast<uint64_t*>(ptile + 2 * BM * PP);
    uint64_t* empty = full + p.stages;
    uint64_t* wfull = empty + p.stages;
    uint64_t* tfull = wfull + 1;                           // [2]
    uint64_t* tempty = tfull + 2;                          // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmA);
        ptx::prefetch_tmap(&tmB);
        for (int i = 0; i < p.stages; ++i) { ptx::mbar_init(&full[i], 1); ptx::mbar_init(&empty[i], 1); }
        ptx::mbar_init(wfull, 1);
        for (int i = 0; i < 2; ++i) { ptx::mbar_init(&tfull[i], 1); ptx::mbar_init(&tempty[i], 4); }
        ptx::fence_barrier_init();
    }
    if (warp == 1) ptx::tmem_alloc(tmem_slot, 64);         // two accumulators of 32 columns
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            if (blockIdx.x < p.num_tiles) {
                ptx::mbar_expect_tx(wfull, (uint32_t)(p.cchunks * w_tile));
                for (int cc = 0; cc < p.cchunks; ++cc) ptx::tma_load_2d(smemW + (size_t)cc * w_tile, &tmB, wfull, cc * BK, 0);
            }
            int stage = 0; uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                const int twi = tile % p.tiles_w, t2 = tile / p.tiles_w, thi = t2 % p.tiles_h, n = t2 / p.tiles_h;
                for (int cc = 0; cc < p.cchunks; ++cc) {
                    ptx::mbar_wait(&empty[stage], phase ^ 1);
                    ptx::mbar_expect_tx(&full[stage], (uint32_t)a_tile);
                    ptx::tma_load_4d(smemA + (size_t)stage * a_tile, &tmA, &full[stage], cc * BK, twi * OW_T - 1, thi * OH_T - 1, n);
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = ptx::umma_idesc_bf16(BM, PN, 0, 0);
            int stage = 0; uint32_t phase = 0;
            int as = 0; uint32_t aphase = 0;
            if (blockIdx.x < p.num_tiles) ptx::mbar_wait(wfull, 0);
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                ptx::mbar_wait(&tempty[as], aphase ^ 1);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(as * PN);
                for (int cc = 0; cc < p.cchunks; ++cc) {
                    ptx::mbar_wait(&full[stage], phase);
                    ptx::tc_fence_after();
                    const uint64_t adesc = ptx::umma_smem_desc(ptx::smem_u32(smemA + (size_t)stage * a_tile), 0, 1024);
                    const uint64_t bdesc = ptx::umma_smem_desc(ptx::smem_u32(smemW + (size_t)cc * w_tile), 0, 1024);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k)
                        ptx::umma_bf16(d_tmem, adesc + (uint64_t)(k * UMMA_K * 2 / 16), bdesc + (uint64_t)(k * UMMA_K * 2 / 16), idesc,
                                       (cc | k) != 0 ? 1u : 0u);
                    ptx::umma_commit(&empty[stage]);
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
                ptx::umma_commit(&tfull[as]);
                if (++as == 2) { as = 0; aphase ^= 1; }
            }
        }
    } else {
        // epilogue warps 2..5: TMEM lane = halo pixel (row-major 8 x 16)
        const int quarter = warp & 3;
        const int q = quarter * 32 + lane;                 // halo pixel of this thread
        const int et = (warp - 2) * 32 + lane;             // 0..127: thread index among the epilogue warps
        int as = 0; uint32_t aphase = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
            const int twi = tile % p.tiles_w, t2 = tile / p.tiles_w, thi = t2 % p.tiles_h, n = t2 / p.tiles_h;
            ptx::mbar_wait(&tfull[as], aphase);
            ptx::tc_fence_after();
            uint32_t r[32];
            ptx::tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * PN), r);
            ptx::tmem_ld_wait();
            ptx::tc_fence_before();
            float* pt = ptile + (size_t)as * BM * PP;
#pragma unroll
            for (int j = 0; j < 27; ++j) pt[q * PP + j] = __uint_as_float(r[j]);
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&tempty[as]);   // the accumulator is free for the tile after next
            asm volatile("bar.sync 1, 128;" ::: "memory");  // P of the whole halo tile is in shared memory
            for (int o = et; o < OW_T * OH_T; o += 128) {
                const int orow = o / OW_T, ocol = o - orow * OW_T;
                const int h = thi * OH_T + orow, w = twi * OW_T + ocol;
                if (h < p.H && w < p.W) {
                    float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
                    for (int kh = 0; kh < 3; ++kh)
#pragma unroll
                        for (int kw = 0; kw < 3; ++kw) {
                            const float* src = pt + ((orow + kh) * TWH + (ocol + kw)) * PP + (kh * 3 + kw) * p.Co;
#pragma unroll
                            for (int co = 0; co < 3; ++co)
                                if (co < p.Co) acc[co] += src[co];
                        }
                    const int64_t off = (((int64_t)n * p.H + h) * p.W + w) * p.Co;
#pragma unroll
                    for (int co = 0; co < 3; ++co)
                        if (co < p.Co) {
                            float t = acc[co] + (p.bias ? __ldg(p.bias + co) : 0.f);
                            t = act_f(t, p.act, p.alpha) * p.gain;
                            if (p.y_f32) reinterpret_cast<float*>(p.y)[off + co] = t;
                            else reinterpret_cast<bf16*>(p.y)[off + co] = __float2bfloat16_rn(t);
                        }
                }
            }
            // (the P buffer `as` is rewritten two tiles later, after another bar.sync of these warps: no second barrier needed)
            if (++as == 2) { as = 0; aphase ^= 1; }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, 64);
    }
}


// ---------------------------------------------------------------------------------------------------
// Narrow-INPUT 3x3 convolution (Ci <= 3: encoder.conv_in 3 -> 128, VGG conv1_1 3 -> 64; autoencoder.py:114, lpips.py): the A
// operand [128 pixels][K = 9 Ci <= 27, zero-padded to 64] is BUILT in shared memory by four producer warps straight from the
// image (one pixel row per thread, written in the UMMA K-major SWIZZLE_128B layout), instead of materialising a 64-channel im2col
// tensor in HBM (537 MB written and read back per call at B = 64) and running a 1x1 convolution over it.  Two UMMA k-steps
// (K = 32) per tile; the packed weight [Co][64] (mode 4) is resident; the epilogue is the common one.
// ---------------------------------------------------------------------------------------------------
constexpr int NTHREADS_NIN = 448;      // TMA / MMA / 8 epilogue warps / 4 producer warps

template <typename TI>
__global__ void __launch_bounds__(NTHREADS_NIN, 1)
conv_fwd_tc_narrowin_kernel(const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmY, const FwdParams p,
                            const TI* __restrict__ x, int Cin, int a_stages, int out_tma) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int a_tile = BM * BK * 2;                        // 16 KB
    const int w_tile = p.BN * BK * 2;                      // one channel tile of the resident weight
    uint8_t* smemW = smem;
    uint8_t* smemA = smem + (size_t)p.co_tiles * w_tile;
    uint8_t* smemO = smemA + (size_t)a_stages * a_tile;    // out_tma: [2 buffers][BN / 64 halves][128 pixels][128 B] output staging (SWIZZLE_128B)
    const int o_half = BM * 128, o_buf = (p.BN / 64) * o_half;
    uint64_t* full = reinterpret_cast<uint64_t*>(smemO + (out_tma ? 2 * (size_t)o_buf : 0));
    uint64_t* empty = full + a_stages;
    uint64_t* wfull = empty + a_stages;
    uint64_t* tfull = wfull + 1;                           // [2]
    uint64_t* tempty = tfull + 2;                          // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // the K columns 9 Cin .. 63 of every A tile are zero and never rewritten: clear the stages once
    for (int i = threadIdx.x; i < a_stages * a_tile / 16; i += NTHREADS_NIN) reinterpret_cast<uint4*>(smemA)[i] = make_uint4(0u, 0u, 0u, 0u);
    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmB);
        for (int i = 0; i < a_stages; ++i) { ptx::mbar_init(&full[i], 4); ptx::mbar_init(&empty[i], 1); }
        ptx::mbar_init(wfull, 1);
        for (int i = 0; i < 2; ++i) { ptx::mbar_init(&tfull[i], 1); ptx::mbar_init(&tempty[i], 8); }
        ptx::fence_barrier_init();
    }
    if (warp == 1) ptx::tmem_alloc(tmem_slot, 256);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // the zero fill is visible to the async proxy (UMMA reads)
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0 && blockIdx.x < p.num_tiles) {
            ptx::mbar_expect_tx(wfull, (uint32_t)(p.co_tiles * w_tile));
            for (int ct = 0; ct < p.co_tiles; ++ct) ptx::tma_load_2d(smemW + (size_t)ct * w_tile, &tmB, wfull, 0, ct * p.BN);
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = ptx::umma_idesc_bf16(BM, p.BN, 0, 0);
            int stage = 0; uint32_t phase = 0;
            int as = 0; uint32_t aphase = 0;
            if (blockIdx.x < p.num_tiles) ptx::mbar_wait(wfull, 0);
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                const int ct = tile % p.co_tiles;
                ptx::mbar_wait(&tempty[as], aphase ^ 1);
                ptx::mbar_wait(&full[stage], phase);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(as * p.BN);
                const uint64_t adesc = ptx::umma_smem_desc(ptx::smem_u32(smemA + (size_t)stage * a_tile), 0, 1024);
                const uint64_t bdesc = ptx::umma_smem_desc(ptx::smem_u32(smemW + (size_t)ct * w_tile), 0, 1024);
#pragma unroll
                for (int k = 0; k < 2; ++k)                  // K = 32 of the 64 columns (the rest is zero on both sides)
                    ptx::umma_bf16(d_tmem, adesc + (uint64_t)(k * UMMA_K * 2 / 16), bdesc + (uint64_t)(k * UMMA_K * 2 / 16), idesc, k != 0 ? 1u : 0u);
                ptx::umma_commit(&empty[stage]);
                ptx::umma_commit(&tfull[as]);
                if (++stage == a_stages) { stage = 0; phase ^= 1; }
                if (++as == 2) { as = 0; aphase ^= 1; }
            }
        }
    } else if (warp < 10 && !out_tma) {
        epilogue_loop(p, tmem_base, tfull, tempty, warp, lane);
    } else if (warp < 10) {
        // ================= epilogue through shared memory + TMA store (bf16 output, no residual) =================
        // thread = pixel row: a 32-channel chunk is 64 B per pixel at a pixel stride of 2 Co bytes -- stored straight to global
        // memory every lane of a warp store hits a different 128-byte line (2048 LSU transactions per 32 KB tile: the kernel was
        // bound by that, 2.8 us per tile).  Staged in the TMA SWIZZLE_128B layout instead and written by two bulk tensor stores per
        // tile (which also clip partial tiles at the image border).
        if (warp == 2 && lane == 0) ptx::prefetch_tmap(&tmY);
        const int quarter = warp & 3, half = (warp - 2) >> 2;
        const int row = quarter * 32 + lane;
        const int nch = (p.BN - half * 32 + 63) / 64;
        int as = 0; uint32_t aphase = 0;
        int buf = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
            const int ct = tile % p.co_tiles, pt = tile / p.co_tiles;
            const int twi = pt % p.tiles_w, t2 = pt / p.tiles_w, thi = t2 % p.tiles_h, n = t2 / p.tiles_h;
            const int co0 = ct * p.BN;
            // the bulk store that read this staging buffer two tiles ago has finished reading it
            if (warp == 2 && lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            asm volatile("bar.sync 2, 256;" ::: "memory");
            ptx::mbar_wait(&tfull[as], aphase);
            ptx::tc_fence_after();
            uint8_t* ob = smemO + (size_t)buf * o_buf;
            for (int s_ = 0; s_ < nch; ++s_) {
                const int c = half * 32 + s_ * 64;
                uint32_t r[32];
                ptx::tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * p.BN + c), r);
                ptx::tmem_ld_wait();
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
                if (p.bias || p.act != VQB_ACT_NONE || p.gain != 1.0f) {
                    const float4* bp = reinterpret_cast<const float4*>(p.bias ? p.bias + co0 + c : nullptr);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float4 q4 = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (p.bias) q4 = __ldg(bp + j);
                        v[4 * j] += q4.x; v[4 * j + 1] += q4.y; v[4 * j + 2] += q4.z; v[4 * j + 3] += q4.w;
                    }
                    if (p.act == VQB_ACT_RELU) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f) * p.gain;
                    } else if (p.act == VQB_ACT_LRELU) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = (v[j] > 0.f ? v[j] : v[j] * p.alpha) * p.gain;
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = act_f(v[j], p.act, p.alpha) * p.gain;
                    }
                }
                uint8_t* orow = ob + (size_t)(c >> 6) * o_half + (size_t)row * 128;
                const int jb = (c & 63) >> 3;                       // first 16-byte chunk of this 32-channel piece inside the 128-byte row
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    uint4 u;
                    __nv_bfloat162* hb = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
                    for (int e = 0; e < 4; ++e) hb[e] = __floats2bfloat162_rn(v[i * 8 + 2 * e], v[i * 8 + 2 * e + 1]);
                    *reinterpret_cast<uint4*>(orow + (((jb + i) ^ (row & 7)) << 4)) = u;
                }
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&tempty[as]);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> visible to the bulk store
            asm volatile("bar.sync 2, 256;" ::: "memory");
            if (warp == 2 && lane == 0) {
                for (int hf = 0; hf < p.BN / 64; ++hf)
                    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                                     reinterpret_cast<uint64_t>(&tmY)),
                                 "r"(ptx::smem_u32(ob + (size_t)hf * o_half)), "r"(co0 + hf * 64), "r"(twi * p.tw), "r"(thi * p.th), "r"(n)
                                 : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            buf ^= 1;
            if (++as == 2) { as = 0; aphase ^= 1; }
        }
        if (warp == 2 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");    // all output bytes written before exit
    } else {
        // ================= A-tile producers: thread = pixel row of the 16 x 8 tile =================
        const int row = (warp - 10) * 32 + lane;
        const int wi = row % p.tw, hi = row / p.tw;
        int stage = 0; uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
            const int pt = tile / p.co_tiles;
            const int twi = pt % p.tiles_w, t2 = pt / p.tiles_w, thi = t2 % p.tiles_h, n = t2 / p.tiles_h;
            const int h = thi * p.th + hi, w = twi * p.tw + wi;
            {   // this thread's pixel of the tile two iterations ahead -> L2 (the image is streamed: every tile is a first touch)
                const int tile2 = tile + 2 * (int)gridDim.x;
                if (tile2 < p.num_tiles) {
                    const int pt2 = tile2 / p.co_tiles;
                    const int twi2 = pt2 % p.tiles_w, t22 = pt2 / p.tiles_w, thi2 = t22 % p.tiles_h, n2 = t22 / p.tiles_h;
                    const int h2 = thi2 * p.th + hi, w2 = twi2 * p.tw + wi;
                    if (h2 < p.H && w2 < p.W)
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(x + (((int64_t)n2 * p.H + h2) * p.W + w2) * Cin) : "memory");
                }
            }
            // 32 K values: (tap, c) for tap < 9, c < Cin (9 Cin <= 27), zero beyond
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0.f;
            if (h < p.H && w < p.W) {
#pragma unroll
                for (int tap = 0; tap < 9; ++tap) {
                    const int ih = h + tap / 3 - 1, iw = w + tap % 3 - 1;
                    if (ih >= 0 && ih < p.H && iw >= 0 && iw < p.W) {
                        const TI* src = x + (((int64_t)n * p.H + ih) * p.W + iw) * Cin;
#pragma unroll
                        for (int c = 0; c < 3; ++c)
                            if (c < Cin) v[tap * 3 + c] = (float)src[c];
                    }
                }
            }
            ptx::mbar_wait(&empty[stage], phase ^ 1);
            uint8_t* dst = smemA + (size_t)stage * a_tile + (size_t)row * 128;
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
                uint4 u;
                __nv_bfloat162* hb = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
                for (int e = 0; e < 4; ++e) hb[e] = __floats2bfloat162_rn(v[ch * 8 + 2 * e], v[ch * 8 + 2 * e + 1]);
                *reinterpret_cast<uint4*>(dst + ((ch ^ (row & 7)) << 4)) = u;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&full[stage]);
            if (++stage == a_stages) { stage = 0; phase ^= 1; }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, 256);
    }
}

// ---------------------------------------------------------------------------------------------------
struct WgradParams {
    int N, H, W, Ci, Co, KH, KW, pad;
    int tw, th, nb, tiles_w, tiles_h, tiles_n;
    int CN, co_tiles, ci_tiles, stages;
    int ptiles_total, ptiles_per_split;
    float* dwp;
};

__global__ void __launch_bounds__(NTHREADS, 1)
conv_wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmDy, const __grid_constant__ CUtensorMap tmX, const WgradParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int blk_bytes = BK * 128;                     // one [64 pixels][64 channels] bf16 box = 8 KB
    const int a_bytes = (BM / 64) * blk_bytes;          // dy: 128 co
    const int b_bytes = (p.CN / 64) * blk_bytes;        // x : CN ci
    const int stage_bytes = a_bytes + b_bytes;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
    uint64_t* empty = full + p.stages;
    uint64_t* tfull = empty + p.stages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // tile decode: blockIdx.x = ((tap * ci_tiles + cit) * co_tiles + cot), blockIdx.y = pixel split
    int cot = blockIdx.x % p.co_tiles; int t1 = blockIdx.x / p.co_tiles; int cit = t1 % p.ci_tiles; int tap = t1 / p.ci_tiles;
    const int kh = tap / p.KW, kw = tap - kh * p.KW;
    const int co0 = cot * BM, ci0 = cit * p.CN;
    const int pt_begin = blockIdx.y * p.ptiles_per_split;
    int pt_end = pt_begin + p.ptiles_per_split; if (pt_end > p.ptiles_total) pt_end = p.ptiles_total;
    const int nsteps = pt_end - pt_begin;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmDy);
        ptx::prefetch_tmap(&tmX);
        for (int i = 0; i < p.stages; ++i) { ptx::mbar_init(&full[i], 1); ptx::mbar_init(&empty[i], 1); }
        ptx::mbar_init(tfull, 1);
        ptx::fence_barrier_init();
    }
    uint32_t tcols = 32; while ((int)tcols < p.CN) tcols <<= 1;
    if (warp == 1) ptx::tmem_alloc(tmem_slot, tcols);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (nsteps > 0) {
        if (warp == 0) {
            if (lane == 0) {
                int stage = 0; uint32_t phase = 0;
                for (int pt = pt_begin; pt < pt_end; ++pt) {
                    int twi = pt % p.tiles_w; int t2 = pt / p.tiles_w; int thi = t2 % p.tiles_h; int tni = t2 / p.tiles_h;
                    int w0 = twi * p.tw, h0 = thi * p.th, n0 = tni * p.nb;
                    ptx::mbar_wait(&empty[stage], phase ^ 1);
                    uint8_t* sa = smem + (size_t)stage * stage_bytes;
                    ptx::mbar_expect_tx(&full[stage], (uint32_t)stage_bytes);
                    for (int j = 0; j < BM / 64; ++j)
                        ptx::tma_load_4d(sa + j * blk_bytes, &tmDy, &full[stage], co0 + j * 64, w0, h0, n0);
                    for (int j = 0; j < p.CN / 64; ++j)
                        ptx::tma_load_4d(sa + a_bytes + j * blk_bytes, &tmX, &full[stage], ci0 + j * 64, w0 + kw - p.pad, h0 + kh - p.pad, n0);
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
            }
        } else if (warp == 1) {
            if (lane == 0) {
                const uint32_t idesc = ptx::umma_idesc_bf16(BM, p.CN, 1, 1);
                int stage = 0; uint32_t phase = 0;
                for (int s = 0; s < nsteps; ++s) {
                    ptx::mbar_wait(&full[stage], phase);
                    ptx::tc_fence_after();
                    const uint32_t sa = ptx::smem_u32(smem + (size_t)stage * stage_bytes);
                    // MN-major: LBO = stride between 64-channel blocks (one 8 KB box), SBO = 1024 (8 pixel rows)
                    const uint64_t adesc = ptx::umma_smem_desc(sa, (uint32_t)blk_bytes, 1024);
                    const uint64_t bdesc = ptx::umma_smem_desc(sa + a_bytes, (uint32_t)blk_bytes, 1024);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k)      // 16 pixel rows = 2048 bytes per UMMA
                        ptx::umma_bf16(tmem_base, adesc + (uint64_t)(k * UMMA_K * 128 / 16), bdesc + (uint64_t)(k * UMMA_K * 128 / 16),
                                       idesc, (s | k) != 0 ? 1u : 0u);
                    ptx::umma_commit(&empty[stage]);
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
                ptx::umma_commit(tfull);
            }
        } else {
            const int quarter = warp & 3;
            const int co = co0 + quarter * 32 + lane;
            ptx::mbar_wait(tfull, 0);
            ptx::tc_fence_after();
            const uint32_t t_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
            float* out = p.dwp + ((int64_t)tap * p.Ci + ci0) * p.Co + co;
            for (int c = 0; c < p.CN; c += 32) {
                uint32_t r[32];
                ptx::tmem_ld32(t_addr + (uint32_t)c, r);
                ptx::tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) atomicAdd(out + (int64_t)(c + j) * p.Co, __uint_as_float(r[j]));
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, tcols);
    }
}


// ---------------------------------------------------------------------------------------------------
// Weight gradient with a NARROW operand (3 channels: the image side of encoder.conv_in, or dy of the 128 -> 3 head):
//   dwp[(tap * 3 + c)][cw] += sum_pix nar[pix + tap][c] * wide[pix][cw],   dwp = [64][Cw] fp32 (rows >= 27 stay zero).
// The [64 pixels][64 im2col columns] operand of every reduction step is built in shared memory by producer warps (the layout of
// conv_fwd_tc_narrowin_kernel's A tile, read here as an MN-major operand: rows = pixels = the reduction index); the wide
// operand arrives by TMA.  Replaces im2col-to-HBM + the generic 1x1 weight-gradient kernel.
// ---------------------------------------------------------------------------------------------------
constexpr int NTHREADS_NWG = 320;      // TMA / MMA / 4 epilogue warps / 4 producer warps (two groups alternating steps)

template <typename TI>
__global__ void __launch_bounds__(NTHREADS_NWG, 1)
conv_wgrad_tc_narrow_kernel(const __grid_constant__ CUtensorMap tmWide, const WgradParams p, const TI* __restrict__ nar, int Cn) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int blk_bytes = BK * 128;                     // one [64 pixels][64 channels] bf16 block = 8 KB
    const int a_bytes = (BM / 64) * blk_bytes;          // wide operand: 128 channels
    const int stage_bytes = a_bytes + blk_bytes;        // + the built [64 pixels][64 columns] block
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
    uint64_t* empty = full + p.stages;
    uint64_t* tfull = empty + p.stages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cot = blockIdx.x;                          // 128-channel tile of the wide operand
    const int co0 = cot * BM;
    const int pt_begin = blockIdx.y * p.ptiles_per_split;
    int pt_end = pt_begin + p.ptiles_per_split; if (pt_end > p.ptiles_total) pt_end = p.ptiles_total;
    const int nsteps = pt_end - pt_begin;

    // columns 27 .. 63 of every built block are zero and never rewritten
    for (int st = 0; st < p.stages; ++st)
        for (int i = threadIdx.x; i < blk_bytes / 16; i += NTHREADS_NWG)
            reinterpret_cast<uint4*>(smem + (size_t)st * stage_bytes + a_bytes)[i] = make_uint4(0u, 0u, 0u, 0u);
    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmWide);
        for (int i = 0; i < p.stages; ++i) { ptx::mbar_init(&full[i], 3); ptx::mbar_init(&empty[i], 1); }   // TMA expect_tx + 2 producer warps
        ptx::mbar_init(tfull, 1);
        ptx::fence_barrier_init();
    }
    if (warp == 1) ptx::tmem_alloc(tmem_slot, 64);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (nsteps > 0) {
        if (warp == 0) {
            if (lane == 0) {
                int stage = 0; uint32_t phase = 0;
                for (int pt = pt_begin; pt < pt_end; ++pt) {
                    const int twi = pt % p.tiles_w, t2 = pt / p.tiles_w, thi = t2 % p.tiles_h, tni = t2 / p.tiles_h;
                    ptx::mbar_wait(&empty[stage], phase ^ 1);
                    uint8_t* sa = smem + (size_t)stage * stage_bytes;
                    ptx::mbar_expect_tx(&full[stage], (uint32_t)a_bytes);
                    for (int j = 0; j < BM / 64; ++j)
                        ptx::tma_load_4d(sa + j * blk_bytes, &tmWide, &full[stage], co0 + j * 64, twi * p.tw, thi * p.th, tni * p.nb);
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
            }
        } else if (warp == 1) {
            if (lane == 0) {
                const uint32_t idesc = ptx::umma_idesc_bf16(BM, 64, 1, 1);
                int stage = 0; uint32_t phase = 0;
                for (int s_ = 0; s_ < nsteps; ++s_) {
                    ptx::mbar_wait(&full[stage], phase);
                    ptx::tc_fence_after();
                    const uint32_t sa = ptx::smem_u32(smem + (size_t)stage * stage_bytes);
                    const uint64_t adesc = ptx::umma_smem_desc(sa, (uint32_t)blk_bytes, 1024);
                    const uint64_t bdesc = ptx::umma_smem_desc(sa + a_bytes, (uint32_t)blk_bytes, 1024);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k)
                        ptx::umma_bf16(tmem_base, adesc + (uint64_t)(k * UMMA_K * 128 / 16), bdesc + (uint64_t)(k * UMMA_K * 128 / 16), idesc,
                                       (s_ | k) != 0 ? 1u : 0u);
                    ptx::umma_commit(&empty[stage]);
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
                ptx::umma_commit(tfull);
            }
        } else if (warp < 6) {
            const int quarter = warp & 3;
            const int co = co0 + quarter * 32 + lane;
            ptx::mbar_wait(tfull, 0);
            ptx::tc_fence_after();
            uint32_t r[32];
            ptx::tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16), r);
            ptx::tmem_ld_wait();
            // D[cw lane][j column]: only the 27 real rows of dwp
#pragma unroll
            for (int j = 0; j < 27; ++j) atomicAdd(p.dwp + (int64_t)j * p.Co + co, __uint_as_float(r[j]));
        } else {
            // producers: group g = (warp - 6) / 2 builds the blocks of steps s with (s & 1) == g; thread = pixel row of the block
            const int g = (warp - 6) >> 1;
            const int row = ((warp - 6) & 1) * 32 + lane;                  // 0..63
            const int wi = row % p.tw, r2 = row / p.tw, hi = r2 % p.th, ni = r2 / p.th;
            for (int s_ = g; s_ < nsteps; s_ += 2) {
                const int pt = pt_begin + s_;
                const int stage = s_ % p.stages;
                const uint32_t phase = (uint32_t)((s_ / p.stages) & 1);
                const int twi = pt % p.tiles_w, t2 = pt / p.tiles_w, thi = t2 % p.tiles_h, tni = t2 / p.tiles_h;
                const int w = twi * p.tw + wi, h = thi * p.th + hi, n = tni * p.nb + ni;
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = 0.f;
                if (h < p.H && w < p.W && n < p.N) {
#pragma unroll
                    for (int tap = 0; tap < 9; ++tap) {
                        const int ih = h + tap / 3 - 1, iw = w + tap % 3 - 1;
                        if (ih >= 0 && ih < p.H && iw >= 0 && iw < p.W) {
                            const TI* src = nar + (((int64_t)n * p.H + ih) * p.W + iw) * Cn;
#pragma unroll
                            for (int c = 0; c < 3; ++c) v[tap * 3 + c] = (float)src[c];
                        }
                    }
                }
                ptx::mbar_wait(&empty[stage], phase ^ 1);
                uint8_t* dst = smem + (size_t)stage * stage_bytes + a_bytes + (size_t)row * 128;
#pragma unroll
                for (int ch = 0; ch < 4; ++ch) {
                    uint4 u;
                    __nv_bfloat162* hb = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
                    for (int e = 0; e < 4; ++e) hb[e] = __floats2bfloat162_rn(v[ch * 8 + 2 * e], v[ch * 8 + 2 * e + 1]);
                    *reinterpret_cast<uint4*>(dst + ((ch ^ (row & 7)) << 4)) = u;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&full[stage]);
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, 64);
    }
}

// ---------------------------------------------------------------------------------------------------
// 3x3 weight gradient with halo reuse: a CTA owns (128 co) x (CN ci) x (one kernel row kh = three taps) and a range
// of 8x8-pixel tiles.  Per tile it loads dy [64 px][128 co] once and ONE x halo box {64 ch, 10, 8} per channel block;
// the B operand of tap kw is that halo addressed with a start offset of kw rows and SBO = halo pitch (MN-major:
// an 8-row K atom = 8 consecutive pixels along w).  Three accumulators (one per kw) live in TMEM.
// ---------------------------------------------------------------------------------------------------
struct WgradHaloParams {
    int N, H, W, Ci, Co;
    int tiles_w, tiles_h, CN, co_tiles, ci_tiles, stages;
    int ptiles_total, ptiles_per_split;
    float* dwp;
    int ktw, kh0, kw0;   // taps per kernel row and first frame row / column (3, 0, 0 for a 3x3 kernel; see FwdParams)
};

__global__ void __launch_bounds__(NTHREADS, 1)
conv_wgrad_tc_halo_kernel(const __grid_constant__ CUtensorMap tmDy, const __grid_constant__ CUtensorMap tmX, const WgradHaloParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    constexpr int PITCH = 10;                           // halo row pitch in pixels (8 + 2)
    constexpr int DY_BLK = 64 * 128;                    // [64 pixels][64 co] = 8 KB
    constexpr int X_BLK = 8 * PITCH * 128;              // [8 rows][10 cols][64 ci] = 10 KB (multiple of 1024)
    const int a_bytes = (BM / 64) * DY_BLK;
    const int b_bytes = (p.CN / 64) * X_BLK;
    const int stage_bytes = a_bytes + b_bytes;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
    uint64_t* empty = full + p.stages;
    uint64_t* tfull = empty + p.stages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // blockIdx.x = ((kh * ci_tiles + cit) * co_tiles + cot), blockIdx.y = pixel split
    int cot = blockIdx.x % p.co_tiles; int t1 = blockIdx.x / p.co_tiles; int cit = t1 % p.ci_tiles; int kh = t1 / p.ci_tiles;
    const int co0 = cot * BM, ci0 = cit * p.CN;
    const int pt_begin = blockIdx.y * p.ptiles_per_split;
    int pt_end = pt_begin + p.ptiles_per_split; if (pt_end > p.ptiles_total) pt_end = p.ptiles_total;
    const int nsteps = pt_end - pt_begin;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmDy);
        ptx::prefetch_tmap(&tmX);
        for (int i = 0; i < p.stages; ++i) { ptx::mbar_init(&full[i], 1); ptx::mbar_init(&empty[i], 1); }
        ptx::mbar_init(tfull, 1);
        ptx::fence_barrier_init();
    }
    uint32_t tcols = 32; while ((int)tcols < 3 * p.CN) tcols <<= 1;
    if (warp == 1) ptx::tmem_alloc(tmem_slot, tcols);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (nsteps > 0) {
        if (warp == 0) {
            if (lane == 0) {
                int stage = 0; uint32_t phase = 0;
                for (int pt = pt_begin; pt < pt_end; ++pt) {
                    int twi = pt % p.tiles_w; int t2 = pt / p.tiles_w; int thi = t2 % p.tiles_h; int n = t2 / p.tiles_h;
                    int w0 = twi * 8, h0 = thi * 8;
                    ptx::mbar_wait(&empty[stage], phase ^ 1);
                    uint8_t* sa = smem + (size_t)stage * stage_bytes;
                    ptx::mbar_expect_tx(&full[stage], (uint32_t)stage_bytes);
                    for (int j = 0; j < BM / 64; ++j)
                        ptx::tma_load_4d(sa + j * DY_BLK, &tmDy, &full[stage], co0 + j * 64, w0, h0, n);
                    for (int j = 0; j < p.CN / 64; ++j)
                        ptx::tma_load_4d(sa + a_bytes + j * X_BLK, &tmX, &full[stage], ci0 + j * 64, w0 - 1, h0 + p.kh0 + kh - 1, n);
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
            }
        } else if (warp == 1) {
            if (lane == 0) {
                const uint32_t idesc = ptx::umma_idesc_bf16(BM, p.CN, 1, 1);
                int stage = 0; uint32_t phase = 0;
                for (int s = 0; s < nsteps; ++s) {
                    ptx::mbar_wait(&full[stage], phase);
                    ptx::tc_fence_after();
                    const uint32_t sa = ptx::smem_u32(smem + (size_t)stage * stage_bytes);
                    const uint64_t adesc = ptx::umma_smem_desc(sa, (uint32_t)DY_BLK, 1024);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {              // UMMA j: pixel rows 2j, 2j+1 of the 8x8 tile (16 K-rows)
#pragma unroll
                        for (int kw = 0; kw < 3; ++kw) {
                            if (kw >= p.ktw) break;
                            const uint64_t bdesc = ptx::umma_smem_desc(sa + a_bytes + (uint32_t)((2 * j * PITCH + p.kw0 + kw) * 128),
                                                                       (uint32_t)X_BLK, (uint32_t)(PITCH * 128));
                            ptx::umma_bf16(tmem_base + (uint32_t)(kw * p.CN), adesc + (uint64_t)(j * 2048 / 16), bdesc, idesc,
                                           (s | j) != 0 ? 1u : 0u);
                        }
                    }
                    ptx::umma_commit(&empty[stage]);
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
                ptx::umma_commit(tfull);
            }
        } else {
            const int quarter = warp & 3;
            const int co = co0 + quarter * 32 + lane;
            ptx::mbar_wait(tfull, 0);
            ptx::tc_fence_after();
            for (int kw = 0; kw < p.ktw; ++kw) {
                const uint32_t t_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(kw * p.CN);
                float* out = p.dwp + ((int64_t)(kh * p.ktw + kw) * p.Ci + ci0) * p.Co + co;
                for (int c = 0; c < p.CN; c += 32) {
                    uint32_t r[32];
                    ptx::tmem_ld32(t_addr + (uint32_t)c, r);
                    ptx::tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) atomicAdd(out + (int64_t)(c + j) * p.Co, __uint_as_float(r[j]));
                }
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, tcols);
    }
}

int sm_count() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

// VQB_HALO_MODE: 0 = generic kernel only; 1 = halo pitch 10 (default); 2 = pitch 16; 3 = pitch 16 + base-offset field;
// 4 = pitch 10 + base-offset field.  (2-4 exist to pin down the descriptor semantics on hardware; see tests.)
int halo_mode() {
    static int mode = -1;
    if (mode < 0) {
        const char* e = getenv("VQB_HALO_MODE");
        mode = e ? atoi(e) : 1;
        if (mode < 0 || mode > 4) mode = 1;
    }
    return mode;
}

}  // namespace

// test hook: override the halo mode at run time (-1 = re-read the environment)
extern "C" void vqb_set_halo_mode(int mode);
static int g_halo_override = -1;
extern "C" void vqb_set_halo_mode(int mode) { g_halo_override = mode; }

// Cx: channels of the x tensor.  Cx == Ci is the plain bf16 convolution.  Split-precision operands (strict numeric mode): x holds
// [hi | lo] bf16 halves of C fp32 channels (Cx = 2C) and the packed weight holds `terms` blocks per tap (Ci = terms * C):
// [wh | wh | wl] (3 terms: xh.wh + xl.wh + xh.wl) or [wh | wh | wl | wl] (4 terms, + xl.wl); the k loop's channel coordinate
// wraps modulo Cx, so the SAME kernels accumulate all terms in one fp32 TMEM accumulator.
// `sub` (optional): a T x T-tap sub-convolution y[n,h,w,:] = sum_{a,b<T} x[n, h+off+a, w+off+b, :] . wp[:, (a*T+b)*Ci ...] over an input of
// Hx x Wx pixels (zero outside), executed by the 3x3 halo kernels on the tap sub-rectangle [off+1, off+1+T)^2 of their frame.
struct SubConv { int Hx, Wx, T, off; };

static int conv_fwd_tc_impl(const void* x, const void* wp, const float* bias, const void* residual, void* y, int y_dtype, int N,
                            int H, int W, int Ci, int Co, int KH, int KW, int pad, int act, float alpha, float gain,
                            cudaStream_t stream, int Cx, double* gn_sums, int gn_groups, const SubConv* sub) {
    VQB_CHECK_ARG(N > 0 && H > 0 && W > 0 && KH > 0 && KW > 0 && pad >= 0, "conv2d_fwd(tcgen05): bad geometry");
    const int Hx = sub ? sub->Hx : H, Wx = sub ? sub->Wx : W;
    const int wk = sub ? sub->T * sub->T * Ci : KH * KW * Ci;            // reduction length of the packed weight
    VQB_CHECK_ARG(Ci % 64 == 0 && (Co % 64 == 0 || Co <= 16), "conv2d_fwd(tcgen05): need Ci %% 64 == 0 and (Co %% 64 == 0 or Co <= 16) (got %d, %d)", Ci, Co);
    VQB_CHECK_ARG(H + 2 * pad - KH + 1 == H && W + 2 * pad - KW + 1 == W, "conv2d_fwd(tcgen05): only 'same' convolutions");
    VQB_CHECK_ARG(((uintptr_t)x & 15) == 0 && ((uintptr_t)wp & 15) == 0 && ((uintptr_t)y & 15) == 0, "conv2d_fwd(tcgen05): unaligned pointer");
    FwdParams p;
    if (Cx <= 0) Cx = Ci;
    VQB_CHECK_ARG(Cx % 64 == 0 && (2 * Ci) % Cx == 0, "conv2d_fwd(tcgen05): bad split-operand channel count %d for Ci %d", Cx, Ci);
    p.N = N; p.H = H; p.W = W; p.Ci = Ci; p.Co = Co; p.KH = KH; p.KW = KW; p.pad = pad; p.Cx = Cx;
    p.narrow = (Co % 64 != 0);
    p.gn_sums = nullptr; p.gn_cpg = 0;
    p.ntaps = 9; p.ktw = 3; p.kh0 = 0; p.kw0 = 0; p.b_resident = 0;
    if (sub) { p.ntaps = sub->T * sub->T; p.ktw = sub->T; p.kh0 = p.kw0 = sub->off + 1; }
    if (gn_sums) {
        const int cpg = (gn_groups > 0 && Co % gn_groups == 0) ? Co / gn_groups : 0;
        if (p.narrow || !(cpg == 4 || cpg == 8 || cpg == 16)) {
            vqb_set_error("conv2d_fwd(tcgen05): fused GroupNorm statistics need 4, 8 or 16 channels per group (Co=%d, groups=%d)", Co, gn_groups);
            return VQB_ERR_UNSUPPORTED;
        }
        const bool halo_ok = KH == 3 && KW == 3 && pad == 1 && H >= 16 && W >= 8 &&
                             ((Co % 256 == 0 && Co / 256 <= 2) || (Co % 256 != 0 && Co % 128 == 0 && H >= 32));
        if (!halo_ok) {
            vqb_set_error("conv2d_fwd(tcgen05): fused GroupNorm statistics are built into the CTA-pair and swapped-operand 3x3 kernels only");
            return VQB_ERR_UNSUPPORTED;
        }
        p.gn_sums = gn_sums; p.gn_cpg = cpg;
    }
    // narrow heads: UMMA N = 16, the weight box rows beyond Co are zero-filled by TMA
    p.BN = p.narrow ? 16 : ((Co % 256 == 0) ? 256 : ((Co % 128 == 0) ? 128 : 64));
    p.co_tiles = p.narrow ? 1 : Co / p.BN;
    p.cchunks = Ci / BK;
    p.ksteps = KH * KW * p.cchunks;
    p.bias = bias; p.residual = residual; p.y = y; p.y_f32 = (y_dtype == VQB_F32); p.act = act; p.alpha = alpha; p.gain = gain;
    static const int res_prefetch = getenv("VQB_RES_PREFETCH") ? atoi(getenv("VQB_RES_PREFETCH")) : 1;
    p.res_prefetch = res_prefetch;
    p.pitch = 0; p.bo_mode = 0; p.a_tile_bytes = 0; p.a_stages = 0; p.b_stages = 0;
    const int mode = g_halo_override >= 0 ? g_halo_override : halo_mode();
    const bool halo = mode != 0 && KH == 3 && KW == 3 && pad == 1 && H >= 16 && W >= 8;
    CUtensorMap tmA, tmB;
    int rc = make_weight_map(&tmB, wp, Co, wk, p.BN); if (rc) return rc;
    // VQB_CONV_2CTA: 0 = never, 1 (default) = 256-channel output tiles, 2 = also 128-channel tiles (slower than the swapped-operand
    // kernel below: with N = 128 a CTA pair still feeds 128 x 16 of A per 64 cycles -- measured 1.08 vs 1.42 PFLOP/s)
    static const int use_2cta = getenv("VQB_CONV_2CTA") ? atoi(getenv("VQB_CONV_2CTA")) : 1;
    if (halo && mode == 1 && use_2cta && !p.narrow && (p.BN == 256 || (use_2cta == 2 && p.BN == 128))) {
        // CTA pairs: 2 x 128 pixels x BN channels per cluster tile (see conv_fwd_tc_halo2_kernel)
        p.tw = 8; p.th = 16; p.nb = 1; p.MT = 1; p.pitch = 10; p.bo_mode = 0; p.stages = 0;
        p.tiles_w = (W + 7) / 8; p.tiles_h = (H + 15) / 16; p.tiles_n = N;
        const int ptiles = p.tiles_w * p.tiles_h * p.tiles_n;
        p.num_tiles = ((ptiles + 1) / 2) * p.co_tiles;             // cluster tiles
        p.a_tile_bytes = (((p.th + 2) * p.pitch * 128) + 1023) / 1024 * 1024;
        const int b_stage = (p.BN / 2) * BK * 2;
        p.a_stages = 3;
        p.b_stages = (SMEM_LIMIT - 2048 - p.a_stages * p.a_tile_bytes) / b_stage; if (p.b_stages > 12) p.b_stages = 12;
        rc = make_weight_map(&tmB, wp, Co, wk, p.BN / 2); if (rc) return rc;
        rc = make_act_map(&tmA, x, N, Hx, Wx, Cx, p.pitch, p.th + 2, 1); if (rc) return rc;
        size_t smem = (size_t)p.a_stages * p.a_tile_bytes + (size_t)p.b_stages * b_stage + 1024 + 512;
        VQB_CUDA(cudaFuncSetAttribute(conv_fwd_tc_halo2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int clusters = p.num_tiles < sm_count() / 2 ? p.num_tiles : sm_count() / 2;
        conv_fwd_tc_halo2_kernel<<<2 * clusters, NTHREADS_FWD, smem, stream>>>(tmA, tmB, p);
        VQB_CHECK_LAUNCH("conv2d_fwd_tc_halo2");
        return VQB_OK;
    }
    if (halo && mode == 1 && p.BN == 128 && !p.narrow && H >= 32) {
        // Co tiles of 128: swapped operand roles (M = co, N = 256 pixels), see conv_fwd_tc_halo_t_kernel
        p.tw = 8; p.th = 32; p.nb = 1; p.MT = 1; p.pitch = 10; p.bo_mode = 0; p.stages = 0;
        p.tiles_w = (W + 7) / 8; p.tiles_h = (H + 31) / 32; p.tiles_n = N;
        p.co_tiles = Co / BM;
        p.num_tiles = p.tiles_w * p.tiles_h * N * p.co_tiles;
        p.a_tile_bytes = ((34 * 10 * 128) + 1023) / 1024 * 1024;
        const int w_stage = BM * BK * 2;
        p.a_stages = 2;
        const int tr_bytes = 8 * 32 * 36 * 4;                       // epilogue transpose tiles (8 warps)
        p.b_stages = (SMEM_LIMIT - 2048 - tr_bytes - p.a_stages * p.a_tile_bytes) / w_stage; if (p.b_stages > 12) p.b_stages = 12;
        rc = make_act_map(&tmA, x, N, Hx, Wx, Cx, 10, 34, 1); if (rc) return rc;
        size_t smem = (size_t)p.a_stages * p.a_tile_bytes + (size_t)p.b_stages * w_stage + 1024 + 512 + tr_bytes;
        VQB_CUDA(cudaFuncSetAttribute(conv_fwd_tc_halo_t_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int grid = p.num_tiles < sm_count() ? p.num_tiles : sm_count();
        conv_fwd_tc_halo_t_kernel<<<grid, NTHREADS_FWD, smem, stream>>>(tmA, tmB, p);
        VQB_CHECK_LAUNCH("conv2d_fwd_tc_halo_t");
        return VQB_OK;
    }
    if (halo) {
        p.tw = 8; p.th = 16; p.nb = 1;
        p.tiles_w = (W + 7) / 8; p.tiles_h = (H + 15) / 16; p.tiles_n = N;
        p.pitch = (mode == 2 || mode == 3) ? 16 : 10;
        if (sub) { p.pitch = 10; }
        p.bo_mode = (mode == 3 || mode == 4) ? 1 : 0;
        p.a_tile_bytes = (((p.th + 2) * p.pitch * 128) + 1023) / 1024 * 1024;
        const int ptiles = p.tiles_w * p.tiles_h * p.tiles_n;
        p.MT = (p.BN <= 128 && ptiles >= 2 * sm_count()) ? 2 : 1;
        p.num_tiles = ((ptiles + p.MT - 1) / p.MT) * p.co_tiles;
        const int a_stage = p.MT * p.a_tile_bytes, b_stage = p.BN * BK * 2;
        p.a_stages = 2;
        static const int allow_resident = getenv("VQB_CONV_BRES") ? atoi(getenv("VQB_CONV_BRES")) : 1;
        if (allow_resident && p.co_tiles == 1 && p.ntaps * p.cchunks * b_stage <= 72 * 1024) {
            p.b_resident = 1;
            p.b_stages = p.ntaps * p.cchunks;
            p.a_stages = (SMEM_LIMIT - 2048 - p.b_stages * b_stage) / a_stage; if (p.a_stages > 4) p.a_stages = 4;
        } else {
            p.b_stages = (SMEM_LIMIT - 2048 - p.a_stages * a_stage) / b_stage; if (p.b_stages > 12) p.b_stages = 12;
        }
        VQB_CHECK_ARG(p.b_stages >= 2, "conv2d_fwd(tcgen05 halo): shared memory budget");
        p.stages = 0;
        rc = make_act_map(&tmA, x, N, Hx, Wx, Cx, p.pitch, p.th + 2, 1); if (rc) return rc;
        size_t smem = (size_t)p.a_stages * a_stage + (size_t)p.b_stages * b_stage + 1024 + 512;
        VQB_CUDA(cudaFuncSetAttribute(conv_fwd_tc_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int grid = p.num_tiles < sm_count() ? p.num_tiles : sm_count();
        conv_fwd_tc_halo_kernel<<<grid, NTHREADS_FWD, smem, stream>>>(tmA, tmB, p);
        VQB_CHECK_LAUNCH("conv2d_fwd_tc_halo");
        return VQB_OK;
    }
    if (sub) { vqb_set_error("conv2d_fwd_sub(tcgen05): needs the 3x3 halo kernels (output of at least 16 x 8 pixels)"); return VQB_ERR_UNSUPPORTED; }
    pick_tile(BM, H, W, p.tw, p.th, p.nb);
    p.tiles_w = (W + p.tw - 1) / p.tw; p.tiles_h = (H + p.th - 1) / p.th; p.tiles_n = (N + p.nb - 1) / p.nb;
    const int ptiles = p.tiles_w * p.tiles_h * p.tiles_n;
    p.MT = (p.BN <= 128 && ptiles >= 2 * sm_count()) ? 2 : 1;     // 2 x 128 pixels per CTA tile when the accumulators fit TMEM
    p.num_tiles = ((ptiles + p.MT - 1) / p.MT) * p.co_tiles;
    const int stage_bytes = p.MT * BM * BK * 2 + p.BN * BK * 2;
    p.stages = (SMEM_LIMIT - 2048) / stage_bytes; if (p.stages > 8) p.stages = 8;
    rc = make_act_map(&tmA, x, N, Hx, Wx, Cx, p.tw, p.th, p.nb); if (rc) return rc;
    size_t smem = (size_t)p.stages * stage_bytes + 1024 + 256;
    VQB_CUDA(cudaFuncSetAttribute(conv_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int grid = p.num_tiles < sm_count() ? p.num_tiles : sm_count();
    conv_fwd_tc_kernel<<<grid, NTHREADS_FWD, smem, stream>>>(tmA, tmB, p);
    VQB_CHECK_LAUNCH("conv2d_fwd_tc");
    return VQB_OK;
}

int vqb_conv2d_fwd_tc(const void* x, const void* wp, const float* bias, const void* residual, void* y, int y_dtype, int N,
                      int H, int W, int Ci, int Co, int KH, int KW, int pad, int act, float alpha, float gain,
                      cudaStream_t stream, int Cx, double* gn_sums, int gn_groups) {
    return conv_fwd_tc_impl(x, wp, bias, residual, y, y_dtype, N, H, W, Ci, Co, KH, KW, pad, act, alpha, gain, stream, Cx, gn_sums, gn_groups,
                            nullptr);
}

static int sub_check(const char* who, int N, int Hx, int Wx, int H, int W, int Ci, int Co, int T, int off) {
    if (!(N > 0 && Hx > 0 && Wx > 0 && H >= 16 && W >= 8 && Ci % 64 == 0 && Co % 64 == 0 && (T == 2 || T == 3) && off >= -1 && off + T <= 2)) {
        vqb_set_error("%s: unsupported sub-convolution (N=%d in %dx%d out %dx%d Ci=%d Co=%d T=%d off=%d): needs an output of >= 16 x 8 pixels, "
                      "64-multiple channels, T in {2,3}, the taps inside the 3x3 frame", who, N, Hx, Wx, H, W, Ci, Co, T, off);
        return VQB_ERR_UNSUPPORTED;
    }
    return VQB_OK;
}

extern "C" int vqb_conv2d_sub_supported(int N, int Hx, int Wx, int H, int W, int Ci, int Co, int T, int off) {
    const int mode = g_halo_override >= 0 ? g_halo_override : halo_mode();
    return mode == 1 && N > 0 && Hx > 0 && Wx > 0 && H >= 16 && W >= 8 && Ci % 64 == 0 && Co % 64 == 0 && (T == 2 || T == 3) && off >= -1 &&
           off + T <= 2;
}

extern "C" int vqb_conv2d_fwd_sub(const void* x, const void* wp, const float* bias, const void* residual, void* y, int y_dtype, int N, int Hx,
                                  int Wx, int H, int W, int Ci, int Co, int T, int off, int act, float act_alpha, float gain, void* stream) {
    VQB_CHECK_ARG(x && wp && y, "conv2d_fwd_sub: null pointer");
    int rc = sub_check("conv2d_fwd_sub", N, Hx, Wx, H, W, Ci, Co, T, off); if (rc) return rc;
    SubConv sc{Hx, Wx, T, off};
    return conv_fwd_tc_impl(x, wp, bias, residual, y, y_dtype, N, H, W, Ci, Co, 3, 3, 1, act, act_alpha, gain, as_stream(stream), Ci, nullptr, 0, &sc);
}

static int conv_wgrad_tc_impl(const void* x, const void* dy, float* dwp, int N, int H, int W, int Ci, int Co, int KH, int KW,
                              int pad, cudaStream_t stream, const SubConv* sub) {
    VQB_CHECK_ARG(N > 0 && H > 0 && W > 0 && KH > 0 && KW > 0 && pad >= 0, "conv2d_wgrad(tcgen05): bad geometry");
    VQB_CHECK_ARG(Ci % 64 == 0 && Co % 128 == 0, "conv2d_wgrad(tcgen05): need Ci %% 64 == 0 and Co %% 128 == 0 (got %d, %d)", Ci, Co);
    VQB_CHECK_ARG(H + 2 * pad - KH + 1 == H && W + 2 * pad - KW + 1 == W, "conv2d_wgrad(tcgen05): only 'same' convolutions");
    const int mode = g_halo_override >= 0 ? g_halo_override : halo_mode();
    if (sub && !(mode != 0 && H >= 8 && W >= 8)) { vqb_set_error("conv2d_wgrad_sub(tcgen05): needs the halo kernel"); return VQB_ERR_UNSUPPORTED; }
    if (mode != 0 && KH == 3 && KW == 3 && pad == 1 && H >= 8 && W >= 8) {
        WgradHaloParams h;
        h.N = N; h.H = H; h.W = W; h.Ci = Ci; h.Co = Co;
        h.ktw = 3; h.kh0 = 0; h.kw0 = 0;
        if (sub) { h.ktw = sub->T; h.kh0 = h.kw0 = sub->off + 1; }
        h.tiles_w = (W + 7) / 8; h.tiles_h = (H + 7) / 8;
        h.CN = (Ci % 128 == 0) ? 128 : 64;
        h.co_tiles = Co / BM; h.ci_tiles = Ci / h.CN;
        const int stage_bytes = (BM / 64) * 64 * 128 + (h.CN / 64) * 8 * 10 * 128;
        h.stages = (SMEM_LIMIT - 2048) / stage_bytes; if (h.stages > 8) h.stages = 8;
        h.ptiles_total = h.tiles_w * h.tiles_h * N;
        const int out_tiles = h.ktw * h.ci_tiles * h.co_tiles;           // one CTA per (kernel row, ci tile, co tile)
        // ONE wave: TMEM (3 x 128 accumulator columns) admits one CTA per SM, so out_tiles * splits must not exceed the SM count
        // -- ceil(2 * SMs / out_tiles) gave 297 CTAs for the 128-channel layers: two waves plus ONE straggler CTA (ncu: SMs
        // active 57 % of the launch) and twice the atomic-combine traffic.
        static const int waves = getenv("VQB_WGRAD_WAVES") ? atoi(getenv("VQB_WGRAD_WAVES")) : 1;
        int splits = waves == 2 ? (sm_count() * 2 + out_tiles - 1) / out_tiles : (out_tiles <= sm_count() ? sm_count() / out_tiles : 1);
        int max_splits = (h.ptiles_total + 7) / 8; if (max_splits < 1) max_splits = 1;
        if (splits > max_splits) splits = max_splits;
        if (splits < 1) splits = 1;
        h.ptiles_per_split = (h.ptiles_total + splits - 1) / splits;
        splits = (h.ptiles_total + h.ptiles_per_split - 1) / h.ptiles_per_split;
        h.dwp = dwp;
        CUtensorMap tmDy, tmX;
        int rc = make_act_map(&tmDy, dy, N, H, W, Co, 8, 8, 1); if (rc) return rc;
        rc = make_act_map(&tmX, x, N, sub ? sub->Hx : H, sub ? sub->Wx : W, Ci, 10, 8, 1); if (rc) return rc;
        size_t smem = (size_t)h.stages * stage_bytes + 1024 + 256;
        VQB_CUDA(cudaFuncSetAttribute(conv_wgrad_tc_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dim3 grid(out_tiles, splits);
        conv_wgrad_tc_halo_kernel<<<grid, NTHREADS, smem, stream>>>(tmDy, tmX, h);
        VQB_CHECK_LAUNCH("conv2d_wgrad_tc_halo");
        return VQB_OK;
    }
    WgradParams p;
    p.N = N; p.H = H; p.W = W; p.Ci = Ci; p.Co = Co; p.KH = KH; p.KW = KW; p.pad = pad;
    pick_tile(BK, H, W, p.tw, p.th, p.nb);
    p.tiles_w = (W + p.tw - 1) / p.tw; p.tiles_h = (H + p.th - 1) / p.th; p.tiles_n = (N + p.nb - 1) / p.nb;
    p.CN = (Ci % 256 == 0) ? 256 : ((Ci % 128 == 0) ? 128 : 64);
    p.co_tiles = Co / BM; p.ci_tiles = Ci / p.CN;
    const int stage_bytes = (BM / 64 + p.CN / 64) * BK * 128;
    p.stages = (SMEM_LIMIT - 2048) / stage_bytes; if (p.stages > 8) p.stages = 8;
    p.ptiles_total = p.tiles_w * p.tiles_h * p.tiles_n;
    const int out_tiles = KH * KW * p.ci_tiles * p.co_tiles;
    int splits = out_tiles <= sm_count() ? sm_count() / out_tiles : 1;       // one wave of CTAs (see the halo launcher)
    int max_splits = (p.ptiles_total + 7) / 8; if (max_splits < 1) max_splits = 1;   // >= 8 pixel tiles (512 pixels) per CTA
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    p.ptiles_per_split = (p.ptiles_total + splits - 1) / splits;
    splits = (p.ptiles_total + p.ptiles_per_split - 1) / p.ptiles_per_split;
    p.dwp = dwp;
    CUtensorMap tmDy, tmX;
    int rc = make_act_map(&tmDy, dy, N, H, W, Co, p.tw, p.th, p.nb); if (rc) return rc;
    rc = make_act_map(&tmX, x, N, H, W, Ci, p.tw, p.th, p.nb); if (rc) return rc;
    size_t smem = (size_t)p.stages * stage_bytes + 1024 + 256;
    VQB_CUDA(cudaFuncSetAttribute(conv_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(out_tiles, splits);
    conv_wgrad_tc_kernel<<<grid, NTHREADS, smem, stream>>>(tmDy, tmX, p);
    VQB_CHECK_LAUNCH("conv2d_wgrad_tc");
    return VQB_OK;
}

int vqb_conv2d_wgrad_tc(const void* x, const void* dy, float* dwp, int N, int H, int W, int Ci, int Co, int KH, int KW,
                        int pad, cudaStream_t stream) {
    return conv_wgrad_tc_impl(x, dy, dwp, N, H, W, Ci, Co, KH, KW, pad, stream, nullptr);
}

// weight gradient of vqb_conv2d_fwd_sub: dwp[(a*T+b)*Ci + ci][co] (fp32, caller zero-fills) += sum_pix x[n, h+off+a, w+off+b, ci] dy[n,h,w,co]
extern "C" int vqb_conv2d_wgrad_sub(const void* x, const void* dy, float* dwp, int N, int Hx, int Wx, int H, int W, int Ci, int Co, int T, int off,
                                    void* stream) {
    VQB_CHECK_ARG(x && dy && dwp, "conv2d_wgrad_sub: null pointer");
    int rc = sub_check("conv2d_wgrad_sub", N, Hx, Wx, H, W, Ci, Co, T, off); if (rc) return rc;
    VQB_CHECK_ARG(Co % 128 == 0, "conv2d_wgrad_sub: Co must be a multiple of 128");
    SubConv sc{Hx, Wx, T, off};
    return conv_wgrad_tc_impl(x, dy, dwp, N, H, W, Ci, Co, 3, 3, 1, as_stream(stream), &sc);
}

// y = act(conv3x3(x, w) + bias) * gain for Co <= 3 (see conv_fwd_tc_narrowout_kernel); wp = [32][Ci] bf16, row (tap * Co + co) =
// w[co][:, tap] (tap = kh * 3 + kw), rows >= 9 Co zero
extern "C" int vqb_conv2d_fwd_narrowout(const void* x, const void* wp, const float* bias, void* y, int y_dtype, int N, int H, int W, int Ci,
                                        int Co, int act, float act_alpha, float gain, void* stream) {
    VQB_CHECK_ARG(x && wp && y && N > 0 && H > 0 && W > 0, "conv2d_fwd_narrowout: bad arguments");
    VQB_CHECK_ARG(Ci % 64 == 0 && Ci <= 512 && Co >= 1 && Co <= 3, "conv2d_fwd_narrowout: needs Ci %% 64 == 0, Ci <= 512, Co <= 3 (got %d, %d)", Ci, Co);
    VQB_CHECK_ARG(((uintptr_t)x & 15) == 0 && ((uintptr_t)wp & 15) == 0, "conv2d_fwd_narrowout: unaligned pointer");
    NarrowOutParams p;
    p.N = N; p.H = H; p.W = W; p.Ci = Ci; p.Co = Co;
    p.tiles_w = (W + 13) / 14; p.tiles_h = (H + 5) / 6;
    p.num_tiles = p.tiles_w * p.tiles_h * N;
    p.cchunks = Ci / BK;
    static const int stages_env = getenv("VQB_NARROW_OUT_STAGES") ? atoi(getenv("VQB_NARROW_OUT_STAGES")) : 0;
    const size_t fixed = (size_t)p.cchunks * 4096 + (size_t)2 * BM * 33 * 4 + 256 + 1024;
    p.stages = (int)((SMEM_LIMIT - fixed) / (BM * BK * 2));             // as many 16 KB input boxes in flight as fit (11 at Ci = 128)
    if (stages_env > 0 && stages_env < p.stages) p.stages = stages_env;
    if (p.stages > 12) p.stages = 12;
    p.bias = bias; p.y = y; p.y_f32 = (y_dtype == VQB_F32); p.act = act; p.alpha = act_alpha; p.gain = gain;
    CUtensorMap tmA, tmB;
    int rc = make_act_map(&tmA, x, N, H, W, Ci, 16, 8, 1); if (rc) return rc;
    rc = make_weight_map(&tmB, wp, 32, Ci, 32); if (rc) return rc;
    size_t smem = fixed + (size_t)p.stages * BM * BK * 2;
    VQB_CUDA(cudaFuncSetAttribute(conv_fwd_tc_narrowout_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int grid = p.num_tiles < sm_count() ? p.num_tiles : sm_count();
    conv_fwd_tc_narrowout_kernel<<<grid, NTHREADS, smem, as_stream(stream)>>>(tmA, tmB, p);
    VQB_CHECK_LAUNCH("conv2d_fwd_narrowout");
    return VQB_OK;
}

// y = act(conv3x3(x, w) + bias [+ residual]) * gain for Ci == 3 packed as K index (tap * 3 + c) (see conv_fwd_tc_narrowin_kernel);
// x NHWC [N,H,W,3] fp32 or bf16 (rounded to bf16 for the tensor cores, as the im2col route does); wp = mode-4 packed weight [Co][64]
extern "C" int vqb_conv2d_fwd_narrowin(const void* x, int x_dtype, const void* wp, const float* bias, const void* residual, void* y,
                                       int y_dtype, int N, int H, int W, int Ci, int Co, int act, float act_alpha, float gain, void* stream) {
    VQB_CHECK_ARG(x && wp && y && N > 0 && H > 0 && W > 0, "conv2d_fwd_narrowin: bad arguments");
    VQB_CHECK_ARG(Ci == 3 && Co % 64 == 0 && Co <= 512, "conv2d_fwd_narrowin: needs Ci == 3, Co %% 64 == 0, Co <= 512 (got %d, %d)", Ci, Co);
    VQB_CHECK_ARG(x_dtype == VQB_F32 || x_dtype == VQB_BF16, "conv2d_fwd_narrowin: x must be fp32 or bf16");
    FwdParams p;
    p.N = N; p.H = H; p.W = W; p.Ci = 64; p.Co = Co; p.KH = 1; p.KW = 1; p.pad = 0; p.Cx = 64;
    p.tw = 16; p.th = 8; p.nb = 1; p.MT = 1;
    p.tiles_w = (W + 15) / 16; p.tiles_h = (H + 7) / 8; p.tiles_n = N;
    p.BN = (Co % 128 == 0) ? 128 : 64;
    p.co_tiles = Co / p.BN;
    p.num_tiles = p.tiles_w * p.tiles_h * N * p.co_tiles;
    p.stages = 0; p.ksteps = 0; p.cchunks = 1;
    p.pitch = 0; p.bo_mode = 0; p.a_tile_bytes = 0; p.a_stages = 0; p.b_stages = 0;
    p.bias = bias; p.residual = residual; p.y = y; p.y_f32 = (y_dtype == VQB_F32); p.act = act; p.narrow = 0;
    static const int res_prefetch = getenv("VQB_RES_PREFETCH") ? atoi(getenv("VQB_RES_PREFETCH")) : 1;
    p.res_prefetch = res_prefetch;
    p.alpha = act_alpha; p.gain = gain; p.gn_sums = nullptr; p.gn_cpg = 0;
    p.ntaps = 1; p.ktw = 1; p.kh0 = 0; p.kw0 = 0; p.b_resident = 1;
    CUtensorMap tmB, tmY;
    int rc = make_weight_map(&tmB, wp, Co, 64, p.BN); if (rc) return rc;
    const int a_stages = 4;
    // bf16 output without a residual leaves through shared memory + TMA bulk stores (VQB_NARROW_IN_TMA=0: the common pixel-major epilogue)
    static const int tma_out = getenv("VQB_NARROW_IN_TMA") ? atoi(getenv("VQB_NARROW_IN_TMA")) : 1;
    const int out_tma = (tma_out && y_dtype == VQB_BF16 && !residual && ((uintptr_t)y & 15) == 0) ? 1 : 0;
    if (out_tma) { rc = make_act_map(&tmY, y, N, H, W, Co, 16, 8, 1); if (rc) return rc; }
    else tmY = tmB;
    size_t smem = (size_t)Co * 128 + (size_t)a_stages * BM * BK * 2 + (out_tma ? (size_t)2 * (p.BN / 64) * BM * 128 : 0) + 256 + 1024;
    cudaStream_t st = as_stream(stream);
    int grid = p.num_tiles < sm_count() ? p.num_tiles : sm_count();
    if (x_dtype == VQB_F32) {
        VQB_CUDA(cudaFuncSetAttribute(conv_fwd_tc_narrowin_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        conv_fwd_tc_narrowin_kernel<float><<<grid, NTHREADS_NIN, smem, st>>>(tmB, tmY, p, (const float*)x, Ci, a_stages, out_tma);
    } else {
        VQB_CUDA(cudaFuncSetAttribute(conv_fwd_tc_narrowin_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        conv_fwd_tc_narrowin_kernel<bf16><<<grid, NTHREADS_NIN, smem, st>>>(tmB, tmY, p, (const bf16*)x, Ci, a_stages, out_tma);
    }
    VQB_CHECK_LAUNCH("conv2d_fwd_narrowin");
    return VQB_OK;
}

// dwp [64][Cw] fp32 (caller zero-fills; rows (tap * 3 + c), tap = kh * 3 + kw) += sum_pix narrow[pix + tap - (1,1)][c] * wide[pix][cw]:
// the weight gradient of a 3x3 'same' convolution whose one side has 3 channels (see conv_wgrad_tc_narrow_kernel).  narrow: NHWC
// [N,H,W,3] fp32 or bf16; wide: NHWC [N,H,W,Cw] bf16, Cw a multiple of 128.
extern "C" int vqb_conv2d_wgrad_narrow(const void* narrow, int n_dtype, const void* wide, float* dwp, int N, int H, int W, int Cn, int Cw,
                                       void* stream) {
    VQB_CHECK_ARG(narrow && wide && dwp && N > 0 && H > 0 && W > 0, "conv2d_wgrad_narrow: bad arguments");
    VQB_CHECK_ARG(Cn == 3 && Cw % 128 == 0, "conv2d_wgrad_narrow: needs 3 narrow channels and Cw %% 128 == 0 (got %d, %d)", Cn, Cw);
    VQB_CHECK_ARG(n_dtype == VQB_F32 || n_dtype == VQB_BF16, "conv2d_wgrad_narrow: narrow operand must be fp32 or bf16");
    WgradParams p;
    p.N = N; p.H = H; p.W = W; p.Ci = 64; p.Co = Cw; p.KH = 1; p.KW = 1; p.pad = 0;
    pick_tile(BK, H, W, p.tw, p.th, p.nb);
    p.tiles_w = (W + p.tw - 1) / p.tw; p.tiles_h = (H + p.th - 1) / p.th; p.tiles_n = (N + p.nb - 1) / p.nb;
    p.CN = 64; p.co_tiles = Cw / BM; p.ci_tiles = 1;
    const int stage_bytes = (BM / 64 + 1) * BK * 128;
    p.stages = 8;
    p.ptiles_total = p.tiles_w * p.tiles_h * p.tiles_n;
    int splits = p.co_tiles <= sm_count() ? sm_count() / p.co_tiles : 1;
    int max_splits = (p.ptiles_total + 7) / 8; if (max_splits < 1) max_splits = 1;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    p.ptiles_per_split = (p.ptiles_total + splits - 1) / splits;
    splits = (p.ptiles_total + p.ptiles_per_split - 1) / p.ptiles_per_split;
    p.dwp = dwp;
    CUtensorMap tmWide;
    int rc = make_act_map(&tmWide, wide, N, H, W, Cw, p.tw, p.th, p.nb); if (rc) return rc;
    size_t smem = (size_t)p.stages * stage_bytes + 1024 + 256;
    cudaStream_t st = as_stream(stream);
    dim3 grid(p.co_tiles, splits);
    if (n_dtype == VQB_F32) {
        VQB_CUDA(cudaFuncSetAttribute(conv_wgrad_tc_narrow_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        conv_wgrad_tc_narrow_kernel<float><<<grid, NTHREADS_NWG, smem, st>>>(tmWide, p, (const float*)narrow, Cn);
    } else {
        VQB_CUDA(cudaFuncSetAttribute(conv_wgrad_tc_narrow_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        conv_wgrad_tc_narrow_kernel<bf16><<<grid, NTHREADS_NWG, smem, st>>>(tmWide, p, (const bf16*)narrow, Cn);
    }
    VQB_CHECK_LAUNCH("conv2d_wgrad_narrow");
    return VQB_OK;
}
