// ONE-launch vector quantisation (reference: VectorQuantizer.forward / EMAVectorQuantizer.forward,
// vqvae/modules/vector_quantizers.py:23-61, 128-180; EntropyVectorQuantizer's argmin part :337-343):
//   pairwise L2 distance of N latents against the K x D codebook -> argmin -> gather -> straight-through value -> sum (e-z)^2
//   -> code histogram -> EMA cluster sums, with the exact-index contract of the strict kernel (vq.cu: fp32 distances in the
//   reference's operation order, first-index ties).
//
// Structure (cluster of 2 CTAs = 256 latent rows, tcgen05 cta_group::2, 10 warps per CTA):
//   prologue   warps 2-9 read this CTA's 128 z rows ONCE from HBM (fp32, coalesced), round them to fp16 and write them straight
//              into the UMMA K-major SWIZZLE_128B layout in shared memory, where they stay for the whole kernel (A operand,
//              64 KB at D = 256); |z|^2 comes from the same pass.
//   warp 0     streams the fp16 copy of the codebook (prepared once per codebook version by vq_prep_codebook or by the EMA
//              update itself) through a 6-stage TMA ring; each CTA stages only HALF of every 256-code tile (16 KB).
//   warp 1     (leader CTA) issues M = 256, N = 256 UMMAs (kind::f16, fp32 accumulators in TMEM, double buffered).  ONE fp16
//              product per k-step: the dot products carry a RIGOROUSLY bounded error of ~2^-10 |z||e_k| per code (eps_k below), a
//              third of the tensor work of a bf16 hi/lo split that would give 2^-16.
//   warps 2-9  scan each accumulator tile: d~ = |e|^2 - 2 dot, a running UPPER bound ub = min_k (d~_k + eps_k) of the smallest
//              fp32 distance, and an online CANDIDATE LIST per row: every code whose lower bound d~_k - eps_k does not exceed the
//              running ub (a superset of the codes that can still be the fp32 argmin at the end).  The bounds are per code
//              (eps_k grows with |e_k|): a trained codebook mixes large used codes with decayed unused ones.  After the last tile
//              a row with a single surviving candidate is decided -- no other code can have the smallest fp32 distance; for the
//              others the survivors, and only they (typically two), are re-evaluated EXACTLY (sequential fp32 FMA chain,
//              (|z|^2 + |e|^2) - 2 dot: the strict kernel's arithmetic bit for bit), so the indices equal the strict kernel's in
//              every case.
//   finish     same warps, same launch: idx (int64), q = z + (e - z), sum (e-z)^2, histogram and EMA cluster sums
//              (vector red.global.add), z re-read from L2.
//
// Roofline: algorithmic bytes 4ND + 4KD + 4ND + 8N (+8K + 12KD); tensor work 2NKD FLOP.  At K = 1024 the op sits
// at the fp16 ridge (215 FLOP/B): the search is tensor-bound, prologue and finish are HBM / L2-bound.
#include "common.cuh"
#include "ptx.cuh"
#include <mutex>

namespace {

constexpr int TM = 128;          // latent rows per CTA (256 per cluster)
constexpr int TN = 256;          // codes per accumulator tile (128 staged per CTA)
constexpr int TK = 64;           // k elements per shared-memory tile row (128 bytes)
constexpr int NTH = 320;         // warp 0 TMA, warp 1 MMA, warps 2-9 prologue / scan / exact / finish
constexpr int ESTAGES = 6;
constexpr int CAP = 32;          // candidate-list entries per (row, column half): 16 overflowed on the reference's INITIAL codebook
                                 // (U(+-1/K) codes against O(1) latents: dozens of codes within the error band at K = 8192 -- 200
                                 // of 8192 rows took the exact full scan, 1.7 ms in the first steps of the K = 8192 configuration)
constexpr int Z_TILE = TM * TK * 2;          // 16 KB
constexpr int E_TILE = (TN / 2) * TK * 2;    // 16 KB (this CTA's half)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode3() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

int make_code_map(CUtensorMap* m, const void* base, int64_t rows, int cols, int box_rows) {
    EncodeTiledFn enc = get_encode3();
    if (!enc) { vqb_set_error("cuTensorMapEncodeTiled unavailable"); return VQB_ERR_CUDA; }
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { vqb_set_error("cuTensorMapEncodeTiled(vq_fused) failed: %d", (int)r); return VQB_ERR_CUDA; }
    return VQB_OK;
}

// How far the APPROXIMATE distance d~_k = |e_k|^2 - 2 (zh . eh_k) from the tensor cores (zh, eh = z, e rounded to fp16) may lie
// from the strict kernel's fp32 value d_k (both without the row's |z|^2, which is the same operand for every code of a row
// and cancels in comparisons):  |d~_k - d_k| <= eps_k = EPS_A |z||e_k| + EPS_B (|z|^2 + |e_k|^2) + EPS_R |e_k|^2 + EPS_G sqrt(D) (|z| + |e_k|)
//   * operand rounding: |zh_i - z_i| <= 2^-12 |z_i| (11 significant bits, round to nearest) or <= 2^-25 below the fp16 normal
//     range, likewise e: |zh.eh - z.e| <= (2^-11 + 2^-24) |z||e| + 2^-25 sqrt(D) (|z| + |e|)  (Cauchy-Schwarz);
//     fp32 accumulation in TMEM over 16 k-steps (not round-to-nearest): <= 2^-17 |z||e|;
//     on the distance (x2):  (2^-10 + 2^-16 + 2^-23) |z||e| + 2^-24 sqrt(D) (|z| + |e|);
//   * the fp32 evaluation itself: one rounding of (|z|^2 + |e|^2) and one of the subtraction, each <= 1 ulp of a value
//     <= (|z|+|e|)^2 <= 2 (|z|^2+|e|^2), i.e. <= 2^-22 (|z|^2+|e|^2), and the sequential 256-term FMA chain of the dot product,
//     <= 256 * 2^-24 |z||e| (x2 on the distance) = 2^-15 |z||e|.
//   * the bookkeeping of the scan itself (d~ = fma(-2, dot, |e|^2), d~ - eps, d~ + eps, ub + eps: at most four roundings, per pair of
//     codes compared, of values <= |e|^2 + 2|z||e| + eps): <= 2^-22 |e|^2 + 2^-21 |z||e| (+ a 2^-22 fraction of eps) per code.
//   EPS_A = 1.04e-3 >= 2^-10 + 2^-15 + 2^-16 + 2^-21 + 2^-23 (= 1.0231e-3: the 1.6 % slack covers the roundings of evaluating eps
//   itself, the approximate square roots and the approximate |z|^2 of the prologue), EPS_B = 2.4e-7 >= 2^-22 on |z|^2 + |e|^2,
//   EPS_R = 2.4e-7 >= 2^-22 on |e|^2 alone, EPS_G = 6.0e-8 >= 2^-24.
// The bound is PER CODE (round 2 used the largest code norm of the whole codebook for every code: a trained EMA codebook holds a
// few large used codes next to thousands of unused ones that have decayed towards the origin; with the global norm all of those
// lie within one band of each other, every row overflowed its candidate list before it met its real neighbour and took the exact
// scan of the whole codebook -- 30.8 ms per launch at K = 8192 after ~100 training steps, profiles/r02_bench_final_1gpu.json cfg5).
// Code k can be the fp32 argmin only if  lo_k = d~_k - eps_k  <=  min_j (d~_j + eps_j) = ub.
// (fp16 overflow -- |z_i| or |e_i| > 65504 -- gives inf / NaN dot products: such a row keeps no candidate and takes the exact
// scan of the whole codebook.)
constexpr float EPS_A = 1.04e-3f, EPS_B = 2.4e-7f, EPS_R = 2.4e-7f, EPS_G = 6.0e-8f;

// The auxiliary buffer `cb_sq` of a codebook (4 K floats), written by the two preparation kernels below and read by the search:
//   [0, K)        |e_k|^2 with EXACTLY the summation of row_sqnorm_kernel (vq.cu): lane-strided fp32 FMA chains, then the xor
//                 butterfly -- the strict kernel and the exact re-rank read the same values;
//   [K, 2K)       pa_k = EPS_A |e_k|                                       } the code-dependent part of the error bound eps_k
//   [2K, 3K)      qa_k = (EPS_B + EPS_R) |e_k|^2 + EPS_G sqrt(D) |e_k|     } (see below): eps_k = |z| pa_k + qa_k + row term
//   [3K, 3K + 2 ceil(K/32))   the same pair for the LARGEST norm of every 32-code chunk (the scan's chunk-wide test); the second
//                 value carries a minus sign when the chunk mixes large and small norms.
// One block = 32 warps = the 32 codes of one chunk.
__device__ __forceinline__ void prep_aux(float* __restrict__ sq, int K, int D, int code, bool valid, float s, int warp, int lane) {
    __shared__ float en_s[32], en_lo_s[32];
    const float fin = (valid && s < INFINITY) ? s : 0.f;                   // non-finite norms: the search takes the exact scan anyway
    const float en = sqrtf(fin);
    const float gd = EPS_G * sqrtf((float)D);
    if (lane == 0) {
        en_s[warp] = en;
        en_lo_s[warp] = valid ? en : INFINITY;
        if (valid) { sq[code] = s; sq[K + code] = EPS_A * en; sq[2 * K + code] = fmaf((EPS_B + EPS_R) * en, en, gd * en); }
    }
    __syncthreads();
    if (warp == 0) {
        float mx = en_s[lane], mi = en_lo_s[lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            mi = fminf(mi, __shfl_xor_sync(0xffffffffu, mi, o));
        }
        // a NEGATIVE second coefficient marks a chunk of mixed norms (largest > 1.5 x smallest): there the scan evaluates the
        // per-code bounds even for a single candidate; in a homogeneous chunk the chunk-wide bound is at most 1.5 x too wide
        const float qc = fmaf((EPS_B + EPS_R) * mx, mx, gd * mx);
        if (lane == 0) { sq[3 * K + 2 * blockIdx.x] = EPS_A * mx; sq[3 * K + 2 * blockIdx.x + 1] = (mx > 1.5f * mi) ? -qc : qc; }
    }
}

// codebook fp32 [K][D] -> fp16 copy [K][D] and the auxiliary buffer
__global__ void __launch_bounds__(1024) vq_prep_codebook_kernel(const float* __restrict__ cb, __half* __restrict__ hf, float* __restrict__ sq, int K, int D) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * 32 + warp;
    const bool valid = row < K;
    float s = 0.f;
    if (valid)
        for (int d = lane; d < D; d += 32) {
            const float v = cb[row * D + d];
            hf[row * D + d] = __float2half_rn(v);
            s = fmaf(v, v, s);
        }
    s = warp_sum(s);
    prep_aux(sq, K, D, (int)row, valid, s, warp, lane);
}

// EMA state update (vector_quantizers.py:158-169, same arithmetic as vq_ema_update_kernel in vq.cu) that also leaves the
// fp16 copy and the auxiliary buffer of the NEW codebook for the next step's search: no separate preparation launch on the EMA path.
__global__ void __launch_bounds__(1024) vq_ema_update_prep_kernel(float* __restrict__ ema_count, float* __restrict__ ema_weight, float* __restrict__ cb,
                                                                  const float* __restrict__ counts, const float* __restrict__ dw, __half* __restrict__ hf,
                                                                  float* __restrict__ sq, int K, int D, float decay, float eps, float batch) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int code = blockIdx.x * 32 + warp;
    const bool valid = code < K;
    float s = 0.f, cnt = 0.f;
    if (valid) {
        const float c = ema_count[code] * decay + (1.0f - decay) * counts[code];
        cnt = (c + eps) / (batch + (float)K * eps) * batch;
        for (int d = lane; d < D; d += 32) {
            const int64_t o = (int64_t)code * D + d;
            const float w = ema_weight[o] * decay + (1.0f - decay) * dw[o];
            ema_weight[o] = w;
            const float v = w / cnt;
            cb[o] = v;
            hf[o] = __float2half_rn(v);
            s = fmaf(v, v, s);
        }
    }
    s = warp_sum(s);
    __syncwarp();
    if (valid && lane == 0) ema_count[code] = cnt;
    prep_aux(sq, K, D, code, valid, s, warp, lane);
}

struct FusedParams {
    int64_t N;
    int K, D, kchunks, ctiles, order;
    const float* z;              // [N][D] fp32
    const float* cb;             // [K][D] fp32
    const float* cb_sq;          // [K]
    float* q_out;                // [N][D] or NULL
    int64_t* idx_out;            // [N]
    double* sse;                 // [1] or NULL
    float* counts;               // [K] or NULL
    float* dw;                   // [K][D] or NULL
    int* undecided;              // [2] or NULL: rows that took the exact re-rank, rows among them that needed the full scan
    long long* trace;            // [ctas][8] or NULL: globaltimer stamps of the phase boundaries (tools/vq_phases.py)
};

__device__ __forceinline__ long long gtimer() { long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define VQ_TRACE(slot) do { if (p.trace && et == 0) p.trace[(size_t)blockIdx.x * 8 + (slot)] = gtimer(); } while (0)

__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "WAITC_%=:\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONEC_%=;\n\t"
        "bra WAITC_%=;\n\t"
        "DONEC_%=:\n\t}" ::"r"(ptx::smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// instruction descriptor of kind::f16 with FP16 A / B (format 0), fp32 accumulate, both operands K-major
__device__ __forceinline__ uint32_t umma_idesc_f16(int M, int N) { return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void epi_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}


// exact fp32 distance of one (row, code) pair with the strict kernel's arithmetic (vq.cu: vq_assign_tile): |z|^2 from four
// interleaved FMA chains combined as (s0+s1)+(s2+s3), the dot product as ONE sequential FMA chain over d, then the
// reference's operation order.  z_sq does not depend on the code; it is recomputed here so that every lane holds it.
__device__ __forceinline__ float exact_distance(const float* __restrict__ zrow, const float* __restrict__ erow, float e2, int D, int order) {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f, acc = 0.f;
    for (int d = 0; d < D; d += 4) {
        const float4 zv = *reinterpret_cast<const float4*>(zrow + d);
        const float4 ev = __ldg(reinterpret_cast<const float4*>(erow + d));
        s0 = __fmaf_rn(zv.x, zv.x, s0); s1 = __fmaf_rn(zv.y, zv.y, s1); s2 = __fmaf_rn(zv.z, zv.z, s2); s3 = __fmaf_rn(zv.w, zv.w, s3);
        acc = __fmaf_rn(zv.x, ev.x, acc); acc = __fmaf_rn(zv.y, ev.y, acc); acc = __fmaf_rn(zv.z, ev.z, acc); acc = __fmaf_rn(zv.w, ev.w, acc);
    }
    const float zsq = __fadd_rn(__fadd_rn(s0, s1), __fadd_rn(s2, s3));
    const float two_dot = __fmul_rn(2.0f, acc);
    return (order == 0) ? __fsub_rn(__fadd_rn(zsq, e2), two_dot) : __fadd_rn(__fsub_rn(zsq, two_dot), e2);
}

// Append (lo, code) to the candidate list of `slot` (shared memory, [CAP][256]); `cnt` = its length.  When the list is full it
// is first compacted against the current bound (entries whose lower bound rose above the running upper bound are dropped).  If
// CAP entries remain, the list keeps the CAP SMALLEST lower bounds and drop[slot] remembers the smallest one that was turned away:
// at the end the list is complete iff that value lies above the final bound (otherwise: more than CAP genuine near-ties --
// duplicated codes -- and the row takes the exact scan of the whole codebook).
// Deliberately NOT inlined: the scan loop is unrolled 32x and an inlined copy per element made the loop body ~100 KB of
// code (ncu: the scan warps stalled on instruction fetch); a thread gets here ~ln K times per row half.
__device__ __noinline__ int cand_push(float d, int code, float bound, int cnt, int slot, float* cand_d, uint16_t* cand_c, float* drop) {
    if (!(d < INFINITY)) return cnt;                                        // padding codes (|e|^2 = inf) and NaN never qualify
    int n = cnt;
    if (n == CAP) {
        int m = 0;
        for (int i = 0; i < CAP; ++i) {
            const float di = cand_d[i * 256 + slot];
            if (di <= bound) { cand_d[m * 256 + slot] = di; cand_c[m * 256 + slot] = cand_c[i * 256 + slot]; ++m; }
        }
        n = m;
    }
    if (n < CAP) { cand_d[n * 256 + slot] = d; cand_c[n * 256 + slot] = (uint16_t)code; return n + 1; }
    int im = 0;
    float dm = cand_d[slot];
    for (int i = 1; i < CAP; ++i) { const float di = cand_d[i * 256 + slot]; if (di > dm) { dm = di; im = i; } }
    float out = d;
    if (d < dm) { out = dm; cand_d[im * 256 + slot] = d; cand_c[im * 256 + slot] = (uint16_t)code; }
    drop[slot] = fminf(drop[slot], out);
    return n;
}

// the same arithmetic on rows staged in shared memory (plain loads)
__device__ __forceinline__ float exact_distance_smem(const float* zrow, const float* erow, float e2, int D, int order) {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f, acc = 0.f;
    for (int d = 0; d < D; d += 4) {
        const float4 zv = *reinterpret_cast<const float4*>(zrow + d);
        const float4 ev = *reinterpret_cast<const float4*>(erow + d);
        s0 = __fmaf_rn(zv.x, zv.x, s0); s1 = __fmaf_rn(zv.y, zv.y, s1); s2 = __fmaf_rn(zv.z, zv.z, s2); s3 = __fmaf_rn(zv.w, zv.w, s3);
        acc = __fmaf_rn(zv.x, ev.x, acc); acc = __fmaf_rn(zv.y, ev.y, acc); acc = __fmaf_rn(zv.z, ev.z, acc); acc = __fmaf_rn(zv.w, ev.w, acc);
    }
    const float zsq = __fadd_rn(__fadd_rn(s0, s1), __fadd_rn(s2, s3));
    const float two_dot = __fmul_rn(2.0f, acc);
    return (order == 0) ? __fsub_rn(__fadd_rn(zsq, e2), two_dot) : __fadd_rn(__fsub_rn(zsq, two_dot), e2);
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NTH, 1)
vq_fused_kernel(const __grid_constant__ CUtensorMap tmEh, const FusedParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* smemZ = smem;                                               // [kchunks] tiles of [128 rows][64 k] fp16
    uint8_t* smemE = smemZ + (size_t)p.kchunks * Z_TILE;                 // [ESTAGES] tiles of [128 codes][64 k] fp16
    float* cand_d = reinterpret_cast<float*>(smemE + (size_t)ESTAGES * E_TILE);      // [CAP][256]
    uint16_t* cand_c = reinterpret_cast<uint16_t*>(cand_d + CAP * 256);              // [CAP][256]
    float* e2_s = reinterpret_cast<float*>(cand_c + CAP * 256);                       // [2][256] code norms of the tile in flight
    float* pa_s = e2_s + 2 * TN;                                         // [2][256] EPS_A |e_k|                       } eps_k = |z| pa_k + qa_k
    float* qa_s = pa_s + 2 * TN;                                         // [2][256] EPS_B |e_k|^2 + EPS_G sqrt(D) |e_k| }         + row term
    float* ck_s = qa_s + 2 * TN;                                         // [2][8][2] the same pair for the largest norm of each 32-code chunk
    float* drop_s = ck_s + 32;                                           // [256] smallest lower bound turned away from a full list
    float* zsq_s = drop_s + 256;                                         // [128] |z|^2 (approximate order; bounds only)
    float* rbest = zsq_s + TM;                                           // [2][128] running upper bound of the minimum per (half, row)
    int* rcode = reinterpret_cast<int*>(rbest + 2 * TM);                 // [128] final code per row, [128] fp16-overflow flag per row
    int* rcnt = rcode + 2 * TM;                                          // [2][128] list length, bit 31 = take the exact full scan
    float* red_s = reinterpret_cast<float*>(rcnt + 2 * TM);              // [8] per-warp partial sums, [8] emax
    uint64_t* zfull = reinterpret_cast<uint64_t*>(red_s + 16);
    uint64_t* efull = zfull + 1;                                          // [ESTAGES] (the leader's are used)
    uint64_t* eempty = efull + ESTAGES;                                   // [ESTAGES]
    uint64_t* tfull = eempty + ESTAGES;                                   // [2]
    uint64_t* tempty = tfull + 2;                                         // [2] (the leader's are used)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = ptx::cluster_ctarank();
    const int64_t r0 = (int64_t)blockIdx.x * TM;                          // blockIdx.x = 2 * cluster + rank

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmEh);
        ptx::mbar_init(zfull, 2);                                          // one arrival per CTA of the pair
        for (int i = 0; i < ESTAGES; ++i) { ptx::mbar_init(&efull[i], 1); ptx::mbar_init(&eempty[i], 1); }
        for (int i = 0; i < 2; ++i) { ptx::mbar_init(&tfull[i], 1); ptx::mbar_init(&tempty[i], 16); }      // 8 scan warps x 2 CTAs
        ptx::fence_barrier_init();
    }
    if (warp == 1) ptx::tmem_alloc2(tmem_slot, 512);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::cluster_sync();                              // barriers and TMEM of BOTH CTAs exist before any remote signal
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ---- codebook stream: this CTA's 128-code half of every (tile j, k chunk c, hi | lo) stage ----------------------
        if (lane == 0) {
            int s = 0; uint32_t ph = 0;
            for (int j = 0; j < p.ctiles; ++j)
                for (int c = 0; c < p.kchunks; ++c) {
                    ptx::mbar_wait(&eempty[s], ph ^ 1);
                    if (rank == 0) ptx::mbar_expect_tx(&efull[s], 2u * (uint32_t)E_TILE);
                    ptx::tma_load_2d_2sm(smemE + (size_t)s * E_TILE, &tmEh, ptx::mapa_rank(ptx::smem_u32(&efull[s]), 0), c * TK,
                                         j * TN + (int)rank * (TN / 2));
                    if (++s == ESTAGES) { s = 0; ph ^= 1; }
                }
        }
    } else if (warp == 1) {
        // ---- MMA issue (leader CTA only): M = 256 rows (128 per CTA), N = 256 codes (128 staged per CTA) -----------------
        if (lane == 0 && rank == 0) {
            const uint32_t idesc = umma_idesc_f16(2 * TM, TN);
            mbar_wait_cluster(zfull, 0);
            ptx::tc_fence_after();
            int s = 0; uint32_t ph = 0;
            int as = 0; uint32_t aph = 0;
            for (int j = 0; j < p.ctiles; ++j) {
                ptx::mbar_wait(&tempty[as], aph ^ 1);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(as * TN);
                for (int c = 0; c < p.kchunks; ++c) {
                    ptx::mbar_wait(&efull[s], ph);
                    ptx::tc_fence_after();
                    const uint64_t bdesc = ptx::umma_smem_desc(ptx::smem_u32(smemE + (size_t)s * E_TILE), 0, 1024);
                    const uint64_t adesc = ptx::umma_smem_desc(ptx::smem_u32(smemZ + (size_t)c * Z_TILE), 0, 1024);
#pragma unroll
                    for (int k = 0; k < TK / 16; ++k)
                        ptx::umma2_bf16(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (c | k) != 0 ? 1u : 0u);
                    ptx::umma2_commit(&eempty[s]);
                    if (++s == ESTAGES) { s = 0; ph ^= 1; }
                }
                ptx::umma2_commit(&tfull[as]);
                if (++as == 2) { as = 0; aph ^= 1; }
            }
        }
    } else {
        const int ew = warp - 2;                                            // 0..7
        const int et = ew * 32 + lane;                                      // 0..255
        VQ_TRACE(0);
        // ---- prologue: z rows -> bf16 hi / lo in the UMMA K-major SWIZZLE_128B layout (resident), |z|^2 ------------------
        // one warp per row and iteration: lane L owns k = 8L .. 8L+7 (one 16-byte chunk of the hi and of the lo tile)
        {
            const int cidx = lane >> 3, j16 = lane & 7;
            for (int rr = 0; rr < TM / 8; rr += 8) {
                float4 v[8][2];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int row = ew * (TM / 8) + rr + u;
                    const int64_t g = r0 + row;
                    v[u][0] = v[u][1] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (g < p.N && lane * 8 < p.D) {
                        const float4* src = reinterpret_cast<const float4*>(p.z + g * p.D + lane * 8);
                        v[u][0] = src[0]; v[u][1] = src[1];
                    }
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int row = ew * (TM / 8) + rr + u;
                    const float f[8] = {v[u][0].x, v[u][0].y, v[u][0].z, v[u][0].w, v[u][1].x, v[u][1].y, v[u][1].z, v[u][1].w};
                    uint32_t hw[4];
                    float s = 0.f;
                    int bad = 0;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const __half2 h2 = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
                        hw[i] = *reinterpret_cast<const uint32_t*>(&h2);
                        s = fmaf(f[2 * i], f[2 * i], s); s = fmaf(f[2 * i + 1], f[2 * i + 1], s);
                        bad |= !(fabsf(f[2 * i]) < 65504.f) | !(fabsf(f[2 * i + 1]) < 65504.f);      // fp16 overflow, inf, NaN
                    }
                    if (lane * 8 < p.D) {
                        const uint32_t off = (uint32_t)row * 128u + (uint32_t)((j16 ^ (row & 7)) << 4);
                        *reinterpret_cast<uint4*>(smemZ + (size_t)cidx * Z_TILE + off) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                    }
                    s = warp_sum(s);
                    bad = __any_sync(0xffffffffu, bad != 0);
                    if (lane == 0) { zsq_s[row] = s; rcode[TM + row] = bad; }
                }
            }
            // largest code norm (threshold only): every CTA scans the K norms once (L2-resident, 4 KB at K = 1024)
            float m = 0.f;
            for (int k = et; k < p.K; k += 256) { const float v = p.cb_sq[k]; m = (v < 4.29e9f) ? fmaxf(m, v) : INFINITY; }   // inf: fp16 overflow / NaN
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
            if (lane == 0) red_s[8 + ew] = m;
            fence_proxy_async_smem();                                       // generic-proxy stores -> visible to the UMMA (async proxy)
            epi_sync();
            if (et == 0) ptx::mbar_arrive_cluster(ptx::mapa_rank(ptx::smem_u32(zfull), 0));
        }
        VQ_TRACE(1);                                                        // prologue done (z staged, |z|^2)
        const float sqrt_d = sqrtf((float)p.D);
        float emax = red_s[8];
#pragma unroll
        for (int i = 1; i < 8; ++i) emax = fmaxf(emax, red_s[8 + i]);

        // ---- scan: running minimum + online candidate list per (row, column half) -----------------------------------------
        const int quarter = warp & 3, half = ew >> 2;
        const int row = quarter * 32 + lane;
        const int slot = half * TM + row;
        const float gd = EPS_G * sqrt_d;
        const float zn = sqrtf(zsq_s[row]);
        const float rterm = fmaf(EPS_B, zsq_s[row], gd * zn);              // the part of eps that depends on the row only
        float ub = INFINITY;                                                // running min of the upper bounds d~ + eps
        int cnt = 0;                                                        // list length
        int as = 0; uint32_t aph = 0;
        drop_s[et] = INFINITY;
        // code norms of a tile and the coefficients of their error bounds (prepared with the codebook, see prep_aux) go from the
        // auxiliary buffer to shared memory by cp.async, 16 bytes per thread (64 + 64 + 64 granules of four codes, 4 of two chunk
        // pairs): nothing is held in registers across the scan of a tile (holding the four values of the next tile there spilled:
        // 46 K local loads per launch that missed the 28 KB L1 a third of the time, right before the per-tile barrier)
        const int nchunk2 = 2 * ((p.K + 31) / 32);
        auto fetch_norms = [&](const int tile, const int buf) {
            if (et < 196) {
                const int which = et >> 6, gq = (et & 63) * 4;
                float* dst; const float* src; bool in;
                if (which < 3) {
                    const int code = tile * TN + gq;
                    in = tile < p.ctiles && code < p.K;                     // K % 8 == 0: a granule is entirely in or out
                    dst = (which == 0 ? e2_s : which == 1 ? pa_s : qa_s) + buf * TN + gq;
                    src = p.cb_sq + (size_t)which * p.K + code;
                } else {
                    in = tile < p.ctiles && tile * 16 + gq < nchunk2;
                    dst = ck_s + buf * 16 + gq;
                    src = p.cb_sq + (size_t)3 * p.K + tile * 16 + gq;
                }
                if (in) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(ptx::smem_u32(dst)), "l"(src) : "memory");
                else {
                    const float f = (which == 0) ? INFINITY : 0.f;          // padding codes never qualify (d~ = inf)
                    *reinterpret_cast<float4*>(dst) = make_float4(f, f, f, f);
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        fetch_norms(0, 0);                                                  // tile 0
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        for (int j = 0; j < p.ctiles; ++j) {
            epi_sync();                                                     // e2_s[as] written; every warp has left tile j-1
            // norms of the NEXT tile: loaded now, stored after this tile's scan (buffer as^1 was tile j-1's, free since the sync)
            fetch_norms(j + 1, as ^ 1);                                     // buffer as^1 was tile j-1's: free since the sync
            ptx::mbar_wait(&tfull[as], aph);
            ptx::tc_fence_after();
            const uint32_t t_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * TN);
            // The accumulator chunks are read one AHEAD: tcgen05.ld of chunk i+1 is issued before chunk i is scanned, so the TMEM read
            // latency overlaps the ~100 instructions of a chunk (read -> wait -> scan in sequence left the eight scan warps, two per
            // scheduler, latency-bound: 3.2 us per 256-code tile against 1.1 us of tensor work -- tools/vq_phases.py).
            auto scan_chunk = [&](const uint32_t (&r)[32], const int c) {
                const float* e2c = e2_s + as * TN + c;
                // all 32 distances first (independent FMAs) and their minimum by a tree.  Only if the chunk minimum, lowered by the
                // error bound of the chunk's LARGEST code norm, reaches the running upper bound can the chunk hold candidates at
                // all; only then are the per-code bounds evaluated: lo_k = d~_k - eps_k, the upper bound tightened by
                // min_k (d~_k + eps_k), and every code with lo_k <= ub appended (usually one).  ub only decreases, so a code that
                // is not appended here lies above the final bound as well.
                float d[32];
#pragma unroll
                for (int u = 0; u < 32; u += 4) {
                    const float4 e4 = *reinterpret_cast<const float4*>(e2c + u);
                    d[u] = fmaf(-2.0f, __uint_as_float(r[u]), e4.x); d[u + 1] = fmaf(-2.0f, __uint_as_float(r[u + 1]), e4.y);
                    d[u + 2] = fmaf(-2.0f, __uint_as_float(r[u + 2]), e4.z); d[u + 3] = fmaf(-2.0f, __uint_as_float(r[u + 3]), e4.w);
                }
                float m[16];
#pragma unroll
                for (int u = 0; u < 16; ++u) m[u] = fminf(d[u], d[u + 16]);
#pragma unroll
                for (int w2 = 8; w2 > 0; w2 >>= 1)
#pragma unroll
                    for (int u = 0; u < w2; ++u) m[u] = fminf(m[u], m[u + w2]);
                const float mn = m[0];
                const float2 ck = *reinterpret_cast<const float2*>(ck_s + (as * 8 + (c >> 5)) * 2);
                const float epsc = fmaf(zn, ck.x, fabsf(ck.y)) + rterm;
                if (mn - epsc <= ub && mn < INFINITY) {
                    // (at warp level this branch is the COMMON case -- some lane of 32 improves its running bound in most chunks --
                    // so its usual form stays as light as the comparison against one band: the codes that pass the chunk-wide bound
                    // first, the per-code bound only for those)
                    ub = fminf(ub, mn + epsc);                              // the chunk minimum's own upper bound (eps_k <= epsc)
                    const float band = ub + epsc;                           // = min(old ub, mn + epsc) + epsc: at most mn + 2 epsc
                    unsigned near = 0;
#pragma unroll
                    for (int u = 0; u < 32; ++u) near |= (d[u] <= band) ? (1u << u) : 0u;
                    const float* pac = pa_s + as * TN + c;
                    const float* qac = qa_s + as * TN + c;
                    const int nn = __popc(near);
                    if (nn == 1) {                                            // the usual case: the chunk minimum alone
                        const int u = __ffs(near) - 1;
                        float lo = mn - epsc;                                 // homogeneous chunk: the chunk-wide bound is the code's
                        if (ck.y < 0.f) {                                     // mixed norms (warp-uniform): the code's own bound
                            const float eu = fmaf(zn, pac[u], qac[u]) + rterm;
                            ub = fminf(ub, mn + eu);
                            lo = mn - eu;
                        }
                        if (lo <= ub) {
                            const int code = j * TN + c + u;
                            if (cnt < CAP) { cand_d[cnt * 256 + slot] = lo; cand_c[cnt * 256 + slot] = (uint16_t)code; ++cnt; }
                            else cnt = cand_push(lo, code, ub, cnt, slot, cand_d, cand_c, drop_s);
                        }
                    } else if (nn <= 4) {
#pragma unroll 1
                        while (near) {
                            const int u = __ffs(near) - 1;
                            near &= near - 1;
                            // the value of element u without a dynamically indexed register array: a select chain
                            float du = d[0];
#pragma unroll
                            for (int t = 1; t < 32; ++t) du = (u == t) ? d[t] : du;
                            const float eu = fmaf(zn, pac[u], qac[u]) + rterm;
                            ub = fminf(ub, du + eu);
                            if (du - eu <= ub) cnt = cand_push(du - eu, j * TN + c + u, ub, cnt, slot, cand_d, cand_c, drop_s);
                        }
                    } else {
                        // many codes inside the chunk-wide band (a chunk that mixes one large-norm code with decayed ones): all
                        // 32 per-code bounds at once, then the few that remain
                        float hmin = INFINITY;
#pragma unroll
                        for (int u = 0; u < 32; u += 4) {
                            const float4 p4 = *reinterpret_cast<const float4*>(pac + u);
                            const float4 q4 = *reinterpret_cast<const float4*>(qac + u);
                            const float e0 = fmaf(zn, p4.x, q4.x) + rterm, e1 = fmaf(zn, p4.y, q4.y) + rterm;
                            const float e2 = fmaf(zn, p4.z, q4.z) + rterm, e3 = fmaf(zn, p4.w, q4.w) + rterm;
                            hmin = fminf(fminf(hmin, d[u] + e0), fminf(d[u + 1] + e1, fminf(d[u + 2] + e2, d[u + 3] + e3)));
                            d[u] -= e0; d[u + 1] -= e1; d[u + 2] -= e2; d[u + 3] -= e3;
                        }
                        ub = fminf(ub, hmin);
                        near = 0;
#pragma unroll
                        for (int u = 0; u < 32; ++u) near |= (d[u] <= ub) ? (1u << u) : 0u;
#pragma unroll 1
                        while (near) {
                            const int u = __ffs(near) - 1;
                            near &= near - 1;
                            float du = d[0];
#pragma unroll
                            for (int t = 1; t < 32; ++t) du = (u == t) ? d[t] : du;
                            cnt = cand_push(du, j * TN + c + u, ub, cnt, slot, cand_d, cand_c, drop_s);
                        }
                    }
                }
            };
            {
                uint32_t ra[32], rb[32];
                const int c0 = half * 32;
                ptx::tmem_ld32(t_addr + (uint32_t)c0, ra);
#pragma unroll 1
                for (int it = 0; it < TN / 64; it += 2) {
                    ptx::tmem_ld_wait();
                    ptx::tmem_ld32(t_addr + (uint32_t)(c0 + (it + 1) * 64), rb);
                    scan_chunk(ra, c0 + it * 64);
                    ptx::tmem_ld_wait();
                    if (it + 2 < TN / 64) ptx::tmem_ld32(t_addr + (uint32_t)(c0 + (it + 2) * 64), ra);
                    scan_chunk(rb, c0 + (it + 1) * 64);
                }
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive_cluster(ptx::mapa_rank(ptx::smem_u32(&tempty[as]), 0));
            if (++as == 2) { as = 0; aph ^= 1; }
            asm volatile("cp.async.wait_group 0;" ::: "memory");            // the next tile's norms have landed (visible after the sync)
        }
        // rows (or codebooks) that do not survive the fp16 rounding take the exact scan of the whole codebook
        rbest[slot] = ub; rcnt[slot] = (rcode[TM + row] != 0 || !(emax < INFINITY)) ? (int)0x80000000 : cnt;
        epi_sync();
        VQ_TRACE(2);                                                        // scan of every code tile done

        // ---- decide (one thread per row): a single surviving candidate is the fp32 argmin; the other rows are appended to the
        //      CTA's re-rank work list (balanced over the eight warps afterwards)
        int* und_n = reinterpret_cast<int*>(tmem_slot + 1);               // spare word next to the TMEM address slot
        int* und_rows = rcode + TM;                                         // the fp16-overflow flags kept there were consumed above
        if (et == 0) *und_n = 0;
        epi_sync();
        if (et < TM) {
            const float bound = fminf(rbest[et], rbest[TM + et]);
            int my_ca = rcnt[et], my_cb = rcnt[TM + et];
            if (drop_s[et] <= bound || drop_s[TM + et] <= bound) { my_ca |= (int)0x80000000; rcnt[et] = my_ca; }   // the lists are incomplete
            int keep = 0, code = -1;
            if (((my_ca | my_cb) >> 31) == 0) {
                for (int i = 0; i < my_ca; ++i) if (cand_d[i * 256 + et] <= bound) { ++keep; code = cand_c[i * 256 + et]; }
                for (int i = 0; i < my_cb; ++i) if (cand_d[i * 256 + TM + et] <= bound) { ++keep; code = cand_c[i * 256 + TM + et]; }
            }
            const int64_t g = r0 + et;
            if (g < p.N) {
                if (keep == 1) { rcode[et] = code; p.idx_out[g] = (int64_t)code; }
                else { rcode[et] = -1; und_rows[atomicAdd(und_n, 1)] = et; }
            } else rcode[et] = 0;
        }
        epi_sync();

        // ---- exact re-rank of the near-tied rows (one warp per row, round-robin over the work list): ONLY the surviving
        //      candidates are evaluated -------------------------------------------------------------------------------------
        VQ_TRACE(3);                                                        // decided rows written, re-rank work list built
        float* xbuf = reinterpret_cast<float*>(smemZ);          // staging: every MMA has completed (last tfull), the operand tiles are free
        float sse_local = 0.f;
        const int n_und = *und_n;
        for (int wi = ew; wi < n_und; wi += 8) {
            const int rw = und_rows[wi];
            const int64_t g = r0 + rw;
            const float* zrow = p.z + g * p.D;
            const int ca = rcnt[rw], cb2 = rcnt[TM + rw];
            float dist = INFINITY; int c2 = 0x7fffffff;
            bool have = false;
            if (((ca | cb2) >> 31) == 0) {
                const float bound = fminf(rbest[rw], rbest[TM + rw]);
              for (int cbase = 0; cbase < ca + cb2; cbase += 32) {       // the concatenated list of both halves, 32 entries per pass
                float dme = INFINITY; int cme = 0x7fffffff;
                const int li = cbase + lane;
                if (li < ca) { dme = cand_d[li * 256 + rw]; cme = cand_c[li * 256 + rw]; }
                else if (li < ca + cb2) { dme = cand_d[(li - ca) * 256 + TM + rw]; cme = cand_c[(li - ca) * 256 + TM + rw]; }
                const bool keep = dme <= bound;
                unsigned kmask = __ballot_sync(0xffffffffu, keep);
                have = have || kmask != 0;
                // The fp32 dot product is ONE sequential FMA chain (bit-exactness with the strict kernel), so a lane walking global
                // memory pays an L2 round trip every few steps (ncu: 5 us per row).  Instead the warp stages the z row and up to
                // four candidate rows in shared memory with coalesced loads issued back to back (the operand tiles are dead by
                // now), and one lane per candidate runs its chain from there.
                float* xz = xbuf + ew * (5 * 256);
                while (kmask) {
                    int codes4[4]; int ng = 0;
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) {
                        codes4[cc] = -1;
                        if (kmask) { codes4[cc] = __shfl_sync(0xffffffffu, cme, __ffs(kmask) - 1); kmask &= kmask - 1; ++ng; }
                    }
                    float4 zr[2], er[4][2];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int d = lane * 4 + h * 128;
                        if (d < p.D) {
                            zr[h] = *reinterpret_cast<const float4*>(zrow + d);
#pragma unroll
                            for (int cc = 0; cc < 4; ++cc)
                                if (codes4[cc] >= 0) er[cc][h] = __ldg(reinterpret_cast<const float4*>(p.cb + (int64_t)codes4[cc] * p.D + d));
                        }
                    }
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int d = lane * 4 + h * 128;
                        if (d < p.D) {
                            *reinterpret_cast<float4*>(xz + d) = zr[h];
#pragma unroll
                            for (int cc = 0; cc < 4; ++cc)
                                if (codes4[cc] >= 0) *reinterpret_cast<float4*>(xz + 256 * (1 + cc) + d) = er[cc][h];
                        }
                    }
                    __syncwarp();
                    if (lane < ng) {
                        int code_l = codes4[0];
#pragma unroll
                        for (int cc = 1; cc < 4; ++cc) code_l = (lane == cc) ? codes4[cc] : code_l;
                        const float dl = exact_distance_smem(xz, xz + 256 * (1 + lane), p.cb_sq[code_l], p.D, p.order);
                        if (dl < dist || (dl == dist && code_l < c2)) { dist = dl; c2 = code_l; }
                    }
                    __syncwarp();
                }
              }
            }
            if (!have) {
                // list overflow (more near-ties than it holds: duplicated codes) or no finite approximate distance at all
                // (fp16 overflow, NaN): exact scan of every code
                if (lane == 0 && p.undecided) atomicAdd(p.undecided + 1, 1);
                for (int k = lane; k < p.K; k += 32) {
                    const float dk = exact_distance(zrow, p.cb + (int64_t)k * p.D, p.cb_sq[k], p.D, p.order);
                    if (dk < dist || (dk == dist && k < c2)) { dist = dk; c2 = k; }
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float od = __shfl_xor_sync(0xffffffffu, dist, o);
                const int oc = __shfl_xor_sync(0xffffffffu, c2, o);
                if (od < dist || (od == dist && oc < c2)) { dist = od; c2 = oc; }
            }
            int code = c2;
            if (code < 0 || code >= p.K) code = 0;        // NaN rows: keep memory-safe (torch.argmin would return the NaN position)
            if (lane == 0) { rcode[rw] = code; p.idx_out[g] = (int64_t)code; }
        }
        epi_sync();                                                         // every row of the CTA has its code
        VQ_TRACE(4);                                                        // exact re-rank done

        // ---- finish, 8 rows in flight per warp: q = z + (e - z), sum (e - z)^2, EMA cluster sums (z re-read from L2) ----------
        for (int rr = 0; rr < TM / 8; rr += 8) {
            float4 zv[8][2], ev[8][2];
            int codes[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int rw = ew * (TM / 8) + rr + u;
                const int64_t g = r0 + rw;
                codes[u] = (g < p.N) ? rcode[rw] : -1;
                if (lane == u && codes[u] >= 0 && p.counts) atomicAdd(p.counts + codes[u], 1.0f);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int d = lane * 4 + h * 128;
                    zv[u][h] = ev[u][h] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (codes[u] >= 0 && d < p.D) {
                        zv[u][h] = *reinterpret_cast<const float4*>(p.z + g * p.D + d);
                        ev[u][h] = __ldg(reinterpret_cast<const float4*>(p.cb + (int64_t)codes[u] * p.D + d));
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int64_t g = r0 + ew * (TM / 8) + rr + u;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int d = lane * 4 + h * 128;
                    if (codes[u] >= 0 && d < p.D) {
                        const float4 z4 = zv[u][h], e4 = ev[u][h];
                        const float4 df = make_float4(e4.x - z4.x, e4.y - z4.y, e4.z - z4.z, e4.w - z4.w);
                        sse_local = fmaf(df.x, df.x, sse_local); sse_local = fmaf(df.y, df.y, sse_local);
                        sse_local = fmaf(df.z, df.z, sse_local); sse_local = fmaf(df.w, df.w, sse_local);
                        if (p.q_out)                                          // flat_x + (quantized - flat_x).detach()
                            *reinterpret_cast<float4*>(p.q_out + g * p.D + d) = make_float4(z4.x + df.x, z4.y + df.y, z4.z + df.z, z4.w + df.w);
                        if (p.dw) red_add_v4(p.dw + (int64_t)codes[u] * p.D + d, z4.x, z4.y, z4.z, z4.w);
                    }
                }
            }
        }
        sse_local = warp_sum(sse_local);
        if (lane == 0) red_s[ew] = sse_local;
        epi_sync();
        if (et == 0 && p.sse) {
            double t = 0.0;
            for (int i = 0; i < 8; ++i) t += (double)red_s[i];
            atomicAdd(p.sse, t);
        }
        if (et == 0 && p.undecided && n_und) atomicAdd(p.undecided, n_und);
        VQ_TRACE(5);                                                        // finish (gather / STE / statistics) done
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::cluster_sync();                              // the peer may still be reading this CTA's shared memory / TMEM until here
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc2(tmem_base, 512);
    }
}

size_t fused_smem_bytes(int D) {
    const int kchunks = D / TK;
    return (size_t)kchunks * Z_TILE + (size_t)ESTAGES * E_TILE + (size_t)CAP * 256 * 6 + (size_t)3 * 2 * TN * 4 + (32 + 256) * 4 + TM * 4 + 2 * TM * 4 * 3 +
           16 * 4 + (1 + 2 * ESTAGES + 4) * 8 + 16 + 1024 + 64;
}

}  // namespace

extern "C" int vqb_vq_prep_codebook(const float* codebook, void* cb_half, float* cb_sq, int K, int D, void* stream) {
    VQB_CHECK_ARG(codebook && cb_half && cb_sq && K > 0 && D > 0, "vq_prep_codebook: bad arguments");
    vq_prep_codebook_kernel<<<(unsigned)((K + 31) / 32), 1024, 0, as_stream(stream)>>>(codebook, (__half*)cb_half, cb_sq, K, D);
    VQB_CHECK_LAUNCH("vq_prep_codebook");
    return VQB_OK;
}

extern "C" int vqb_vq_ema_update_prep(float* ema_count, float* ema_weight, float* codebook, const float* counts, const float* dw,
                                      void* cb_half, float* cb_sq, int K, int D, float decay, float eps, float batch, void* stream) {
    VQB_CHECK_ARG(ema_count && ema_weight && codebook && counts && dw && cb_half && cb_sq && K > 0 && D > 0,
                  "vq_ema_update_prep: bad arguments");
    vq_ema_update_prep_kernel<<<(unsigned)((K + 31) / 32), 1024, 0, as_stream(stream)>>>(
        ema_count, ema_weight, codebook, counts, dw, (__half*)cb_half, cb_sq, K, D, decay, eps, batch);
    VQB_CHECK_LAUNCH("vq_ema_update_prep");
    return VQB_OK;
}

static long long* g_vq_trace = nullptr;
// tuning hook (tools/vq_phases.py): device buffer of [ctas][8] int64 that the next vqb_vq_fused launches fill with globaltimer
// stamps of their phase boundaries; NULL (default) switches the stamps off
extern "C" void vqb_vq_fused_set_trace(void* dev_buf) { g_vq_trace = (long long*)dev_buf; }

extern "C" int vqb_vq_fused(const float* z, const float* codebook, const void* cb_half, const float* cb_sq, int order, float* q_out,
                            int64_t* idx_out, double* sse, float* counts, float* dw, int64_t N, int K, int D, int* undecided_rows_out,
                            void* stream) {
    VQB_CHECK_ARG(z && codebook && cb_half && cb_sq && idx_out, "vq_fused: null pointer");
    VQB_CHECK_ARG(N > 0 && K > 0 && D > 0 && (order == 0 || order == 1), "vq_fused: bad arguments");
    VQB_CHECK_ARG(D % 64 == 0 && D <= 256 && K % 8 == 0 && K <= 65528, "vq_fused: needs D %% 64 == 0, D <= 256, K %% 8 == 0, K <= 65528 (got D=%d K=%d)", D, K);
    VQB_CHECK_ARG(N < (int64_t)1 << 31, "vq_fused: N too large");
    cudaStream_t st = as_stream(stream);
    CUtensorMap tmEh;
    int rc;
    if ((rc = make_code_map(&tmEh, cb_half, K, D, TN / 2))) return rc;
    FusedParams fp;
    fp.N = N; fp.K = K; fp.D = D; fp.kchunks = D / TK; fp.ctiles = (K + TN - 1) / TN; fp.order = order;
    fp.z = z; fp.cb = codebook; fp.cb_sq = cb_sq; fp.q_out = q_out; fp.idx_out = idx_out; fp.sse = sse; fp.counts = counts; fp.dw = dw;
    fp.undecided = undecided_rows_out;
    fp.trace = g_vq_trace;
    const size_t smem = fused_smem_bytes(D);
    if (smem > 227 * 1024) { vqb_set_error("vq_fused: D=%d does not fit the shared-memory budget (%zu B)", D, smem); return VQB_ERR_UNSUPPORTED; }
    static std::once_flag attr_once;
    std::call_once(attr_once, [] { cudaFuncSetAttribute(vq_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); });
    const unsigned ctas = (unsigned)(2 * ceil_div64(N, 2 * TM));             // whole clusters (a cluster covers 256 rows)
    vq_fused_kernel<<<ctas, NTH, smem, st>>>(tmEh, fp);
    VQB_CHECK_LAUNCH("vq_fused");
    return VQB_OK;
}
