// Nearest-code search on the tensor cores with an exactness guarantee (reference: the distance GEMM + argmin of
// vqvae/modules/vector_quantizers.py:37-49, 142-154, 337-343).
//
// Phase A (vq_tc_search_kernel, tcgen05): every latent z and code e is split into two bf16 terms (hi + lo, 16 mantissa
//   bits together); dot = zh.eh + zl.eh + zh.el accumulates in fp32 in TMEM (3 UMMAs per k-step).  The epilogue scans the
//   accumulator columns and keeps, per row, the smallest and second-smallest APPROXIMATE distance |e|^2 - 2 dot.
// Phase B (vq_tc_decide_kernel): a row whose gap (second - best) exceeds a rigorous bound on the approximation error is
//   decided -- its fp32 argmin is provably the same code.  The others (true near-ties, e.g. the reference's U(+-1/K)
//   initial codebook where thousands of codes are within 1e-5 of each other) are compacted into a list ...
// Phase C ... and re-evaluated by the exact fp32 kernel (vq_assign_rows_kernel in vq.cu), with the reference's operation
//   order and first-index tie-break, so the final indices are those of the exact kernel in every case.
// Phase D (vq_finish_kernel in vq.cu): gather, straight-through value, sum (e-z)^2, histogram, EMA cluster sums.
//
// Roofline: phase A is tensor-bound (3 x 2NKD FLOP), phases B-D are HBM-bound (z read twice, q written once).
#include "common.cuh"
#include "ptx.cuh"
#include <mutex>

int vqb_vq_exact_rows(const float* z, const float* codebook, const float* cb_sq, int order, const int* row_list,
                      const int* n_rows_dev, int64_t* idx_out, int64_t N, int K, int D, cudaStream_t st);
int vqb_vq_finish(const float* z, const float* codebook, const int64_t* idx, float* q_out, double* sse, float* counts, float* dw,
                  int64_t N, int K, int D, cudaStream_t st);

namespace {

constexpr int TM = 128;          // latent rows per CTA
constexpr int TN = 256;          // codes per accumulator tile
constexpr int TK = 64;           // k elements per smem tile (128-byte rows)
constexpr int NTH = 320;         // TMA, MMA, 8 epilogue warps

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode2() {
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

int make_2d_map(CUtensorMap* m, const void* base, int64_t rows, int cols, int box_rows) {
    EncodeTiledFn enc = get_encode2();
    if (!enc) { vqb_set_error("cuTensorMapEncodeTiled unavailable"); return VQB_ERR_CUDA; }
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { vqb_set_error("cuTensorMapEncodeTiled(vq) failed: %d", (int)r); return VQB_ERR_CUDA; }
    return VQB_OK;
}

// x fp32 [R][D] -> hi, lo bf16 [R][D] (x ~ hi + lo), sq[r] = |x_r|^2 (fp32, same summation as the exact kernel's norm)
__global__ void split_bf16_kernel(const float* __restrict__ x, bf16* __restrict__ hi, bf16* __restrict__ lo, float* __restrict__ sq,
                                  int64_t R, int D) {
    int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (row >= R) return;
    float s = 0.f;
    for (int d = lane; d < D; d += 32) {
        float v = x[row * D + d];
        bf16 h = __float2bfloat16_rn(v);
        bf16 l = __float2bfloat16_rn(v - __bfloat162float(h));
        hi[row * D + d] = h; lo[row * D + d] = l;
        s = fmaf(v, v, s);
    }
    s = warp_sum(s);
    if (lane == 0 && sq) sq[row] = s;
}

struct SearchParams {
    int64_t N;
    int K, D, kchunks, ctiles;
    const float* cb_sq;          // [K]
    int* best_idx;               // [N]
    float* best_val;             // [N]  approximate  |e|^2 - 2 dot  of the best code
    float* second_val;           // [N]  ... of the runner-up
};

__global__ void __launch_bounds__(NTH, 1)
vq_tc_search_kernel(const __grid_constant__ CUtensorMap tmZh, const __grid_constant__ CUtensorMap tmZl,
                    const __grid_constant__ CUtensorMap tmEh, const __grid_constant__ CUtensorMap tmEl, const SearchParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int z_tile = TM * TK * 2;                       // 16 KB: [128 rows][64 k] bf16
    const int e_tile = TN * TK * 2;                       // 32 KB: [256 codes][64 k] bf16
    constexpr int ESTAGES = 2;
    uint8_t* smemZ = smem;                                 // [2 (hi,lo)][kchunks] z tiles, resident for the whole CTA
    uint8_t* smemE = smemZ + (size_t)2 * p.kchunks * z_tile;
    float* e2s = reinterpret_cast<float*>(smemE + (size_t)ESTAGES * e_tile);         // [K] code norms
    uint64_t* zfull = reinterpret_cast<uint64_t*>(e2s + p.K);
    uint64_t* efull = zfull + 1;
    uint64_t* eempty = efull + ESTAGES;
    uint64_t* tfull = eempty + ESTAGES;                    // [2]
    uint64_t* tempty = tfull + 2;                          // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
    float* merge = reinterpret_cast<float*>(tmem_slot + 4);                           // [128 rows][3] scratch of the upper half

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t r0 = (int64_t)blockIdx.x * TM;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmZh); ptx::prefetch_tmap(&tmZl); ptx::prefetch_tmap(&tmEh); ptx::prefetch_tmap(&tmEl);
        ptx::mbar_init(zfull, 1);
        for (int i = 0; i < ESTAGES; ++i) { ptx::mbar_init(&efull[i], 1); ptx::mbar_init(&eempty[i], 1); }
        for (int i = 0; i < 2; ++i) { ptx::mbar_init(&tfull[i], 1); ptx::mbar_init(&tempty[i], 8); }
        ptx::fence_barrier_init();
    }
    if (warp == 1) ptx::tmem_alloc(tmem_slot, 512);
    for (int k = threadIdx.x; k < p.K; k += NTH) e2s[k] = p.cb_sq[k];
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            ptx::mbar_expect_tx(zfull, (uint32_t)(2 * p.kchunks * z_tile));
            for (int c = 0; c < p.kchunks; ++c) {
                ptx::tma_load_2d(smemZ + (size_t)c * z_tile, &tmZh, zfull, c * TK, (int)r0);
                ptx::tma_load_2d(smemZ + (size_t)(p.kchunks + c) * z_tile, &tmZl, zfull, c * TK, (int)r0);
            }
            int s = 0; uint32_t ph = 0;
            for (int j = 0; j < p.ctiles; ++j)
                for (int c = 0; c < p.kchunks; ++c)
                    for (int hl = 0; hl < 2; ++hl) {
                        ptx::mbar_wait(&eempty[s], ph ^ 1);
                        ptx::mbar_expect_tx(&efull[s], (uint32_t)e_tile);
                        ptx::tma_load_2d(smemE + (size_t)s * e_tile, hl ? &tmEl : &tmEh, &efull[s], c * TK, j * TN);
                        if (++s == ESTAGES) { s = 0; ph ^= 1; }
                    }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = ptx::umma_idesc_bf16(TM, TN, 0, 0);
            ptx::mbar_wait(zfull, 0);
            int s = 0; uint32_t ph = 0;
            int as = 0; uint32_t aph = 0;
            for (int j = 0; j < p.ctiles; ++j) {
                ptx::mbar_wait(&tempty[as], aph ^ 1);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(as * TN);
                uint32_t first = 1;
                for (int c = 0; c < p.kchunks; ++c)
                    for (int hl = 0; hl < 2; ++hl) {
                        ptx::mbar_wait(&efull[s], ph);
                        ptx::tc_fence_after();
                        const uint64_t bdesc = ptx::umma_smem_desc(ptx::smem_u32(smemE + (size_t)s * e_tile), 0, 1024);
                        const uint64_t zh = ptx::umma_smem_desc(ptx::smem_u32(smemZ + (size_t)c * z_tile), 0, 1024);
                        const uint64_t zl = ptx::umma_smem_desc(ptx::smem_u32(smemZ + (size_t)(p.kchunks + c) * z_tile), 0, 1024);
#pragma unroll
                        for (int k = 0; k < TK / 16; ++k) {
                            ptx::umma_bf16(d_tmem, zh + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, first ? 0u : 1u);   // zh.eh / zh.el
                            first = 0;
                            if (hl == 0) ptx::umma_bf16(d_tmem, zl + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, 1u);      // zl.eh
                        }
                        ptx::umma_commit(&eempty[s]);
                        if (++s == ESTAGES) { s = 0; ph ^= 1; }
                    }
                ptx::umma_commit(&tfull[as]);
                if (++as == 2) { as = 0; aph ^= 1; }
            }
        }
    } else {
        const int quarter = warp & 3, half = (warp - 2) >> 2;
        const int row = quarter * 32 + lane;
        float b1 = INFINITY, b2 = INFINITY; int bi = 0x7fffffff;
        int as = 0; uint32_t aph = 0;
        for (int j = 0; j < p.ctiles; ++j) {
            ptx::mbar_wait(&tfull[as], aph);
            ptx::tc_fence_after();
            const uint32_t t_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * TN);
            for (int c = half * 32; c < TN; c += 64) {
                uint32_t r[32];
                ptx::tmem_ld32(t_addr + (uint32_t)c, r);
                ptx::tmem_ld_wait();
#pragma unroll
                for (int u = 0; u < 32; ++u) {
                    const int code = j * TN + c + u;
                    if (code < p.K) {
                        const float d = fmaf(-2.0f, __uint_as_float(r[u]), e2s[code]);
                        if (d < b1 || (d == b1 && code < bi)) { b2 = b1; b1 = d; bi = code; }
                        else if (d < b2) b2 = d;
                    }
                }
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&tempty[as]);
            if (++as == 2) { as = 0; aph ^= 1; }
        }
        // merge the two column halves of each row (named barrier over the 256 epilogue threads)
        if (half == 1) { merge[row * 3 + 0] = b1; merge[row * 3 + 1] = b2; merge[row * 3 + 2] = __int_as_float(bi); }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (half == 0) {
            const float o1 = merge[row * 3 + 0], o2 = merge[row * 3 + 1]; const int oi = __float_as_int(merge[row * 3 + 2]);
            float n1, n2; int ni;
            if (o1 < b1 || (o1 == b1 && oi < bi)) { n1 = o1; ni = oi; n2 = fminf(b1, o2); }
            else { n1 = b1; ni = bi; n2 = fminf(o1, b2); }
            const int64_t g = r0 + row;
            if (g < p.N) { p.best_idx[g] = ni; p.best_val[g] = n1; p.second_val[g] = n2; }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, 512);
    }
}

// decided rows get their index; the others are appended to row_list (order irrelevant)
__global__ void vq_tc_decide_kernel(const int* __restrict__ best_idx, const float* __restrict__ best_val, const float* __restrict__ second_val,
                                    const float* __restrict__ z_sq, const float* __restrict__ emax_sq, int64_t* __restrict__ idx_out,
                                    int* __restrict__ row_list, int* __restrict__ n_rows, int64_t N) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= N) return;
    const float zs = z_sq[i], es = emax_sq[0];
    // |approx - fp32| <= 2 * (3 * 2^-18 + 2^-16) |z||e| on the distance; threshold with a 4x safety factor, plus the fp32
    // evaluation's own rounding band so that genuine fp32 near-ties always take the exact path
    const float thr = 4.8828125e-4f * sqrtf(zs * es) + 1e-5f * (zs + es);
    if (second_val[i] - best_val[i] > thr) idx_out[i] = (int64_t)best_idx[i];
    else { int slot = atomicAdd(n_rows, 1); row_list[slot] = (int)i; }
}

__global__ void max_kernel(const float* __restrict__ a, float* __restrict__ out, int n) {
    __shared__ float sh[32];
    float m = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) m = fmaxf(m, a[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) { for (int i = 1; i < (int)(blockDim.x >> 5); ++i) m = fmaxf(m, sh[i]); out[0] = m; }
}

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace

extern "C" size_t vqb_vq_tc_workspace_bytes(int64_t N, int K, int D) {
    size_t b = 0;
    b += align_up((size_t)N * D * 2, 256) * 2;          // z hi / lo
    b += align_up((size_t)K * D * 2, 256) * 2;          // e hi / lo
    b += align_up((size_t)N * 4, 256) * 5;              // z_sq, best_idx, best_val, second_val, row_list
    b += align_up((size_t)K * 4, 256);                  // cb_sq
    b += 256;                                            // emax_sq, n_rows
    return b;
}

extern "C" int vqb_vq_assign_tc(const float* z, const float* codebook, int order, float* q_out, int64_t* idx_out, double* sse,
                                float* counts, float* dw, int64_t N, int K, int D, void* workspace, size_t workspace_bytes,
                                int* undecided_rows_out, void* stream) {
    VQB_CHECK_ARG(z && codebook && idx_out && workspace, "vq_assign_tc: null pointer");
    VQB_CHECK_ARG(N > 0 && K > 0 && D > 0 && (order == 0 || order == 1), "vq_assign_tc: bad arguments");
    VQB_CHECK_ARG(D % 64 == 0 && D <= 256 && K % 8 == 0, "vq_assign_tc: needs D %% 64 == 0, D <= 256 and K %% 8 == 0 (got D=%d K=%d)", D, K);
    VQB_CHECK_ARG(N < (int64_t)1 << 31, "vq_assign_tc: N too large");
    VQB_CHECK_ARG(workspace_bytes >= vqb_vq_tc_workspace_bytes(N, K, D), "vq_assign_tc: workspace too small");
    cudaStream_t st = as_stream(stream);
    uint8_t* w = (uint8_t*)workspace;
    auto take = [&](size_t bytes) { uint8_t* p = w; w += align_up(bytes, 256); return p; };
    bf16* zh = (bf16*)take((size_t)N * D * 2); bf16* zl = (bf16*)take((size_t)N * D * 2);
    bf16* eh = (bf16*)take((size_t)K * D * 2); bf16* el = (bf16*)take((size_t)K * D * 2);
    float* z_sq = (float*)take((size_t)N * 4); int* best_idx = (int*)take((size_t)N * 4);
    float* best_val = (float*)take((size_t)N * 4); float* second_val = (float*)take((size_t)N * 4);
    int* row_list = (int*)take((size_t)N * 4);
    float* cb_sq = (float*)take((size_t)K * 4);
    float* emax_sq = (float*)take(128); int* n_rows = (int*)take(128);

    split_bf16_kernel<<<(unsigned)ceil_div64(N * 32, 256), 256, 0, st>>>(z, zh, zl, z_sq, N, D);
    split_bf16_kernel<<<(unsigned)ceil_div64((int64_t)K * 32, 256), 256, 0, st>>>(codebook, eh, el, cb_sq, K, D);
    max_kernel<<<1, 256, 0, st>>>(cb_sq, emax_sq, K);
    VQB_CUDA(cudaMemsetAsync(n_rows, 0, sizeof(int), st));
    VQB_CHECK_LAUNCH("vq_tc split");

    CUtensorMap tmZh, tmZl, tmEh, tmEl;
    int rc;
    if ((rc = make_2d_map(&tmZh, zh, N, D, TM))) return rc;
    if ((rc = make_2d_map(&tmZl, zl, N, D, TM))) return rc;
    if ((rc = make_2d_map(&tmEh, eh, K, D, TN))) return rc;
    if ((rc = make_2d_map(&tmEl, el, K, D, TN))) return rc;
    SearchParams sp;
    sp.N = N; sp.K = K; sp.D = D; sp.kchunks = D / TK; sp.ctiles = (K + TN - 1) / TN;
    sp.cb_sq = cb_sq; sp.best_idx = best_idx; sp.best_val = best_val; sp.second_val = second_val;
    size_t smem = (size_t)2 * sp.kchunks * TM * TK * 2 + (size_t)2 * TN * TK * 2 + (size_t)K * 4 + 256 + 128 * 3 * 4 + 1024 + 256;
    if (smem > 227 * 1024) { vqb_set_error("vq_assign_tc: K=%d does not fit the shared-memory budget", K); return VQB_ERR_UNSUPPORTED; }
    VQB_CUDA(cudaFuncSetAttribute(vq_tc_search_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    vq_tc_search_kernel<<<(unsigned)ceil_div64(N, TM), NTH, smem, st>>>(tmZh, tmZl, tmEh, tmEl, sp);
    VQB_CHECK_LAUNCH("vq_tc_search");
    vq_tc_decide_kernel<<<(unsigned)ceil_div64(N, 256), 256, 0, st>>>(best_idx, best_val, second_val, z_sq, emax_sq, idx_out, row_list, n_rows, N);
    VQB_CHECK_LAUNCH("vq_tc_decide");
    if ((rc = vqb_vq_exact_rows(z, codebook, cb_sq, order, row_list, n_rows, idx_out, N, K, D, st))) return rc;
    if ((rc = vqb_vq_finish(z, codebook, idx_out, q_out, sse, counts, dw, N, K, D, st))) return rc;
    if (undecided_rows_out) VQB_CUDA(cudaMemcpyAsync(undecided_rows_out, n_rows, sizeof(int), cudaMemcpyDeviceToDevice, st));
    return VQB_OK;
}
