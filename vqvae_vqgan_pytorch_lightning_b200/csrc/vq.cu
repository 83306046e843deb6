// Fused vector-quantisation kernels (reference: vqvae/modules/vector_quantizers.py).
//
// vq_assign: one kernel does   pairwise L2 distance (never materialised)  ->  running argmin (torch.argmin
// first-index semantics)  ->  gather e[idx]  ->  straight-through forward value  ->  sum (e-z)^2  ->  code
// histogram  ->  EMA cluster sums (atomics).  The reference materialises the N x K distance matrix and an N x K
// one-hot matrix and runs three GEMMs (:37-49, :142-166); here the only HBM traffic is z (read), the codebook
// (read, L2-resident across CTAs), q and idx (write) and the K x D scatter targets.
//
// Numerics (strict path): every distance is evaluated in fp32 with the reference's operation order
//   order 0: (|z|^2 + |e|^2) - 2*dot      order 1: (|z|^2 - 2*dot) + |e|^2
// so that the argmin agrees with the fp32 oracle except where the oracle's own distances tie within an ulp
// ("argmin tie fragility"); the dot product itself is fp32 FMA accumulation.
//
// Tiling: a CTA owns 64 latent rows (kept in shared memory, transposed) and streams the codebook in tiles of
// 128 codes x 16 dims with register-staged double buffering; 256 threads, 4 x 8 micro-tile.
#include "common.cuh"

namespace {

constexpr int VM = 64, VN = 128, VK = 16, VPAD = 4;

__global__ void row_sqnorm_kernel(const float* __restrict__ a, float* __restrict__ out, int64_t R, int D) {
    // one warp per row
    int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (row >= R) return;
    float s = 0.f;
    for (int d = lane; d < D; d += 32) { float v = a[row * D + d]; s = fmaf(v, v, s); }
    s = warp_sum(s);
    if (lane == 0) out[row] = s;
}

__device__ __forceinline__ void
vq_assign_tile(const int64_t tile_idx, const float* __restrict__ z, const float* __restrict__ cb, const float* __restrict__ cb_sq, int order,
                 float* __restrict__ q_out, int64_t* __restrict__ idx_out, double* __restrict__ sse,
                 float* __restrict__ counts, float* __restrict__ dw, int64_t N, int K, int D,
                 const int* __restrict__ row_list, const int* __restrict__ n_rows_dev) {
    // list mode (row_list != NULL): the CTA re-evaluates rows row_list[64*blockIdx.x ...] exactly and only writes idx_out
    extern __shared__ __align__(16) float smem[];
    float (*zs)[VM + VPAD] = reinterpret_cast<float (*)[VM + VPAD]>(smem);                        // [D][VM+PAD]
    float (*Bs)[VK][VN + VPAD] = reinterpret_cast<float (*)[VK][VN + VPAD]>(smem + (size_t)D * (VM + VPAD));   // [2][VK][VN+PAD]
    __shared__ float z_sq[VM];
    __shared__ int best_idx_s[VM];

    const int tid = threadIdx.x;
    const int ty = tid >> 4, tx = tid & 15;
    const int64_t r0 = tile_idx * VM;
    const int64_t n_valid = row_list ? (int64_t)n_rows_dev[0] : N;
    if (r0 >= n_valid) return;
    __shared__ int64_t grow_s[VM];
    if (tid < VM) grow_s[tid] = (r0 + tid < n_valid) ? (row_list ? (int64_t)row_list[r0 + tid] : r0 + tid) : -1;
    __syncthreads();

    // ---- stage the z rows (transposed) and their squared norms
    for (int i = tid; i < VM * (D / 4); i += 256) {
        int row = i / (D / 4), d4 = (i - row * (D / 4)) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (grow_s[row] >= 0) v = *reinterpret_cast<const float4*>(z + grow_s[row] * D + d4);
        zs[d4 + 0][row] = v.x; zs[d4 + 1][row] = v.y; zs[d4 + 2][row] = v.z; zs[d4 + 3][row] = v.w;
    }
    __syncthreads();
    {
        // 4 threads per row
        int row = tid >> 2, part = tid & 3;
        float s = 0.f;
        for (int d = part; d < D; d += 4) { float v = zs[d][row]; s = fmaf(v, v, s); }
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        if (part == 0) z_sq[row] = s;
    }
    __syncthreads();

    float best_v[4];
    int best_i[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { best_v[i] = INFINITY; best_i[i] = 0x7fffffff; }
    float zsq[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) zsq[i] = z_sq[ty * 4 + i];

    const int lr = tid >> 2, lkc = (tid & 3) * 4;   // codebook loader: codes lr, lr+64; dims lkc..lkc+3
    const int DK = D / VK + ((D % VK) ? 1 : 0);
    const int CT = (K + VN - 1) / VN;
    // list mode: blockIdx.y selects a slice of the code tiles (partial minima are merged with a 64-bit atomicMin)
    const int ct_per = (CT + gridDim.y - 1) / gridDim.y;
    const int ct_begin = blockIdx.y * ct_per;
    const int ct_end = (ct_begin + ct_per < CT) ? ct_begin + ct_per : CT;
    const int total = (ct_end > ct_begin ? ct_end - ct_begin : 0) * DK;
    if (total == 0) return;
    float b_reg[2][4];
    auto load_b = [&](int it) {
        int ct = ct_begin + it / DK, dk = it % DK;
        int d = dk * VK + lkc;
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            int code = ct * VN + lr + r * 64;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (code < K && d < D) v = *reinterpret_cast<const float4*>(cb + (int64_t)code * D + d);
            b_reg[r][0] = v.x; b_reg[r][1] = v.y; b_reg[r][2] = v.z; b_reg[r][3] = v.w;
        }
    };
    auto store_b = [&](int buf) {
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int j = 0; j < 4; ++j) Bs[buf][lkc + j][lr + r * 64] = b_reg[r][j];
    };

    float acc[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    load_b(0);
    store_b(0);
    __syncthreads();
    for (int it = 0; it < total; ++it) {
        const int cur = it & 1;
        const int ct = ct_begin + it / DK, dk = it % DK;
        if (it + 1 < total) load_b(it + 1);
        const int dbase = dk * VK;
        const int klim = (D - dbase < VK) ? (D - dbase) : VK;
#pragma unroll
        for (int k = 0; k < VK; ++k) {
            if (k < klim) {
                float4 a0 = *reinterpret_cast<const float4*>(&zs[dbase + k][ty * 4]);
                float4 b0 = *reinterpret_cast<const float4*>(&Bs[cur][k][tx * 4]);
                float4 b1 = *reinterpret_cast<const float4*>(&Bs[cur][k][64 + tx * 4]);
                float a[4] = {a0.x, a0.y, a0.z, a0.w};
                float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
            }
        }
        if (dk == DK - 1) {
            // distances for this code tile, running first-index argmin
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                int code = ct * VN + ((j < 4) ? (tx * 4 + j) : (64 + tx * 4 + j - 4));
                if (code < K) {
                    float e2 = cb_sq[code];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        float two_dot = 2.0f * acc[i][j];
                        float dist = (order == 0) ? __fsub_rn(__fadd_rn(zsq[i], e2), two_dot)
                                                  : __fadd_rn(__fsub_rn(zsq[i], two_dot), e2);
                        if (dist < best_v[i] || (dist == best_v[i] && code < best_i[i])) { best_v[i] = dist; best_i[i] = code; }
                    }
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[i][j] = 0.f;
            }
        }
        if (it + 1 < total) store_b(cur ^ 1);
        __syncthreads();
    }

    // reduce over the 16 threads (tx) that share rows ty*4..ty*4+3: lanes differ in the low 4 bits
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
            float ov = __shfl_xor_sync(0xffffffffu, best_v[i], o);
            int oi = __shfl_xor_sync(0xffffffffu, best_i[i], o);
            if (ov < best_v[i] || (ov == best_v[i] && oi < best_i[i])) { best_v[i] = ov; best_i[i] = oi; }
        }
        if (tx == 0) {
            best_idx_s[ty * 4 + i] = best_i[i];
            if (row_list && grow_s[ty * 4 + i] >= 0) {
                // order-preserving float -> uint map, code in the low word: atomicMin == (smallest distance, then smallest index)
                uint32_t u = __float_as_uint(best_v[i]);
                u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
                unsigned long long key = ((unsigned long long)u << 32) | (uint32_t)best_i[i];
                atomicMin(reinterpret_cast<unsigned long long*>(idx_out) + grow_s[ty * 4 + i], key);
            }
        }
    }
    __syncthreads();
    if (row_list) return;
    // ---- phase 2: gather, straight-through value, loss, histogram, EMA cluster sums
    const int warp = tid >> 5, lane = tid & 31;
    float sse_local = 0.f;
    for (int rr = 0; rr < VM / 8; ++rr) {
        int row = warp * (VM / 8) + rr;
        int64_t grow = r0 + row;
        if (grow >= N) break;
        int code = best_idx_s[row];
        if (code < 0 || code >= K) code = 0;   // NaN rows: torch.argmin would return the NaN position; keep memory-safe
        if (lane == 0) {
            idx_out[grow] = (int64_t)code;
            if (counts) atomicAdd(counts + code, 1.0f);
        }
        const float* e = cb + (int64_t)code * D;
        for (int d = lane; d < D; d += 32) {
            float zv = zs[d][row];
            float diff = e[d] - zv;
            sse_local = fmaf(diff, diff, sse_local);
            if (q_out) q_out[grow * D + d] = zv + diff;       // flat_x + (quantized - flat_x).detach()
            if (dw) atomicAdd(dw + (int64_t)code * D + d, zv);
        }
    }
    if (sse) {
        sse_local = warp_sum(sse_local);
        if (lane == 0) atomicAdd(sse, (double)sse_local);
    }
}

// phase D of the tensor-core path: one warp per row -- gather, straight-through value, loss, histogram, EMA cluster sums
__global__ void vq_finish_kernel(const float* __restrict__ z, const float* __restrict__ cb, const int64_t* __restrict__ idx,
                                 float* __restrict__ q_out, double* __restrict__ sse, float* __restrict__ counts,
                                 float* __restrict__ dw, int64_t N, int K, int D) {
    int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    float sse_local = 0.f;
    if (row < N) {
        int64_t code = idx[row];
        if (code < 0 || code >= K) code = 0;
        if (lane == 0 && counts) atomicAdd(counts + code, 1.0f);
        const float* e = cb + code * D;
        for (int d = lane; d < D; d += 32) {
            float zv = z[row * D + d];
            float diff = e[d] - zv;
            sse_local = fmaf(diff, diff, sse_local);
            if (q_out) q_out[row * D + d] = zv + diff;
            if (dw) atomicAdd(dw + code * D + d, zv);
        }
    }
    if (sse) {
        sse_local = warp_sum(sse_local);
        __shared__ float sh[8];
        if (lane == 0) sh[threadIdx.x >> 5] = sse_local;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += sh[i];
            atomicAdd(sse, t);
        }
    }
}

__global__ void __launch_bounds__(256)
vq_assign_kernel(const float* __restrict__ z, const float* __restrict__ cb, const float* __restrict__ cb_sq, int order,
                 float* __restrict__ q_out, int64_t* __restrict__ idx_out, double* __restrict__ sse,
                 float* __restrict__ counts, float* __restrict__ dw, int64_t N, int K, int D) {
    vq_assign_tile(blockIdx.x, z, cb, cb_sq, order, q_out, idx_out, sse, counts, dw, N, K, D, nullptr, nullptr);
}

// list mode: a fixed, small grid walks the (device-side counted) undecided rows, so a handful of rows costs a handful of
// CTAs instead of a worst-case grid of early-exit launches
__global__ void __launch_bounds__(256)
vq_assign_rows_kernel(const float* __restrict__ z, const float* __restrict__ cb, const float* __restrict__ cb_sq, int order,
                      int64_t* __restrict__ idx_out, int64_t N, int K, int D, const int* __restrict__ row_list,
                      const int* __restrict__ n_rows_dev) {
    const int64_t n = n_rows_dev[0];
    for (int64_t tile = blockIdx.x; tile * VM < n; tile += gridDim.x) {
        vq_assign_tile(tile, z, cb, cb_sq, order, nullptr, idx_out, nullptr, nullptr, nullptr, N, K, D, row_list, n_rows_dev);
        __syncthreads();
    }
}

__global__ void vq_ema_update_kernel(float* __restrict__ ema_count, float* __restrict__ ema_weight, float* __restrict__ cb,
                                     const float* __restrict__ counts, const float* __restrict__ dw, int K, int D,
                                     float decay, float eps, float batch) {
    // one warp per code
    int code = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (code >= K) return;
    float c = ema_count[code] * decay + (1.0f - decay) * counts[code];
    float cnt = (c + eps) / (batch + (float)K * eps) * batch;
    for (int d = lane; d < D; d += 32) {
        int64_t o = (int64_t)code * D + d;
        float w = ema_weight[o] * decay + (1.0f - decay) * dw[o];
        ema_weight[o] = w;
        cb[o] = w / cnt;
    }
    __syncwarp();
    if (lane == 0) ema_count[code] = cnt;
}

__global__ void vq_backward_kernel(const float* __restrict__ z, const float* __restrict__ q, const int64_t* __restrict__ idx,
                                   const float* __restrict__ g_q, const float* __restrict__ g_loss, float beta, float cb_scale,
                                   float* __restrict__ dz, float* __restrict__ dcb, int64_t N, int K, int D) {
    // one warp per row
    int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (row >= N) return;
    float gl = g_loss ? g_loss[0] : 1.0f;
    float inv = 2.0f / ((float)N * (float)D);
    float cz = gl * beta * inv, cc = gl * cb_scale * inv;
    int64_t code = idx[row];
    if (code < 0 || code >= K) return;
    for (int d = lane; d < D; d += 32) {
        float zv = z[row * D + d], ev = q[row * D + d];
        float diff = zv - ev;
        if (dz) dz[row * D + d] = (g_q ? g_q[row * D + d] : 0.f) + cz * diff;
        if (dcb) atomicAdd(dcb + code * D + d, -cc * diff);
    }
}

__global__ void vq_gather_kernel(const float* __restrict__ cb, const int64_t* __restrict__ idx, float* __restrict__ out,
                                 int64_t N, int K, int D) {
    int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (row >= N) return;
    int64_t code = idx[row];
    for (int d = lane; d < D; d += 32) out[row * D + d] = (code >= 0 && code < K) ? cb[code * D + d] : 0.f;
}

}  // namespace

extern "C" size_t vqb_vq_workspace_bytes(int64_t N, int K, int D) {
    (void)N; (void)D;
    return (size_t)K * sizeof(float);
}

extern "C" int vqb_vq_assign(const float* z, const float* codebook, int order, float* q_out, int64_t* idx_out, double* sse,
                             float* counts, float* dw, int64_t N, int K, int D, void* workspace, size_t workspace_bytes,
                             void* stream) {
    VQB_CHECK_ARG(z && codebook && idx_out && workspace, "vq_assign: null pointer");
    VQB_CHECK_ARG(N > 0 && K > 0 && D > 0, "vq_assign: empty problem N=%lld K=%d D=%d", (long long)N, K, D);
    VQB_CHECK_ARG(order == 0 || order == 1, "vq_assign: order must be 0 or 1");
    VQB_CHECK_ARG(workspace_bytes >= vqb_vq_workspace_bytes(N, K, D), "vq_assign: workspace too small");
    if (D % 4 != 0) { vqb_set_error("vq_assign: embedding_dim must be a multiple of 4 (got %d)", D); return VQB_ERR_UNSUPPORTED; }
    size_t smem = ((size_t)D * (VM + VPAD) + 2 * (size_t)VK * (VN + VPAD)) * sizeof(float);
    if (smem > 227 * 1024) { vqb_set_error("vq_assign: embedding_dim %d needs %zu B of shared memory", D, smem); return VQB_ERR_UNSUPPORTED; }
    cudaStream_t st = as_stream(stream);
    float* cb_sq = (float*)workspace;
    row_sqnorm_kernel<<<(unsigned)ceil_div64((int64_t)K * 32, 256), 256, 0, st>>>(codebook, cb_sq, K, D);
    VQB_CHECK_LAUNCH("vq row_sqnorm");
    VQB_CUDA(cudaFuncSetAttribute(vq_assign_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    vq_assign_kernel<<<(unsigned)ceil_div64(N, VM), 256, smem, st>>>(z, codebook, cb_sq, order, q_out, idx_out, sse, counts, dw, N, K, D);
    VQB_CHECK_LAUNCH("vq_assign");
    return VQB_OK;
}

namespace {
__global__ void vq_keys_init_kernel(const int* __restrict__ row_list, const int* __restrict__ n_rows, int64_t* __restrict__ idx_out, int64_t N) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n_rows[0]) reinterpret_cast<unsigned long long*>(idx_out)[row_list[i]] = ~0ull;
}
__global__ void vq_keys_resolve_kernel(const int* __restrict__ row_list, const int* __restrict__ n_rows, int64_t* __restrict__ idx_out, int K) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n_rows[0]) {
        unsigned long long key = reinterpret_cast<unsigned long long*>(idx_out)[row_list[i]];
        uint32_t code = (uint32_t)(key & 0xffffffffull);
        idx_out[row_list[i]] = (int64_t)((code < (uint32_t)K) ? code : 0);
    }
}
}  // namespace

int vqb_vq_exact_rows(const float* z, const float* codebook, const float* cb_sq, int order, const int* row_list,
                      const int* n_rows_dev, int64_t* idx_out, int64_t N, int K, int D, cudaStream_t st) {
    size_t smem = ((size_t)D * (VM + VPAD) + 2 * (size_t)VK * (VN + VPAD)) * sizeof(float);
    if (smem > 227 * 1024 || D % 4 != 0) { vqb_set_error("vq_exact_rows: unsupported embedding_dim %d", D); return VQB_ERR_UNSUPPORTED; }
    VQB_CUDA(cudaFuncSetAttribute(vq_assign_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // idx_out doubles as the 64-bit (distance, code) key array of the undecided rows; the grid covers the worst case
    // (every row undecided), CTAs beyond the device-side count exit immediately; code tiles are split 8 ways so that a
    // handful of undecided rows still spreads over many SMs
    const int CT = (K + VN - 1) / VN;
    const int ysplit = CT < 8 ? CT : 8;
    vq_keys_init_kernel<<<(unsigned)ceil_div64(N, 256), 256, 0, st>>>(row_list, n_rows_dev, idx_out, N);
    int64_t gx = ceil_div64(N, VM); if (gx > 2 * 148 / ysplit + 1) gx = 2 * 148 / ysplit + 1;     // ~2 CTAs per SM in total
    dim3 grid((unsigned)gx, ysplit);
    vq_assign_rows_kernel<<<grid, 256, smem, st>>>(z, codebook, cb_sq, order, idx_out, N, K, D, row_list, n_rows_dev);
    vq_keys_resolve_kernel<<<(unsigned)ceil_div64(N, 256), 256, 0, st>>>(row_list, n_rows_dev, idx_out, K);
    VQB_CHECK_LAUNCH("vq_exact_rows");
    return VQB_OK;
}

int vqb_vq_finish(const float* z, const float* codebook, const int64_t* idx, float* q_out, double* sse, float* counts, float* dw,
                  int64_t N, int K, int D, cudaStream_t st) {
    vq_finish_kernel<<<(unsigned)ceil_div64(N * 32, 256), 256, 0, st>>>(z, codebook, idx, q_out, sse, counts, dw, N, K, D);
    VQB_CHECK_LAUNCH("vq_finish");
    return VQB_OK;
}

extern "C" int vqb_vq_ema_update(float* ema_count, float* ema_weight, float* codebook, const float* counts, const float* dw,
                                 int K, int D, float decay, float eps, float batch, void* stream) {
    VQB_CHECK_ARG(ema_count && ema_weight && codebook && counts && dw && K > 0 && D > 0, "vq_ema_update: bad arguments");
    vq_ema_update_kernel<<<(unsigned)ceil_div64((int64_t)K * 32, 256), 256, 0, as_stream(stream)>>>(
        ema_count, ema_weight, codebook, counts, dw, K, D, decay, eps, batch);
    VQB_CHECK_LAUNCH("vq_ema_update");
    return VQB_OK;
}

extern "C" int vqb_vq_backward(const float* z, const float* q, const int64_t* idx, const float* g_q, const float* g_loss,
                               float beta, float cb_scale, float* dz, float* dcb, int64_t N, int K, int D, void* stream) {
    VQB_CHECK_ARG(z && q && idx && N > 0 && K > 0 && D > 0, "vq_backward: bad arguments");
    vq_backward_kernel<<<(unsigned)ceil_div64(N * 32, 256), 256, 0, as_stream(stream)>>>(z, q, idx, g_q, g_loss, beta,
                                                                                          cb_scale, dz, dcb, N, K, D);
    VQB_CHECK_LAUNCH("vq_backward");
    return VQB_OK;
}

extern "C" int vqb_vq_gather(const float* codebook, const int64_t* idx, float* out, int64_t N, int K, int D, void* stream) {
    VQB_CHECK_ARG(codebook && idx && out && N > 0 && K > 0 && D > 0, "vq_gather: bad arguments");
    vq_gather_kernel<<<(unsigned)ceil_div64(N * 32, 256), 256, 0, as_stream(stream)>>>(codebook, idx, out, N, K, D);
    VQB_CHECK_LAUNCH("vq_gather");
    return VQB_OK;
}
