// 2-CTA (tcgen05 cta_group::2) GEMM probe: D[256 x N] = A[256 x K] * B[N x K]^T, bf16 in / fp32 out, N = 256, K = 64 * KC.
// Groundwork for CTA-pair convolution kernels: one cluster of two CTAs, CTA r holds A rows [128r, 128r+128) and B rows
// [128r, 128r+128) (= half of the N columns); the leader CTA's single thread issues M = 256 UMMAs that read both CTAs' shared
// memory; each CTA's TMEM receives its own 128 rows x N columns.  Everything is single-shot (no ring) to isolate the protocol:
//   * TMA loads of BOTH CTAs signal the LEADER's full barrier (cp.async.bulk.tensor ... .cta_group::2, barrier address mapped
//     into the leader's shared window with mapa);
//   * tcgen05.commit ... .multicast::cluster arrives on the accumulator barrier of both CTAs.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC -I../../vqvae_vqgan_pytorch_lightning_b200/csrc
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#include "ptx.cuh"

namespace {

constexpr int BM = 128, BK = 64, NTOT = 256, NHALF = 128, UK = 16;

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_rank(uint32_t saddr, uint32_t rank) {
    uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank)); return r;
}
// TMA load into the LOCAL shared memory, completion bytes signalled on a barrier that may live in the peer CTA
__device__ __forceinline__ void tma_load_2d_2sm(void* smem, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            ptx::smem_u32(smem)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ptx::smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma2_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma2_commit_multicast(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     ptx::smem_u32(bar)),
                 "h"(mask)
                 : "memory");
}

// 192 threads: warp 0 = TMA producer, warp 1 = TMEM alloc + (leader only) MMA issuer, warps 2..5 = epilogue (quarter = warp & 3)
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(192, 1)
umma2_probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* __restrict__ D, int KC,
                   int* __restrict__ dbg) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    unsigned char* sA = smem;                                  // [KC][128 x 64 bf16] = KC * 16 KB
    unsigned char* sB = smem + (size_t)KC * BM * BK * 2;       // [KC][128 x 64 bf16]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sB + (size_t)KC * NHALF * BK * 2);
    uint64_t* full = bars;                                     // leader's: all operand bytes of BOTH CTAs
    uint64_t* tfull = bars + 1;                                // accumulator ready (one per CTA, arrived by multicast commit)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();

    if (threadIdx.x == 0) {
        ptx::mbar_init(full, 1);
        ptx::mbar_init(tfull, 1);
        ptx::fence_barrier_init();
    }
    if (warp == 1) tmem_alloc2(tmem_slot, NTOT);
    ptx::tc_fence_before();
    __syncthreads();
    cluster_sync();                                            // barriers initialised and TMEM allocated in BOTH CTAs
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0 && dbg) dbg[rank] = (int)tmem_base;

    if (warp == 0) {
        if (ptx::elect_one()) {
            const uint32_t bytes_per_cta = (uint32_t)KC * (BM + NHALF) * BK * 2;
            if (rank == 0) ptx::mbar_expect_tx(full, 2 * bytes_per_cta);
            const uint32_t leader_full = mapa_rank(ptx::smem_u32(full), 0);
            for (int kc = 0; kc < KC; ++kc) {
                tma_load_2d_2sm(sA + (size_t)kc * BM * BK * 2, &tmA, leader_full, kc * BK, (int)rank * BM);
                tma_load_2d_2sm(sB + (size_t)kc * NHALF * BK * 2, &tmB, leader_full, kc * BK, (int)rank * NHALF);
            }
        }
    } else if (warp == 1) {
        if (rank == 0 && ptx::elect_one()) {
            ptx::mbar_wait(full, 0);
            ptx::tc_fence_after();
            const uint32_t idesc = ptx::umma_idesc_bf16(2 * BM, NTOT, 0, 0);
            for (int kc = 0; kc < KC; ++kc) {
                const uint64_t adesc = ptx::umma_smem_desc(ptx::smem_u32(sA + (size_t)kc * BM * BK * 2), 0, 1024);
                const uint64_t bdesc = ptx::umma_smem_desc(ptx::smem_u32(sB + (size_t)kc * NHALF * BK * 2), 0, 1024);
#pragma unroll
                for (int k = 0; k < BK / UK; ++k)
                    umma2_bf16(tmem_base, adesc + (uint64_t)(k * UK * 2 / 16), bdesc + (uint64_t)(k * UK * 2 / 16), idesc, (kc | k) != 0 ? 1u : 0u);
            }
            umma2_commit_multicast(tfull, 0b11);
        }
    } else {
        const int quarter = warp & 3;
        ptx::mbar_wait(tfull, 0);
        ptx::tc_fence_after();
        const int row = (int)rank * BM + quarter * 32 + lane;
        for (int c = 0; c < NTOT; c += 32) {
            uint32_t r[32];
            ptx::tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c, r);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) D[(size_t)row * NTOT + c + j] = __uint_as_float(r[j]);
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    cluster_sync();                                            // both CTAs are done with TMEM / peer shared memory
    if (warp == 1) {
        ptx::tc_fence_after();
        tmem_dealloc2(tmem_base, NTOT);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_map(CUtensorMap* m, const void* base, int rows, int K, int box_rows) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return 1;
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = reinterpret_cast<EncodeTiledFn>(p)(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, es,
                                                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : 2;
}

}  // namespace

// A [256][K] bf16, B [256][K] bf16 (both K-major), D [256][256] fp32, dbg int[2] (TMEM base of each CTA); K = 64 * KC, KC <= 5
extern "C" int umma2_probe(const void* A, const void* B, float* D, int K, int* dbg, void* stream) {
    if (K % 64 != 0 || K <= 0 || K > 320) return -1;
    const int KC = K / 64;
    CUtensorMap tmA, tmB;
    if (make_map(&tmA, A, 256, K, BM) || make_map(&tmB, B, 256, K, NHALF)) return -2;
    size_t smem = (size_t)KC * (BM + NHALF) * BK * 2 + 1024 + 256;
    if (cudaFuncSetAttribute(umma2_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -3;
    umma2_probe_kernel<<<2, 192, smem, (cudaStream_t)stream>>>(tmA, tmB, D, KC, dbg);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { fprintf(stderr, "umma2_probe launch: %s\n", cudaGetErrorString(e)); return -4; }
    return 0;
}
