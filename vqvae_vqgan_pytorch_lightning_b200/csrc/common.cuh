// Shared helpers for libvqgan_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/vqgan_b200.h"

typedef __nv_bfloat16 bf16;

void vqb_set_error(const char* fmt, ...);

#define VQB_CHECK_ARG(cond, ...)                         \
    do {                                                 \
        if (!(cond)) {                                   \
            vqb_set_error(__VA_ARGS__);                  \
            return VQB_ERR_ARG;                          \
        }                                                \
    } while (0)

#define VQB_CHECK_LAUNCH(name)                                                        \
    do {                                                                              \
        cudaError_t e__ = cudaGetLastError();                                         \
        if (e__ != cudaSuccess) {                                                     \
            vqb_set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));    \
            return VQB_ERR_CUDA;                                                      \
        }                                                                             \
    } while (0)

#define VQB_CUDA(call)                                                                \
    do {                                                                              \
        cudaError_t e__ = (call);                                                     \
        if (e__ != cudaSuccess) {                                                     \
            vqb_set_error("%s failed: %s", #call, cudaGetErrorString(e__));           \
            return VQB_ERR_CUDA;                                                      \
        }                                                                             \
    } while (0)

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- typed scalar / vector access -------------------------------------------------------------------
__device__ __forceinline__ float ld1(const float* p) { return *p; }
__device__ __forceinline__ float ld1(const bf16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void st1(float* p, float v) { *p = v; }
__device__ __forceinline__ void st1(bf16* p, float v) { *p = __float2bfloat16_rn(v); }

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 ld4(const bf16* p) {
    uint2 u = *reinterpret_cast<const uint2*>(p);
    __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&u.x);
    __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&u.y);
    float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
    return make_float4(fa.x, fa.y, fb.x, fb.y);
}
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void st4(bf16* p, float4 v) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&a);
    u.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = u;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ float silu_f(float t) { return t / (1.0f + expf(-t)); }
__device__ __forceinline__ float silu_grad_f(float t) {
    float s = 1.0f / (1.0f + expf(-t));
    return s * (1.0f + t * (1.0f - s));
}

// dtype dispatch helpers: call F<T>(...) with T in {float, bf16}
#define VQB_DISPATCH_1(dt, T, ...)                                     \
    if ((dt) == VQB_F32) { using T = float; __VA_ARGS__ }              \
    else if ((dt) == VQB_BF16) { using T = bf16; __VA_ARGS__ }         \
    else { vqb_set_error("bad dtype %d", (int)(dt)); return VQB_ERR_ARG; }
