// Input pipeline of the training step, one kernel (reference: BaseVQVAE.preprocess_batch with training=True,
// vqvae/modules/abstract_modules/base_autoencoder.py:17-50 -- clamp[0,1] -> kornia RandomResizedCrop(scale 0.7-1, ratio 1,
// bilinear, align_corners=True) -> RandomHorizontalFlip -> Normalize(0.5, 0.5)), fed by what the reference's loaders
// produce (common_utils.py:60-71): NCHW uint8 in [0,255], fp16 or fp32 in [0,1].  Output: NHWC fp32 / bf16 in [-1,1].
//
// HBM-bound: every source pixel of the crop is read (<= 4 taps per output pixel, served by L1/L2), every output pixel is
// written once; algorithmic bytes per image = s_in*C*h_crop*w_crop + s_out*C*OH*OW.
#include "common.cuh"

#include <cuda_fp16.h>

namespace {

template <typename T> __device__ __forceinline__ float load_px(const T* p) { return (float)(*p); }
template <> __device__ __forceinline__ float load_px<__half>(const __half* p) { return __half2float(*p); }
template <> __device__ __forceinline__ float load_px<uint8_t>(const uint8_t* p) { return (float)(*p); }

// box[n] = {x0, y0, x1, y1}: source coordinates of the centres of the first / last crop pixel (inclusive corners);
// align_corners=True resampling maps output u in [0, OW-1] to x0 + u (x1 - x0) / (OW - 1).
template <typename TI, typename TO>
__global__ void crop_flip_normalize_kernel(const TI* __restrict__ img, TO* __restrict__ out, const float* __restrict__ box,
                                           const uint8_t* __restrict__ flip, int N, int C, int H, int W, int OH, int OW,
                                           float in_scale, float mean, float inv_std) {
    const int64_t total = (int64_t)N * OH * OW;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int ow = (int)(i % OW); int64_t r = i / OW; int oh = (int)(r % OH); int n = (int)(r / OH);
        const float x0 = box[4 * n + 0], y0 = box[4 * n + 1], x1 = box[4 * n + 2], y1 = box[4 * n + 3];
        const int u = (flip && flip[n]) ? (OW - 1 - ow) : ow;
        float sx = OW > 1 ? x0 + (float)u * ((x1 - x0) / (float)(OW - 1)) : x0;
        float sy = OH > 1 ? y0 + (float)oh * ((y1 - y0) / (float)(OH - 1)) : y0;
        sx = fminf(fmaxf(sx, 0.f), (float)(W - 1)); sy = fminf(fmaxf(sy, 0.f), (float)(H - 1));
        int ix = (int)floorf(sx), iy = (int)floorf(sy);
        if (ix > W - 2) ix = W > 1 ? W - 2 : 0;
        if (iy > H - 2) iy = H > 1 ? H - 2 : 0;
        const float fx = sx - (float)ix, fy = sy - (float)iy;
        const int dx = W > 1 ? 1 : 0, dy = H > 1 ? W : 0;
        for (int c = 0; c < C; ++c) {
            const TI* p = img + (((int64_t)n * C + c) * H + iy) * W + ix;
            float v00 = load_px<TI>(p), v01 = load_px<TI>(p + dx), v10 = load_px<TI>(p + dy), v11 = load_px<TI>(p + dy + dx);
            // clamp BEFORE resampling, as the reference does (torch.clamp(images, 0, 1) precedes the augmentation)
            v00 = fminf(fmaxf(v00 * in_scale, 0.f), 1.f); v01 = fminf(fmaxf(v01 * in_scale, 0.f), 1.f);
            v10 = fminf(fmaxf(v10 * in_scale, 0.f), 1.f); v11 = fminf(fmaxf(v11 * in_scale, 0.f), 1.f);
            const float top = v00 + fx * (v01 - v00), bot = v10 + fx * (v11 - v10);
            const float v = top + fy * (bot - top);
            out[i * C + c] = (TO)((v - mean) * inv_std);
        }
    }
}

}  // namespace

// in_dtype: 0 = fp32, 1 = fp16, 2 = uint8 (scaled by 1/255)
extern "C" int vqb_crop_flip_normalize(const void* images, int in_dtype, void* out, int out_dtype, const float* boxes,
                                       const uint8_t* flip, int N, int C, int H, int W, int OH, int OW, float mean, float std,
                                       void* stream) {
    VQB_CHECK_ARG(images && out && boxes && N > 0 && C > 0 && H > 0 && W > 0 && OH > 0 && OW > 0 && std > 0.f,
                  "crop_flip_normalize: bad arguments");
    VQB_CHECK_ARG(in_dtype >= 0 && in_dtype <= 2, "crop_flip_normalize: in_dtype must be 0 (fp32), 1 (fp16) or 2 (uint8)");
    const int64_t total = (int64_t)N * OH * OW;
    int64_t blocks = (total + 255) / 256; if (blocks > 148 * 16) blocks = 148 * 16;
    cudaStream_t st = as_stream(stream);
    const float inv_std = 1.0f / std;
#define VQB_AUG_LAUNCH(TI, SCALE)                                                                                             \
    VQB_DISPATCH_1(out_dtype, TO, crop_flip_normalize_kernel<TI, TO><<<(unsigned)blocks, 256, 0, st>>>(                       \
        (const TI*)images, (TO*)out, boxes, flip, N, C, H, W, OH, OW, SCALE, mean, inv_std);)
    if (in_dtype == 0) { VQB_AUG_LAUNCH(float, 1.0f) }
    else if (in_dtype == 1) { VQB_AUG_LAUNCH(__half, 1.0f) }
    else { VQB_AUG_LAUNCH(uint8_t, 1.0f / 255.0f) }
#undef VQB_AUG_LAUNCH
    VQB_CHECK_LAUNCH("crop_flip_normalize");
    return VQB_OK;
}
