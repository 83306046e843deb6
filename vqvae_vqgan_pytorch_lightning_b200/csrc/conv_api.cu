// C-ABI entry points for convolution: dispatch between the fp32 SIMT implicit GEMM (impl 0, conv_simt.cu) and
// the tcgen05/TMA bf16 implicit GEMM (impl 1, conv_tc.cu; impl 2 / 3 = the same kernels on split-precision operands).
#include "common.cuh"
#include <stdlib.h>

int vqb_conv2d_fwd_simt(const void* x, int x_dtype, const float* wp, const float* bias, const void* residual, void* y,
                        int y_dtype, int N, int H, int W, int Ci, int Co, int KH, int KW, int pad, int stride, int act,
                        float alpha, float gain, cudaStream_t stream);
int vqb_conv2d_wgrad_simt(const void* x, int x_dtype, const void* dy, int dy_dtype, float* dwp, int N, int H, int W, int Ci,
                          int Co, int KH, int KW, int pad, int stride, cudaStream_t stream);
int vqb_conv2d_dgrad_simt(const void* dy, int dy_dtype, const float* wd, void* dx, int dx_dtype, int N, int H, int W, int Ci,
                          int Co, int KH, int KW, int pad, int stride, cudaStream_t stream);
int vqb_conv2d_fwd_tc(const void* x, const void* wp, const float* bias, const void* residual, void* y, int y_dtype, int N,
                      int H, int W, int Ci, int Co, int KH, int KW, int pad, int act, float alpha, float gain,
                      cudaStream_t stream, int Cx, double* gn_sums, int gn_groups);
int vqb_conv2d_wgrad_tc(const void* x, const void* dy, float* dwp, int N, int H, int W, int Ci, int Co, int KH, int KW,
                        int pad, cudaStream_t stream);

extern "C" int vqb_conv2d_fwd_gn(int impl, const void* x, int x_dtype, const void* wp, const float* bias, const void* residual,
                                 void* y, int y_dtype, int N, int H, int W, int Ci, int Co, int KH, int KW, int pad, int stride,
                                 int act, float act_alpha, float gain, double* gn_sums, int gn_groups, void* stream);

extern "C" int vqb_conv2d_fwd(int impl, const void* x, int x_dtype, const void* wp, const float* bias, const void* residual,
                              void* y, int y_dtype, int N, int H, int W, int Ci, int Co, int KH, int KW, int pad, int stride,
                              int act, float act_alpha, float gain, void* stream) {
    return vqb_conv2d_fwd_gn(impl, x, x_dtype, wp, bias, residual, y, y_dtype, N, H, W, Ci, Co, KH, KW, pad, stride, act, act_alpha,
                             gain, nullptr, 0, stream);
}

// 1 when vqb_conv2d_fwd_gn can fuse the GroupNorm statistics for this problem (tensor-core impl, >= 128 pixels per image, 4 / 8 / 16
// channels per group), else 0: the caller then runs vqb_gn_stats on the output as before
extern "C" int vqb_conv2d_fwd_gn_supported(int impl, int N, int H, int W, int Ci, int Co, int KH, int KW, int pad, int stride,
                                           int gn_groups) {
    if (impl < 1 || impl > 3 || stride != 1 || gn_groups <= 0 || Co % gn_groups != 0 || Co % 64 != 0 || Ci % 64 != 0) return 0;
    const int cpg = Co / gn_groups;
    if (!(cpg == 4 || cpg == 8 || cpg == 16)) return 0;
    if (H + 2 * pad - KH + 1 != H || W + 2 * pad - KW + 1 != W) return 0;
    // Measured on B200 (tools/step_breakdown.py, B = 64): the swapped-operand kernel (128-channel tiles, H >= 32: the 256^2 and 128^2
    // levels, i.e. the largest tensors) pays +0.11 ms per 128->128 @256^2 launch for a 0.17 ms statistics pass saved: fused.  The
    // CTA-pair kernel (256-channel tiles) pays +0.09 .. +0.2 ms per launch (40 shuffles per 32-channel chunk in an epilogue that
    // has to keep pace with 94 %-busy tensor pipes) for passes of 0.02 .. 0.08 ms, and the epilogue-bound generic kernel tripled
    // on the 64 -> 128 1x1 head: NOT fused unless VQB_GN_FUSE=2 asks for the CTA-pair kernel too (kept for A/B measurements).
    const bool halo = KH == 3 && KW == 3 && pad == 1 && H >= 16 && W >= 8;
    if (!halo) return 0;
    static const int level = getenv("VQB_GN_FUSE") ? atoi(getenv("VQB_GN_FUSE")) : 1;
    if (Co % 256 == 0) return (level >= 2 && Co / 256 <= 2) ? 1 : 0;
    return (Co % 128 == 0 && H >= 32) ? 1 : 0;
}

extern "C" int vqb_conv2d_fwd_gn(int impl, const void* x, int x_dtype, const void* wp, const float* bias, const void* residual,
                                 void* y, int y_dtype, int N, int H, int W, int Ci, int Co, int KH, int KW, int pad, int stride,
                                 int act, float act_alpha, float gain, double* gn_sums, int gn_groups, void* stream) {
    VQB_CHECK_ARG(x && wp && y, "conv2d_fwd: null pointer");
    VQB_CHECK_ARG(!gn_sums || impl >= 1, "conv2d_fwd: fused GroupNorm statistics need a tensor-core impl");
    if (impl == 0)
        return vqb_conv2d_fwd_simt(x, x_dtype, (const float*)wp, bias, residual, y, y_dtype, N, H, W, Ci, Co, KH, KW, pad,
                                   stride, act, act_alpha, gain, as_stream(stream));
    if (impl == 1) {
        VQB_CHECK_ARG(x_dtype == VQB_BF16, "conv2d_fwd(tcgen05): x must be bf16");
        VQB_CHECK_ARG(stride == 1, "conv2d_fwd(tcgen05): stride must be 1");
        return vqb_conv2d_fwd_tc(x, wp, bias, residual, y, y_dtype, N, H, W, Ci, Co, KH, KW, pad, act, act_alpha, gain,
                                 as_stream(stream), Ci, gn_sums, gn_groups);
    }
    if (impl == 2 || impl == 3) {
        // split-precision tcgen05 (strict numeric mode): x = [hi | lo] bf16 halves of Ci fp32 channels (2 * Ci channels), wp packed
        // K-major with 3 (impl 2) or 4 (impl 3) weight blocks per tap -- see vqb_conv2d_fwd_tc
        VQB_CHECK_ARG(x_dtype == VQB_BF16, "conv2d_fwd(tcgen05 split): x must hold bf16 [hi | lo] halves");
        VQB_CHECK_ARG(stride == 1 && Ci % 64 == 0, "conv2d_fwd(tcgen05 split): stride must be 1 and Ci a multiple of 64");
        return vqb_conv2d_fwd_tc(x, wp, bias, residual, y, y_dtype, N, H, W, (impl + 1) * Ci, Co, KH, KW, pad, act, act_alpha, gain,
                                 as_stream(stream), 2 * Ci, gn_sums, gn_groups);
    }
    vqb_set_error("conv2d_fwd: unknown impl %d", impl);
    return VQB_ERR_ARG;
}

extern "C" int vqb_conv2d_wgrad(int impl, const void* x, int x_dtype, const void* dy, int dy_dtype, float* dwp, int N, int H,
                                int W, int Ci, int Co, int KH, int KW, int pad, int stride, void* stream) {
    VQB_CHECK_ARG(x && dy && dwp, "conv2d_wgrad: null pointer");
    if (impl == 0)
        return vqb_conv2d_wgrad_simt(x, x_dtype, dy, dy_dtype, dwp, N, H, W, Ci, Co, KH, KW, pad, stride, as_stream(stream));
    if (impl == 1) {
        VQB_CHECK_ARG(x_dtype == VQB_BF16 && dy_dtype == VQB_BF16, "conv2d_wgrad(tcgen05): x and dy must be bf16");
        VQB_CHECK_ARG(stride == 1, "conv2d_wgrad(tcgen05): stride must be 1");
        return vqb_conv2d_wgrad_tc(x, dy, dwp, N, H, W, Ci, Co, KH, KW, pad, as_stream(stream));
    }
    vqb_set_error("conv2d_wgrad: unknown impl %d", impl);
    return VQB_ERR_ARG;
}

extern "C" int vqb_conv2d_dgrad(const void* dy, int dy_dtype, const void* wd, void* dx, int dx_dtype, int N, int H, int W, int Ci,
                                int Co, int KH, int KW, int pad, int stride, void* stream) {
    VQB_CHECK_ARG(dy && wd && dx, "conv2d_dgrad: null pointer");
    return vqb_conv2d_dgrad_simt(dy, dy_dtype, (const float*)wd, dx, dx_dtype, N, H, W, Ci, Co, KH, KW, pad, stride, as_stream(stream));
}
