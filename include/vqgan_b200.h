/*
 * vqgan_b200.h -- C ABI of libvqgan_b200.so (hand-written sm_100a CUDA kernels for the VQ-VAE / VQGAN
 * training step).
 *
 * Conventions:
 *   - every entry point is extern "C", takes raw DEVICE pointers, plain integer sizes and the CUDA stream to
 *     launch on (cudaStream_t passed as void*), and returns 0 on success or a negative vqb_status;
 *     vqb_last_error() returns a thread-local human-readable message for the last failure;
 *   - nothing is allocated inside the library: the caller owns inputs, outputs and workspaces;
 *   - nothing synchronises: kernels are enqueued on `stream` and the call returns;
 *   - activations are NHWC ("channels last") -- logical [N,C,H,W] tensors stored as [N][H][W][C];
 *   - dtype arguments are vqb_dtype (VQB_F32 = fp32 storage, VQB_BF16 = bf16 storage; arithmetic is fp32 in
 *     the SIMT kernels and bf16 x bf16 -> fp32 in the tcgen05 kernels);
 *   - the reference interface each entry point replaces is cited as file:line relative to the reference repo
 *     (SerezD/vqvae-vqgan-pytorch-lightning @ 277d909).
 *
 * The reference's own native boundary is the StyleGAN2 op loader (custom_ops.get_plugin,
 * vqvae/modules/loss/stylegan2_discriminator/utils/custom_ops.py:49) with POD parameter structs launched on
 * the current ATen stream (ops/bias_act.cpp:32-90, ops/upfirdn2d.cpp:16-92); this header generalises that
 * pattern to the whole training step.  INTEGRATION.md shows the ctypes binding a maintainer adds.
 */
#ifndef VQGAN_B200_H
#define VQGAN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum { VQB_OK = 0, VQB_ERR_ARG = -1, VQB_ERR_CUDA = -2, VQB_ERR_UNSUPPORTED = -3 } vqb_status;
typedef enum { VQB_F32 = 0, VQB_BF16 = 1 } vqb_dtype;
typedef enum { VQB_ACT_NONE = 0, VQB_ACT_TANH = 1, VQB_ACT_SILU = 2, VQB_ACT_LRELU = 3, VQB_ACT_RELU = 4 } vqb_act;

const char* vqb_last_error(void);
/* library / build identification: returns e.g. "vqgan_b200 0.1 sm_100a" */
const char* vqb_version(void);
/* 1 when the current device is compute capability 10.x (tcgen05 kernels usable), else 0 */
int vqb_device_supports_tcgen05(void);

/* ------------------------------------------------------------------------------------------------------
 * Layout plumbing
 * ---------------------------------------------------------------------------------------------------- */
/* NCHW fp32 -> NHWC (dtype out), optionally clamp to [lo,hi] then (x-shift)*scale: the image normalisation
 * of BaseVQVAE.preprocess_batch (vqvae/modules/abstract_modules/base_autoencoder.py:41-50) without kornia
 * augmentation.  do_clamp=0 skips the clamp. */
int vqb_nchw_to_nhwc(const float* x, void* y, int out_dtype, int64_t N, int64_t C, int64_t H, int64_t W,
                     int do_clamp, float lo, float hi, float shift, float scale, void* stream);

/* Training-time input pipeline in one kernel (BaseVQVAE.preprocess_batch with training=True,
 * vqvae/modules/abstract_modules/base_autoencoder.py:17-50): clamp[0,1] -> random resized crop (bilinear,
 * align_corners=True; box[n] = {x0,y0,x1,y1} = source coordinates of the first / last crop pixel centre) -> optional
 * horizontal flip (flip[n] != 0; flip may be NULL) -> (x - mean) / std.  images: NCHW, in_dtype 0 = fp32 in [0,1],
 * 1 = fp16 in [0,1], 2 = uint8 in [0,255] (the three formats of the reference loaders, common_utils.py:60-71);
 * out: NHWC [N][OH][OW][C] of out_dtype. */
int vqb_crop_flip_normalize(const void* images, int in_dtype, void* out, int out_dtype, const float* boxes,
                            const uint8_t* flip, int N, int C, int H, int W, int OH, int OW, float mean, float std,
                            void* stream);
/* NHWC (dtype in) -> NCHW fp32, y = x*scale + shift, optional clamp (base_autoencoder.py:52-61). */
int vqb_nhwc_to_nchw(const void* x, int in_dtype, float* y, int64_t N, int64_t C, int64_t H, int64_t W,
                     float scale, float shift, int do_clamp, float lo, float hi, void* stream);
/* fp32 [rows][C] -> bf16 [rows][2C] = [hi | lo], x ~ hi + lo to 16 mantissa bits: the operand layout of the split-precision
 * tensor-core convolutions (vqb_conv2d_fwd impl 2 / 3), which keep the strict numeric mode's 1e-4 parity on tcgen05. */
int vqb_split_hi_lo(const float* x, void* y, int64_t rows, int C, void* stream);
/* dtype conversion of a flat buffer */
int vqb_convert(const void* x, int in_dtype, void* y, int out_dtype, int64_t n, void* stream);

/* Conv2d weight [Co][Ci][KH][KW] fp32 (nn.Conv2d layout, vqvae/modules/autoencoder.py:55-61) ->
 *   mode 0 (forward  GEMM-B): wp[(kh*KW+kw)*Ci+ci][co]
 *   mode 1 (dgrad    GEMM-B): wp[((KH-1-kh)*KW+(KW-1-kw))*Co+co][ci]   (taps flipped, channels swapped)
 *   mode 2 (tcgen05 forward, K-major): wp[co][(kh*KW+kw)*Ci+ci]
 *   mode 3 (tcgen05 dgrad,   K-major): wp[ci][((KH-1-kh)*KW+(KW-1-kw))*Co+co]
 *   mode 4 / 5: modes 2 / 3 with the reduction dimension zero-padded to 64 (KH*KW*C <= 64): wp[co][64] / wp[ci][64]
 * scale multiplies every weight (StyleGAN2 equalised-lr gain, discriminator.py:148,165). */
int vqb_pack_conv_weight(const float* w, void* wp, int out_dtype, int mode, int Co, int Ci, int KH, int KW,
                         float scale, void* stream);
/* Every kernel-layout weight copy of a model in ONE launch (they are all stale after an optimizer step).  desc_table: device
 * array of n_desc records of vqb_pack_desc_bytes() bytes each --
 *   { const float* w; void* wp; int mode, out_is_bf16, Co, Ci, KH, KW; float scale; int pad; long long start; }
 * (modes as in vqb_pack_conv_weight; `start` = first element of the record in the concatenated output index space, ascending,
 * each record padded to a multiple of 4096 elements: one CTA serves 4096 elements of ONE record);
 * total_elems = end of the last padded record. */
size_t vqb_pack_desc_bytes(void);
int vqb_pack_conv_weights_batched(const void* desc_table, int n_desc, int64_t total_elems, void* stream);
/* Narrow-input 3x3 / pad-1 convolutions (the RGB heads: encoder.conv_in 3->128, and everything that touches the 3-channel
 * side of decoder.conv_out) run on tensor cores as a 64-channel 1x1 implicit GEMM over this im2col tensor:
 * P[n,h,w,j] = x[n,h+kh-1,w+kw-1,c] for j = (kh*3+kw)*C+c < 9*C, zero otherwise (C <= 7).  write_all = 0 writes only the
 * first ceil(9C/8)*8 columns: for a P buffer whose remaining columns the caller keeps at zero across calls. */
int vqb_im2col3x3_narrow(const void* x, int x_dtype, void* P, int p_dtype, int N, int H, int W, int C, int write_all,
                         void* stream);
/* inverse of mode 0 for gradients: dw[co][ci][kh][kw] = scale * dwp[(kh*KW+kw)*Ci+ci][co] */
int vqb_unpack_conv_wgrad(const float* dwp, float* dw, int Co, int Ci, int KH, int KW, float scale, void* stream);
/* the same with accumulate != 0: dw += (dw = the parameter's .grad inside the optimizer's flat gradient buffer: replaces autograd's
 * AccumulateGrad pass, model.py:244-264 manual_backward) and rezero != 0: dwp is cleared behind the read (a persistent per-weight
 * partial-sum buffer then never needs a fill launch) */
int vqb_unpack_conv_wgrad_acc(float* dwp, float* dw, int Co, int Ci, int KH, int KW, float scale, int accumulate, int rezero,
                              void* stream);

/* ------------------------------------------------------------------------------------------------------
 * Convolution as implicit GEMM (replaces F.conv2d behind nn.Conv2d: autoencoder.py:55-61,102,114,133,153,170)
 *   y[n,oh,ow,co] = act( sum_{kh,kw,ci} x[n, oh*stride-pad+kh, ow*stride-pad+kw, ci] * w[co,ci,kh,kw]
 *                        + bias[co] ) * gain + residual[n,oh,ow,co]
 * impl 0 = fp32 SIMT (odd shapes of the strict path), impl 1 = tcgen05/TMA bf16 (fast path; x,y bf16; Ci%64==0, Co%64==0),
 * impl 2 / 3 = the same tcgen05 kernels on SPLIT-PRECISION operands (strict path, 1e-4 parity): x holds [hi | lo] bf16 halves
 *   of Ci fp32 channels (vqb_split_hi_lo: 2*Ci bf16 channels), wp is the K-major pack (mode 2) of the fp32 weight
 *   [wh | wh | wl] (impl 2: xh.wh + xl.wh + xh.wl) or [wh | wh | wl | wl] (impl 3: + xl.wl) along the input-channel axis,
 *   all terms accumulated in one fp32 TMEM accumulator; y is normally fp32.
 * wp is the packed weight for that impl (mode 0 for impl 0, mode 2 for impl 1-3).  bias / residual may be NULL.
 * dgrad is the same call on dy with the dgrad-packed weight (stride 1 only).
 * ---------------------------------------------------------------------------------------------------- */
int vqb_conv2d_fwd(int impl, const void* x, int x_dtype, const void* wp, const float* bias, const void* residual,
                   void* y, int y_dtype, int N, int H, int W, int Ci, int Co, int KH, int KW, int pad, int stride,
                   int act, float act_alpha, float gain, void* stream);
/* vqb_conv2d_fwd that ALSO accumulates the GroupNorm statistics of its output (the consumer is GroupNorm(gn_groups, Co),
 * autoencoder.py:25-39, 66-73): gn_sums [N][gn_groups][2] double, caller zero-fills, += (sum, sum of squares) of the final
 * output values per (image, group) from the epilogue registers -- the separate statistics pass over the activation
 * (vqb_gn_stats) disappears.  Tensor-core impls only; vqb_conv2d_fwd_gn_supported tells whether a problem qualifies
 * (>= 128 pixels per image, 4 / 8 / 16 channels per group); gn_sums == NULL is plain vqb_conv2d_fwd. */
int vqb_conv2d_fwd_gn(int impl, const void* x, int x_dtype, const void* wp, const float* bias, const void* residual,
                      void* y, int y_dtype, int N, int H, int W, int Ci, int Co, int KH, int KW, int pad, int stride,
                      int act, float act_alpha, float gain, double* gn_sums, int gn_groups, void* stream);
int vqb_conv2d_fwd_gn_supported(int impl, int N, int H, int W, int Ci, int Co, int KH, int KW, int pad, int stride,
                                int gn_groups);
/* Narrow-OUTPUT 3x3 'same' convolution, Co <= 3 (decoder.conv_out 128 -> 3 + tanh, autoencoder.py:170,178), bf16 x:
 * one [128 halo pixels] x [32 >= 9 Co] x [Ci] tensor-core GEMM per 16 x 8 halo tile gives the per-tap partial products, the epilogue
 * shift-adds them over the 3x3 neighbourhood for the 14 x 6 interior (+ bias, activation, gain).  wp = [32][Ci] bf16 with row
 * (tap * Co + co) = w[co][:, tap], tap = kh * 3 + kw, remaining rows zero.  y fp32 or bf16 NHWC [N,H,W,Co]. */
int vqb_conv2d_fwd_narrowout(const void* x, const void* wp, const float* bias, void* y, int y_dtype, int N, int H, int W, int Ci,
                             int Co, int act, float act_alpha, float gain, void* stream);
/* Narrow-INPUT 3x3 'same' convolution, Ci == 3 (encoder.conv_in 3 -> 128, autoencoder.py:114; LPIPS-VGG conv1_1 3 -> 64): the A operand
 * [128 pixels][K = 27 -> 64] is built in shared memory by producer warps straight from the NHWC image (fp32 or bf16, rounded to bf16) --
 * no 64-channel im2col tensor in HBM.  wp = mode-4 packed weight [Co][64] of vqb_pack_conv_weight; common bias / activation /
 * residual epilogue; Co a multiple of 64. */
int vqb_conv2d_fwd_narrowin(const void* x, int x_dtype, const void* wp, const float* bias, const void* residual, void* y, int y_dtype,
                            int N, int H, int W, int Ci, int Co, int act, float act_alpha, float gain, void* stream);
/* Weight gradient of a 3x3 'same' convolution one side of which has 3 channels (encoder.conv_in: narrow = the image, wide = dy;
 * decoder.conv_out: narrow = dy, wide = the layer input): dwp [64][Cw] fp32 (caller zero-fills), row (tap * 3 + c) with tap = kh * 3 + kw,
 * += sum_pix narrow[n, h + kh - 1, w + kw - 1, c] * wide[n, h, w, cw].  The im2col operand is built in shared memory (no HBM tensor);
 * narrow fp32 or bf16, wide bf16 with Cw a multiple of 128. */
int vqb_conv2d_wgrad_narrow(const void* narrow, int n_dtype, const void* wide, float* dwp, int N, int H, int W, int Cn, int Cw,
                            void* stream);
/* T x T-tap sub-convolution on the 3x3 halo kernels (tcgen05, bf16 operands):
 *     y[n,h,w,co] = act(bias + sum_{a,b<T} sum_ci x[n, h+off+a, w+off+b, ci] * wp[co][(a*T+b)*Ci + ci]),  x zero outside its Hx x Wx pixels,
 * H x W = output size (may differ from the input's), T in {2,3}, -1 <= off, off + T <= 2.  It carries the discriminator's stride-2
 * 3x3 convolution (stylegan2_discriminator/ops/conv2d_resample.py:119-122) without the 4x surplus of a full-resolution evaluation:
 * on the 2x2 space-to-depth form z' [N, H/2+1, W/2+1, 4C] of the FIR-filtered input (written in that layout by vqb_fir4_s2d) the
 * strided convolution is a 2x2-tap stride-1 convolution over 4C channels (T = 2, off = 0: 16C instead of 36C MACs per output); its
 * input gradient is the same call on dy with the tap-flipped, channel-swapped weight (T = 2, off = -1, output (H/2+1) x (W/2+1)).
 * vqb_conv2d_wgrad_sub: dwp[(a*T+b)*Ci + ci][co] (fp32, caller zero-fills) += sum_pix x[n,h+off+a,w+off+b,ci] * dy[n,h,w,co]. */
int vqb_conv2d_sub_supported(int N, int Hx, int Wx, int H, int W, int Ci, int Co, int T, int off);
int vqb_conv2d_fwd_sub(const void* x, const void* wp, const float* bias, const void* residual, void* y, int y_dtype, int N, int Hx,
                       int Wx, int H, int W, int Ci, int Co, int T, int off, int act, float act_alpha, float gain, void* stream);
int vqb_conv2d_wgrad_sub(const void* x, const void* dy, float* dwp, int N, int Hx, int Wx, int H, int W, int Ci, int Co, int T, int off,
                         void* stream);
/* weight gradient: dwp[(kh*KW+kw)*Ci+ci][co] (fp32, mode-0 packed layout) = sum_pix x_shift * dy.
 * dwp must be zero-filled by the caller (split-K partial sums are accumulated with atomics). */
int vqb_conv2d_wgrad(int impl, const void* x, int x_dtype, const void* dy, int dy_dtype, float* dwp,
                     int N, int H, int W, int Ci, int Co, int KH, int KW, int pad, int stride, void* stream);
/* Input gradient of a STRIDED convolution (stride-1 dgrad is vqb_conv2d_fwd on dy with the mode-1/3 packed weight):
 * dx[N,H,W,Ci] from dy[N,OH,OW,Co], OH = (H+2*pad-KH)/stride+1, wd = mode-1 packed weight (fp32 SIMT gather over the
 * virtually zero-upsampled dy).  Used by the StyleGAN2 discriminator's down-sampling convs (conv2d_resample.py:119-122). */
int vqb_conv2d_dgrad(const void* dy, int dy_dtype, const void* wd, void* dx, int dx_dtype, int N, int H, int W, int Ci, int Co,
                     int KH, int KW, int pad, int stride, void* stream);
/* Test / tuning hook for the tcgen05 forward kernel: 0 = generic per-tap TMA loads only, 1 = 3x3 halo reuse (default),
 * 2-4 = descriptor-semantics probes (see csrc/conv_tc.cu); -1 = follow the VQB_HALO_MODE environment variable. */
void vqb_set_halo_mode(int mode);
/* column sums: out[c] (fp32, caller zero-fills) += sum_p a[p][c]; used for bias gradients */
int vqb_colsum(const void* a, int a_dtype, float* out, int64_t P, int C, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * GroupNorm (custom, UNBIASED variance, eps 1e-6: autoencoder.py:25-39) fused with SiLU
 * ---------------------------------------------------------------------------------------------------- */
/* sums[b][g][2] (double, caller zero-fills) += (sum x, sum x^2) over the group's (C/G)*HW elements */
int vqb_gn_stats(const void* x, int x_dtype, double* sums, int N, int HW, int C, int G, void* stream);
/* stats[b][g] = (mean, 1/sqrt(var_unbiased + eps)) */
int vqb_gn_finalize(const double* sums, float* stats, int N, int HW, int C, int G, float eps, void* stream);
/* y = act( (x-mean)*rstd * gamma[c] + beta[c] ), act in {NONE, SILU} */
int vqb_gn_apply(const void* x, int x_dtype, const float* stats, const float* gamma, const float* beta, void* y,
                 int y_dtype, int N, int HW, int C, int G, int act, void* stream);
/* backward pass 1: part[b][c][2] (double, caller zero-fills) += (sum_pix ds, sum_pix ds*xhat),
 * ds = dy * act'(xhat*gamma+beta) */
int vqb_gn_bwd_reduce(const void* x, int x_dtype, const void* dy, int dy_dtype, const float* stats,
                      const float* gamma, const float* beta, double* part, int N, int HW, int C, int G, int act,
                      void* stream);
/* backward finalize: coef[b][g] = (sum_g/n, sum_g_xhat/(n-1)); dgamma[c], dbeta[c] (overwritten) */
int vqb_gn_bwd_finalize(const double* part, const float* gamma, float* coef, float* dgamma, float* dbeta, int N,
                        int HW, int C, int G, void* stream);
/* backward pass 2: dx = rstd * (g - coef0 - xhat*coef1) [+ add], g = ds*gamma.  `add` (dtype of dx, may be NULL) fuses the
 * accumulation of a second gradient of x -- the ResBlock skip connection (autoencoder.py:77) -- into this pass. */
int vqb_gn_bwd_apply(const void* x, int x_dtype, const void* dy, int dy_dtype, const float* stats,
                     const float* gamma, const float* beta, const float* coef, const void* add, void* dx, int dx_dtype,
                     int N, int HW, int C, int G, int act, void* stream);
/* Two-launch forms of the same arithmetic (the finalize kernels folded into their consumers: 4 -> 2 launches per GroupNorm in
 * each direction of autoencoder.py:25-39):
 *   vqb_gn_apply_sums      = vqb_gn_finalize + vqb_gn_apply: mean / rstd are evaluated (in double) from sums[b][g][2] by every
 *                            block for its own channels; stats_out[b][g][2] (may be NULL) receives them for the backward pass.
 *   vqb_gn_bwd_apply_part  = vqb_gn_bwd_finalize + vqb_gn_bwd_apply: coef[b][g] is evaluated from part[b][c][2] by every block;
 *                            block (0,0) also reduces dgamma[c] / dbeta[c] over the batch -- overwritten, or += when
 *                            accumulate_param_grads != 0 (the destinations are then the parameters' .grad views). */
/* The whole backward in ONE cooperative launch (bf16 tensors, VQB_ACT_NONE / SILU): a persistent grid walks the batch image by image,
 * reduces image b, signals a per-image counter and applies image b-1 while its x / dy are still in L2 -- x and dy are read from HBM
 * once instead of twice.  part [N][C][2] double and counters [N] int32 are zero-filled by the caller; dgamma / dbeta overwritten or
 * += (accumulate_param_grads).  max_ctas > 0 caps the grid (data-parallel steps leave SMs to the overlapped NCCL kernels; a
 * cooperative grid waits until all of its CTAs fit).  vqb_gn_bwd_fused_supported: 1 when the problem qualifies (one image's x + dy
 * between 12 and 40 MB, 8-channel vectors), else the caller uses vqb_gn_bwd_reduce + vqb_gn_bwd_apply_part. */
int vqb_gn_bwd_fused_supported(int x_dtype, int dy_dtype, int dx_dtype, int N, int HW, int C, int G, int act);
int vqb_gn_bwd_fused(const void* x, const void* dy, const float* stats, const float* gamma, const float* beta, double* part,
                     int* counters, const void* add, void* dx, float* dgamma, float* dbeta, int accumulate_param_grads, int N,
                     int HW, int C, int G, int act, int max_ctas, void* stream);
int vqb_gn_apply_sums(const void* x, int x_dtype, const double* sums, const float* gamma, const float* beta, void* y, int y_dtype,
                      float* stats_out, int N, int HW, int C, int G, float eps, int act, void* stream);
int vqb_gn_bwd_apply_part(const void* x, int x_dtype, const void* dy, int dy_dtype, const float* stats, const float* gamma,
                          const float* beta, const double* part, const void* add, void* dx, int dx_dtype, float* dgamma,
                          float* dbeta, int accumulate_param_grads, int N, int HW, int C, int G, int act, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * Resampling (Downsample = avg_pool2d(2,2) autoencoder.py:89-91; Upsample = nearest-exact x2 :103-106)
 * ---------------------------------------------------------------------------------------------------- */
/* y[n,h,w,c] = scale * sum_{i,j<2} x[n,2h+i,2w+j,c]   (x is [N,2H,2W,C]; avg-pool: scale=.25; upsample bwd: 1) */
int vqb_down2(const void* x, void* y, int dtype, int N, int H, int W, int C, float scale, void* stream);
/* y[n,h,w,c] = scale * x[n,h/2,w/2,c]                 (y is [N,2H,2W,C]; upsample: scale=1; avg-pool bwd: .25) */
int vqb_up2(const void* x, void* y, int dtype, int N, int H, int W, int C, float scale, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * Elementwise / reductions used by the loss heads (vqvae/model.py:272-275, loss/loss.py:118-121)
 * ---------------------------------------------------------------------------------------------------- */
/* out[0] += sum (a-b)^2 ; out[1] += sum |a-b|   (double[2], caller zero-fills) */
int vqb_diff_sums(const void* a, int a_dtype, const void* b, int b_dtype, double* out, int64_t n, void* stream);
/* SSIM of the test-time evaluation (vqvae/model.py:495, 529-530, 549: torchmetrics StructuralSimilarityIndexMeasure() defaults --
 * 11 x 11 Gaussian window, sigma 1.5, k1 0.01, k2 0.03, valid windows of the un-padded image): per_image_sum[n] += sum over
 * (C, H-10, W-10) of the SSIM map of image n (double[N], caller zero-fills; the image's SSIM is that sum / (C (H-10) (W-10))).
 * preds / target: NCHW fp32; data_range: ONE float in device memory (max(preds.max - preds.min, target.max - target.min)). */
int vqb_ssim_sums(const float* preds, const float* target, const float* data_range, double* per_image_sum, int N, int C, int H,
                  int W, float k1, float k2, void* stream);
/* da = c2 * 2*(a-b) * up2 + c1 * sign(a-b) * up1, (up2, up1) = upstream ? (upstream[0], upstream[1]) : (1, 1)
 * (device scalars: the upstream gradients of the L2 and L1 means); when y_tanh!=0 `a` is a tanh output and the
 * result is additionally multiplied by (1-a^2) (grad wrt the pre-activation) */
int vqb_diff_grad(const void* a, int a_dtype, const void* b, int b_dtype, void* da, int da_dtype, float c1, float c2,
                  const float* upstream, int y_tanh, int64_t n, void* stream);
/* dx = dy * act'(from the saved OUTPUT y): tanh -> 1-y^2 ; used after a conv with a fused tanh epilogue */
int vqb_act_bwd_from_output(const void* y, int y_dtype, const void* dy, int dy_dtype, void* dx, int dx_dtype, int act,
                            float alpha, float gain, int64_t n, void* stream);

/* The same derivative for NHWC tensors [P][C] of ONE dtype with 16-byte vectors, fused with the bias gradient
 * db[c] += sum_p dx[p][c] (db may be NULL; it must be zero-initialised by the caller): bias_act's backward
 * (stylegan2_discriminator/ops/bias_act.py:143-210) in one pass.  C must be a multiple of 8 (bf16) / 4 (fp32). */
int vqb_act_bwd_bias(const void* y, const void* dy, void* dx, int dtype, int act, float alpha, float gain, int64_t P, int C,
                     float* db, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * VQGAN loss heads: StyleGAN2 discriminator resampling, LPIPS, minibatch-stddev (vqvae/modules/loss/)
 * ---------------------------------------------------------------------------------------------------- */
/* depthwise 4x4 FIR [1,3,3,1]x[1,3,3,1]/64 on the zero-padded input, then keep every `down`-th sample:
 * y[n,oh,ow,c] = sum_{a,b} f[a]f[b] x[n, oh*down-pad+a, ow*down-pad+b, c], OH = (H+2*pad-4)/down+1
 * (upfirdn2d with up=1: ops/upfirdn2d.py:120-208, kernels upfirdn2d.cu:97-341).  bwd = its adjoint (dx is [N,H,W,C]). */
int vqb_fir4_fwd(const void* x, void* y, int dtype, int N, int H, int W, int C, int pad, int down, void* stream);
int vqb_fir4_bwd(const void* dy, void* dx, int dtype, int N, int H, int W, int C, int pad, int down, void* stream);
/* The down = 1 filter between a plain NHWC tensor and the 2x2 space-to-depth layout [N][ceil(H/2)][ceil(W/2)][(dy,dx,c)]:
 *   out_s2d: y = s2d(FIR(x, pad)), x [N,H,W,C], OH = H + 2 pad - 3 (OW alike; padding slots of odd sizes are zero-filled);
 *   in_s2d : y [N,OH,OW,C] = FIR(x, pad) of the logical [N,H,W,C] tensor stored in that layout -- with pad' = 3 - pad and
 *            OH = the forward input height this is the adjoint of the first form (upfirdn2d.py:120-208 backward). */
int vqb_fir4_s2d(const void* x, void* y, int dtype, int N, int H, int W, int C, int OH, int OW, int pad, int in_s2d, int out_s2d,
                 void* stream);
/* y[n,oh,ow,:] = x[n,2*oh+off,2*ow+off,:] and its adjoint (zero_upsample2 writes all of x).  A stride-2 3x3 convolution of the
 * discriminator (conv2d_resample.py:119-122) runs in the bf16 fast mode as the stride-1 tcgen05 convolution at full
 * resolution + decimate2(off=1); its backward is zero_upsample2 + the stride-1 dgrad / wgrad kernels. */
int vqb_decimate2(const void* x, void* y, int dtype, int N, int H, int W, int C, int OH, int OW, int off, void* stream);
int vqb_zero_upsample2(const void* y, void* x, int dtype, int N, int H, int W, int C, int OH, int OW, int off, void* stream);
/* 2x2 / stride-2 max-pool (torchvision VGG16 features): y is [N,H,W,C], x is [N,2H,2W,C]; backward routes the
 * gradient to the first maximal element of each window (torch semantics) and writes all of dx */
int vqb_maxpool2_fwd(const void* x, void* y, int dtype, int N, int H, int W, int C, void* stream);
int vqb_maxpool2_bwd(const void* x, int x_dtype, const void* dy, void* dx, int g_dtype, int N, int H, int W, int C, void* stream);
/* 3x3 / stride-2 max-pool without padding (torchvision AlexNet features, the LPIPS 'alex' trunk of loss.py:182):
 * x [N,H,W,C] -> y [N,(H-3)/2+1,(W-3)/2+1,C]; overlapping windows, backward in gather form with torch's first-argmax rule */
int vqb_maxpool3s2_fwd(const void* x, void* y, int dtype, int N, int H, int W, int C, void* stream);
int vqb_maxpool3s2_bwd(const void* x, int x_dtype, const void* dy, void* dx, int g_dtype, int N, int H, int W, int C, void* stream);
/* y[p][c] = x[p][c]*scale[c] + shift[c]  (BaseNet.z_score, lpips_pytorch/modules/networks.py:48-49) */
int vqb_channel_affine(const void* x, int x_dtype, void* y, int y_dtype, const float* scale, const float* shift, int64_t P, int C,
                       void* stream);
/* LPIPS tap (lpips.py:31-38, utils.py:6-8): out[0] (double, caller zero-fills) += sum_pixels sum_c w[c] (u_c - v_c)^2 with
 * u = fx/(|fx|_c + 1e-10), v = fy/(|fy|_c + 1e-10);  bwd: dfy = scale * upstream[0] * d(out)/d(fy) */
int vqb_lpips_tap_fwd(const void* fx, const void* fy, int dtype, const float* w, double* out, int64_t P, int C, void* stream);
int vqb_lpips_tap_bwd(const void* fx, const void* fy, int dtype, const float* w, const float* upstream, float scale, void* dfy,
                      int g_dtype, int64_t P, int C, void* stream);
/* MinibatchStdLayer (discriminator.py:277-293), num_channels=1: y [N][HW][C+1]; stat [N/G] scratch; sample s is in
 * group s % (N/G) exactly as the reference's reshape(G, -1, ...) */
int vqb_mbstd_fwd(const void* x, void* y, float* stat, int dtype, int N, int G, int HW, int C, void* stream);
int vqb_mbstd_bwd(const void* x, int x_dtype, const void* dy, void* dx, int g_dtype, int N, int G, int HW, int C, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * Vector quantisation (vqvae/modules/vector_quantizers.py)
 * ---------------------------------------------------------------------------------------------------- */
/* Fused pairwise-L2 distance -> argmin -> gather -> straight-through value -> loss partial -> code histogram
 * (+ EMA cluster sums).  Replaces VectorQuantizer.forward :23-61, EMAVectorQuantizer.forward :128-180 (the
 * part before the EMA update), EntropyVectorQuantizer.forward :337-349 (argmin/gather), vec_to_codes
 * :63-84,182-203,358-381.
 *   z        [N][D] fp32 (the NHWC latent, 'b c h w -> (b h w) c')
 *   codebook [K][D] fp32
 *   order    0: d = (|z|^2 + |e|^2) - 2 z.e   (:37-39,142-144)   1: d = (|z|^2 - 2 z.e) + |e|^2   (:337-340)
 *   q_out    [N][D] fp32 or NULL: z + (e[idx] - z)  (forward value of the straight-through estimator :58,177)
 *   idx_out  [N] int64 (first minimal index, torch.argmin semantics)
 *   sse      double[1] or NULL, caller zero-fills: += sum (e[idx]-z)^2
 *   counts   [K] float or NULL, caller zero-fills: += histogram of idx (= encodings.sum(0) :160; bincount model.py:290)
 *   dw       [K][D] float or NULL, caller zero-fills: += sum_{i: idx_i=k} z_i  (= encodings.T @ flat_x :166)
 * workspace: vqb_vq_workspace_bytes(N,K,D) bytes. */
size_t vqb_vq_workspace_bytes(int64_t N, int K, int D);
int vqb_vq_assign(const float* z, const float* codebook, int order, float* q_out, int64_t* idx_out, double* sse,
                  float* counts, float* dw, int64_t N, int K, int D, void* workspace, size_t workspace_bytes,
                  void* stream);
/* The same contract as vqb_vq_assign (identical indices: first-index fp32 argmin in the reference's operation order), with
 * the distance GEMM on the tensor cores: bf16 hi+lo split operands, three tcgen05 UMMAs per k-step, per-row best / second
 * best in the epilogue; rows whose gap exceeds a rigorous error bound are decided there, the rest (genuine near-ties) are
 * re-evaluated by the exact fp32 kernel.  Needs D % 64 == 0, D <= 256, K % 8 == 0.  undecided_rows_out (device int, may be
 * NULL) receives the number of rows that took the exact path.  workspace: vqb_vq_tc_workspace_bytes(N,K,D). */
size_t vqb_vq_tc_workspace_bytes(int64_t N, int K, int D);
int vqb_vq_assign_tc(const float* z, const float* codebook, int order, float* q_out, int64_t* idx_out, double* sse,
                     float* counts, float* dw, int64_t N, int K, int D, void* workspace, size_t workspace_bytes,
                     int* undecided_rows_out, void* stream);
/* ONE-launch form of the same contract (identical indices to vqb_vq_assign), the product path on sm_100: a cluster of two
 * CTAs owns 256 latent rows; z is read once from HBM and rounded to fp16 in the prologue (resident in shared memory), the
 * fp16 copy of the codebook streams through a TMA ring, tcgen05 cta_group::2 UMMAs give the dot products in TMEM with a
 * rigorously bounded error PER CODE (it grows with the code's norm), the scan warps keep a running upper bound of the smallest
 * distance plus an online list of every code whose lower bound does not exceed it, near-tied rows re-evaluate ONLY their
 * surviving candidates with the strict kernel's fp32 arithmetic, and the same launch writes idx, q, sum (e-z)^2, the histogram
 * and the EMA cluster sums.  Replaces vector_quantizers.py:37-61, 142-166.
 *   cb_half [K][D] fp16 and cb_sq [4 K] fp32 (|e_k|^2, then the per-code and per-32-code-chunk coefficients of the error
 *   bound): produced by vqb_vq_prep_codebook (any codebook) or by vqb_vq_ema_update_prep (the EMA path: the update kernel
 *   leaves them for the NEW codebook, so the next step needs no preparation launch).
 * Needs D % 64 == 0, D <= 256, K % 8 == 0, K <= 65528.  sse / counts / dw / undecided_rows_out: caller zero-fills;
 * undecided_rows_out (may be NULL) is int[2]: rows that took the exact re-rank, and those among them that scanned every code. */
int vqb_vq_prep_codebook(const float* codebook, void* cb_half, float* cb_sq, int K, int D, void* stream);
int vqb_vq_fused(const float* z, const float* codebook, const void* cb_half, const float* cb_sq, int order, float* q_out,
                 int64_t* idx_out, double* sse, float* counts, float* dw, int64_t N, int K, int D, int* undecided_rows_out,
                 void* stream);
/* Tuning hook: device buffer [ctas][8] int64 that subsequent vqb_vq_fused launches fill with globaltimer stamps of their phase
 * boundaries (start, prologue, scan, decide, re-rank, finish); NULL (default) = off. */
void vqb_vq_fused_set_trace(void* dev_buf);
/* vqb_vq_ema_update (below) that additionally writes cb_half / cb_sq of the updated codebook. */
int vqb_vq_ema_update_prep(float* ema_count, float* ema_weight, float* codebook, const float* counts, const float* dw,
                           void* cb_half, float* cb_sq, int K, int D, float decay, float eps, float batch, void* stream);
/* EMA codebook update (vector_quantizers.py:158-169), in place:
 *   c = decay*ema_count + (1-decay)*counts ; ema_count = (c+eps)/(b + K*eps)*b   (b = IMAGE batch: defect B7)
 *   ema_weight = decay*ema_weight + (1-decay)*dw ; codebook = ema_weight / ema_count[:,None] */
int vqb_vq_ema_update(float* ema_count, float* ema_weight, float* codebook, const float* counts, const float* dw,
                      int K, int D, float decay, float eps, float batch, void* stream);
/* Backward of the standard/EMA/entropy commitment + codebook MSE terms.  `q` holds the
 * per-row quantised vectors e[idx_i] as produced by vqb_vq_assign (q_out), so the EMA path may already have
 * overwritten the codebook (vector_quantizers.py:169 runs before backward):
 *   dz[i]  = g_q[i] + (g_loss[0] * 2*beta/(N*D)) * (z[i]-q[i])             (g_q may be NULL -> 0)
 *   dcb[k] += (g_loss[0] * 2*cb_scale/(N*D)) * sum_{i:idx_i=k} (q[i]-z[i])   (dcb NULL for EMA; caller zero-fills) */
int vqb_vq_backward(const float* z, const float* q, const int64_t* idx, const float* g_q, const float* g_loss,
                    float beta, float cb_scale, float* dz, float* dcb, int64_t N, int K, int D, void* stream);
/* out[r] = sum_d a[r][d]^2  (|e_k|^2 of the codebook rows) */
int vqb_row_sqnorm(const float* a, float* out, int64_t R, int D, void* stream);

/* Entropy quantizer (EntropyVectorQuantizer.forward, vector_quantizers.py:290-356), organised around the N x K matrix
 * that the implicit-GEMM kernel produces (dot = flat_x @ codebook^T as a 1x1 convolution):
 *   entropy_rows : dot[N][K] is overwritten IN PLACE by logp = log_softmax(-d/T), d = (|z|^2 - 2 dot) + |e|^2 (:337-340);
 *                  idx_out = argmin d (first index); *sample_entropy_sum (double, caller zero-fills) += -sum_k p log p per row
 *   colsum_exp   : out[k] (caller zero-fills) += sum_rows exp(logp[row][k])           (N * avg_probs, :320)
 *   finalize     : out[0] = ratio * (sample_entropy_sum/N - avg_entropy), out[1] = avg_entropy = -sum m log(m+1e-5) (:321-328)
 *   bwd_rows     : logp[N][K] is overwritten IN PLACE by G = dLoss/dd; dz += -2 G E and
 *                  dE += 2 E colsum(G) - 2 G^T Z are then 1x1-convolution dgrad/wgrad calls + combine_dcb */
int vqb_vq_entropy_rows(float* dot_to_logp, const float* z, const float* codebook_sq, float temperature, int64_t* idx_out,
                        double* sample_entropy_sum, int64_t N, int K, int D, int argmax_target, void* stream);
int vqb_vq_colsum_exp(const float* logp, float* out, int64_t N, int K, void* stream);
int vqb_vq_entropy_finalize(const float* colsum_p, const double* sample_entropy_sum, float ratio, float* out, int64_t N, int K,
                            void* stream);
int vqb_vq_entropy_bwd_rows(float* logp_to_g, const float* colsum_p, const float* g_loss, float ratio, float temperature,
                            int64_t N, int K, const int64_t* argmax_idx, void* stream);
int vqb_vq_entropy_combine_dcb(float* dcb, const float* codebook, const float* colsum_g, const float* gtz, int K, int D,
                               void* stream);

/* Gumbel-softmax quantizer rows (GumbelVectorQuantizer.forward, vector_quantizers.py:223-245; F.gumbel_softmax):
 *   fwd: y[row] = softmax((logits - log(exp_noise)) / tau)  (hard != 0: one-hot of its argmax), idx = argmax,
 *        *kl_sum (double, caller zero-fills) += sum_n qy log(qy K + 1e-10), qy = softmax(logits).
 *        exp_noise holds the Exp(1) samples F.gumbel_softmax draws (explicit so CPU and GPU consume identical noise;
 *        NULL = no noise).
 *   bwd: dlogits = (1/tau) soft (dy - sum soft dy) + g_kl[0] * kl_scale * qy (f - sum qy f) */
int vqb_gumbel_rows_fwd(const float* logits, const float* exp_noise, float tau, int hard, float* y, int64_t* idx_out,
                        double* kl_sum, int64_t N, int K, void* stream);
int vqb_gumbel_rows_bwd(const float* logits, const float* exp_noise, float tau, const float* dy, const float* g_kl,
                        float kl_scale, float* dlogits, int64_t N, int K, void* stream);
/* vqb_gumbel_rows_fwd / _bwd with the temperature read from device memory (tau_dev[0]): CUDA-graph replay under the
 * temperature schedule of model.py:219-225. */
int vqb_gumbel_rows_fwd_dev(const float* logits, const float* exp_noise, const float* tau_dev, int hard, float* y,
                            int64_t* idx_out, double* kl_sum, int64_t N, int K, void* stream);
int vqb_gumbel_rows_bwd_dev(const float* logits, const float* exp_noise, const float* tau_dev, const float* dy,
                            const float* g_kl, float kl_scale, float* dlogits, int64_t N, int K, void* stream);

/* gather rows: out[i] = codebook[idx[i]] (BaseVectorQuantizer.codes_to_vec base_quantizer.py:53-61) */
int vqb_vq_gather(const float* codebook, const int64_t* idx, float* out, int64_t N, int K, int D, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * Optimizer (torch.optim.AdamW as configured by VQVAE.configure_optimizers, vqvae/model.py:411-438)
 * ---------------------------------------------------------------------------------------------------- */
/* One fused pass over a flat fp32 range: decoupled weight decay, bias-corrected Adam.  grad_scale multiplies
 * the gradient first (1/world_size after a sum all-reduce). */
int vqb_adamw(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
              float weight_decay, int step, float grad_scale, void* stream);
/* The same update with the per-step scalars in DEVICE memory: hyper = {lr, 1 - beta1^step, sqrt(1 - beta2^step)}.  The launch
 * holds no step-dependent host value, so a whole training step can be captured in a CUDA graph and replayed while the host
 * rewrites three floats per optimizer group (lr schedule of on_train_batch_start, model.py:202-218). */
int vqb_adamw_dev(float* p, const float* g, float* m, float* v, int64_t n, const float* hyper, float beta1, float beta2,
                  float eps, float weight_decay, float grad_scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VQGAN_B200_H */
