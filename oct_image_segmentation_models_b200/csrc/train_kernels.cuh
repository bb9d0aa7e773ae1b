// Launchers of the training kernels (csrc/train_kernels.cu).
#pragma once
#include "common.cuh"

namespace octseg {

// per-step scalars that live on the device so that a captured CUDA graph of the train step replays correctly
struct StepState {
  unsigned long long step;
  float lr_t;
  float pad;
};
int launch_step_advance(StepState *s, float lr, float b1, float b2, cudaStream_t st);

// per-channel sum / sum of squares of z over all N*H*W pixels (double accumulators, [2*C])
template <typename T>
int launch_bn_stats(View<const T> z, double *sums, cudaStream_t st);

// sums -> batch mean / invstd, affine (scale = gamma*invstd, shift = beta - mean*scale), and the
// Keras moving-statistics update (momentum 0.99, Bessel-corrected variance)
int launch_bn_finalize(const double *sums, long long count, int c, float eps, float momentum,
                       const float *gamma, const float *beta, float *moving_mean, float *moving_var,
                       float *mean, float *invstd, float *scale, float *shift, cudaStream_t st);

// bn_finalize + bn_apply_relu (+ 2x2 max-pool when pooled.ptr != NULL) in one pass over z: see the kernel
template <typename T>
int launch_bn_finalize_apply(View<const T> z, const double *sums, long long count, float eps, float momentum,
                             const float *gamma, const float *beta, float *moving_mean, float *moving_var, float *mean,
                             float *invstd, float *scale, float *shift, const T *mask, View<T> a, View<T> pooled,
                             cudaStream_t st);

// a = relu(z*scale + shift) [* mask]   (mask: optional multiplier tensor, same layout as z)
template <typename T>
int launch_bn_apply_relu(View<const T> z, const float *scale, const float *shift, const T *mask,
                         View<T> a, cudaStream_t st);

// dropout multiplier tensor (0 or 1/(1-rate)) in blocked layout, from an injected NHWC uint8
// mask or from a counter-based hash RNG
template <typename T>
int launch_dropout_mask(const uint8_t *mask_nhwc, unsigned long long seed, const StepState *state, float rate,
                        int n, int c, int h, int w, T *out, cudaStream_t st);

// fused head: 1x1 conv + softmax + weighted CE + dlogits + d(head weights) + d(input)
template <typename T>
int launch_head_loss(View<const T> a, const float *wgt, const float *bias, int cin, int K,
                     const uint8_t *labels, const float *class_w, float inv_denominator, View<T> da,
                     float *d_wgt, float *d_bias, double *loss_acc, float *dlog_scratch /* [n*h*w][K], heads wider than 16 channels */,
                     cudaStream_t st);

// BN+ReLU backward, pass 1: sums of dy and dy*zhat per channel ([2*C] doubles)
template <typename T>
int launch_bn_bwd_reduce(View<const T> da, View<const T> z, const float *mean, const float *invstd,
                         const float *gamma, const float *beta, const T *mask, double *sums,
                         cudaStream_t st);
// pass 2: dz = gamma*invstd*(dy - sum_dy/M - zhat*sum_dyz/M); also emits dgamma/dbeta (fp32)
template <typename T>
int launch_bn_bwd_apply(View<const T> da, View<const T> z, const float *mean, const float *invstd,
                        const float *gamma, const float *beta, const T *mask, const double *sums,
                        long long count, View<T> dz, float *d_gamma, float *d_beta, cudaStream_t st);

// da_total = d_skip + scatter of the pooled gradient to the first max of each 2x2 window
template <typename T>
int launch_pool_bwd_add(View<const T> a, View<const T> d_pooled, View<const T> d_skip, View<T> da_total,
                        cudaStream_t st);
// same + the BatchNorm-backward reductions (sum dy, sum dy * zhat) of the block, see pool_bwd_add_kernel
template <typename T>
int launch_pool_bwd_add_bnred(View<const T> a, View<const T> d_pooled, View<const T> d_skip, View<T> da_total, View<const T> z,
                              const float *mean, const float *invstd, const float *gamma, const float *beta, double *sums,
                              cudaStream_t st);

// nearest x2 up-sampling into a dense tensor (weight gradient of the wide up-conv only)
template <typename T>
int launch_upsample2x(View<const T> in, View<T> out, cudaStream_t st);

// weight gradient (accumulates with atomics into zero-initialised dW [kh][kw][cin][cout], db [cout])
template <typename T>
int launch_wgrad(View<const T> a_in, View<const T> dz, int kh, int kw, int pad_top, int pad_left,
                 int ups, int cin, int cout, float *dW, float *db, cudaStream_t st);

// tensor-core (mma.sync bf16) weight gradient, same contract as launch_wgrad (csrc/wgrad_mma.cu)
int launch_wgrad_mma(View<const __nv_bfloat16> a_in, View<const __nv_bfloat16> dz, int kh, int kw, int pad_top,
                     int pad_left, int ups, int cin, int cout, float *dW, float *db, int *status,
                     cudaStream_t st);

// narrow layers (Cin <= 16) run the row-walking kernel, which also reads the LOW-res input of the up-conv directly
// (ups = 1, 2x2 kernel): the caller does not materialise the up-sampled tensor when this returns true
bool wgrad_rows_applicable(int kh, int kw, int cin, int ups);

// tcgen05 weight gradient for dense layers with >= 64 input channels (csrc/wgrad_tc.cu): MN-major UMMA operands straight
// from the blocked tiles, accumulators in TMEM, deterministic split-K reduction through `scratch`
bool wgrad_tc_applicable(int kh, int kw, int cin, int cout, int ups, int h, int w);
size_t wgrad_tc_scratch_floats(int kh, int kw, int cin, int cout, int n, int h, int w);
int launch_wgrad_tc(View<const __nv_bfloat16> a_in, View<const __nv_bfloat16> dz, int kh, int kw, int pad_top, int pad_left,
                    int cin, int cout, float *dW, float *db, float *scratch, size_t scratch_floats, int *status,
                    cudaStream_t st);

// raw image -> blocked T tensor with 8 channels (channel c < cin = x/255, rest 0)
template <typename T>
int launch_image_to_blocked(const void *img, int img_dtype, int n, int h, int w, int cin, T *out,
                            cudaStream_t st);

// weight transforms for the data gradient
//  flip: out[kh-1-a][kw-1-b][co][ci] = w[a][b][ci][co]
int launch_flip_transpose(const float *w, int kh, int kw, int cin, int cout, float *out, cudaStream_t st);
//  up-conv dgrad kernel ((kh+1)x(kw+1), stride 2): see train.cu
int launch_upconv_dgrad_weights(const float *w, int kh, int kw, int cin, int cout, float *out,
                                cudaStream_t st);
// copy the ci==0 slice of a [taps][8][cout] buffer into [taps][cin][cout] (stem wgrad)
int launch_stem_wgrad_extract(const float *tmp, int taps, int cin, int cout, float *dW, cudaStream_t st);

// Keras optimizer_v2 Adam over the flat parameter buffer
int launch_adam(float *p, const float *g, float *m, float *v, long long n, const StepState *state, float b1,
                float b2, float eps, cudaStream_t st);

}  // namespace octseg
