// Launchers of the CUDA-core kernels (memory-bound ops + the fp32-exact conv path).
#pragma once
#include "common.cuh"

namespace octseg {

// x/255 preprocessing table: lut[u] = float32(float64(u)/255)  (reference models/unet.py:87-91)
int init_preprocess_lut();

// First conv block: reads the raw NHWC image (u8 or f32), applies x/255, kxk "same" conv,
// epilogue y = acc*scale + shift (+ReLU), writes blocked output.
template <typename T>
int launch_conv_first(const void *img, int img_dtype, int n, int h, int w, int cin_img,
                      const float *wgt, int kh, int kw, int cout, const float *scale,
                      const float *shift, int relu, View<T> out, cudaStream_t st);

// Generic kxk "same" conv on blocked input (CUDA cores, fp32 accumulate).
// ups=1: the input is first nearest-upsampled x2 (never materialised): reference
// models/unet.py:41-44 (UpSampling2D + Conv2D(dec_kernel, "same")).
template <typename T>
int launch_conv_direct(View<const T> in, const float *wgt, int kh, int kw, int cin, int cout,
                       int ups, const float *scale, const float *shift, int relu, View<T> out,
                       cudaStream_t st);

// Same kernel with explicit geometry: output pixel (y,x) reads input (y*stride+dy-pad_top,
// x*stride+dx-pad_left).  Used by the backward pass (dgrad = conv with transformed weights;
// the up-conv's dgrad is a stride-2 3x3 conv over dz).
template <typename T>
int launch_conv_direct_ex(View<const T> in, const float *wgt, int kh, int kw, int cin, int cout,
                          int ups, int stride, int pad_top, int pad_left, const float *scale,
                          const float *shift, int relu, View<T> out, cudaStream_t st);

// 2x2/stride-2 max pool (reference models/unet.py:37)
// TC stem helpers: widen a uint8 image to the activation type; replicate folded BN per GEMM column
template <typename T>
int launch_u8_to_act(const uint8_t *img, long long count, T *out, cudaStream_t st);
int launch_stem_rep(const float *scale, const float *shift, int cout, float *rep_scale, float *rep_shift, cudaStream_t st);

template <typename T>
int launch_maxpool2(View<const T> in, View<T> out, cudaStream_t st);

// 1x1 conv + bias + softmax head (reference models/unet.py:142-147); probs NHWC fp32,
// labels = first-max argmax (np.argmax semantics, reference common/utils.py:104)
template <typename T>
int launch_head(View<const T> in, const float *wgt /*[cin][K]*/, const float *bias, int cin, int K,
                float *probs, uint8_t *labels, cudaStream_t st);

// u8 labels [n,h,w] -> u8 boundary maps [n,K-1,h,w] (or [n,K-1,w,h] when transposed), reference semantics
int launch_boundary_maps(const uint8_t *labels, int n, int h, int w, int K, int bg_ilm, int bg_csi, int transposed,
                         uint8_t *maps, cudaStream_t st);

// Validation metrics on the device (reference common/custom_metrics.py:19-77 and
// common/custom_losses.py:27-35): per image and class, |{y == c and p_c > 0.5}|, |{p_c > 0.5}|, |{y == c}| --
// the three sums thresholded Dice (micro and macro) is made of -- and the per-image sum of the weighted
// categorical cross-entropy.  probs NHWC fp32 [n,h,w,K], labels u8 [n,h,w]; counts [n][K][3], loss [n] (zeroed by the caller).
int launch_eval_counts(const float *probs, const uint8_t *labels, int n, int h, int w, int K, const float *class_w,
                       unsigned long long *counts, double *loss, cudaStream_t st);

// BN folding for inference: scale = gamma*rsqrt(var+eps), shift = (bias-mean)*scale+beta
int launch_bn_fold(const float *bias, const float *gamma, const float *beta, const float *mean,
                   const float *var, float eps, int c, float *scale, float *shift, cudaStream_t st);

}  // namespace octseg
