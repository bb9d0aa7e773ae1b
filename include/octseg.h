/*
 * octseg.h -- C ABI of the B200-native U-Net hot path (liboctseg.so).
 *
 * The reference (NIH-NEI/oct-image-segmentation-models) has no FFI of its own: every
 * FLOP of this path runs inside TensorFlow, reached through Keras objects.  Each entry
 * point below therefore cites the reference *call site* it replaces; the Python-side
 * mirror (oct_image_segmentation_models_b200/) binds these with ctypes and keeps the
 * reference's Python surface (load_model_and_config / predict / evaluate_model /
 * train_model / *Params) unchanged.  See INTEGRATION.md for the binding stub.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / CUDA types in signatures
 *     (`stream` is a cudaStream_t passed as void*, NULL = the handle's own stream);
 *   - every call returns 0 on success, non-zero on failure; octseg_last_error()
 *     returns a thread-local message for the last failure;
 *   - a handle is bound to one CUDA device and is not thread-safe (one host thread
 *     per GPU, as the reference's single-threaded loops are);
 *   - host tensors are NHWC, row-major, owned by the caller.  On the device the
 *     library keeps activations in an internal channel-blocked layout
 *     [N][C/8][H][W][8]; callers never see it.
 *   - there is NO CPU fallback: without a CUDA device every compute call fails.
 */
#ifndef OCTSEG_H_
#define OCTSEG_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct octseg_net octseg_net;

/* Graph hyper-parameters == keys of the reference's model_config.json
 * (reference: models/unet.py:62-104, models/base_model.py:27-33). */
typedef struct octseg_config {
  int32_t input_channels;
  int32_t num_classes;
  int32_t start_neurons;   /* default 8  */
  int32_t pool_layers;     /* default 4  */
  int32_t conv_layers;     /* default 2  */
  int32_t enc_kh, enc_kw;  /* default 3,3 */
  int32_t dec_kh, dec_kw;  /* default 2,2 */
} octseg_config;

/* precision modes: fp32 (CUDA cores, 1e-4 gate), bf16 (tcgen05, train + predict), fp16 (tcgen05, predict
 * only: same speed as bf16, 8x finer mantissa -> argmax / boundary fidelity on confident trained nets) */
enum { OCTSEG_FP32 = 0, OCTSEG_BF16 = 1, OCTSEG_FP16 = 2 };
/* image dtypes: raw uint8, raw float32 (0..255, x/255 applied on the device), or float32 that the
 * caller already preprocessed with the reference's x/255.0 (models/unet.py:87-91) */
enum { OCTSEG_U8 = 0, OCTSEG_F32 = 1, OCTSEG_F32_PRE = 2 };

/* ---- library / device ---------------------------------------------------------- */
int32_t     octseg_version(void);
const char *octseg_last_error(void);
/* number of visible CUDA devices; 0 (not an error) when there is no driver/GPU */
int32_t     octseg_device_count(void);

/* ---- static graph description (no GPU needed) ----------------------------------- *
 * Replaces introspection of the Keras model built by UNet.build_model()
 * (reference models/unet.py:106-153): weight tensors in `model.get_weights()` order. */
int32_t octseg_param_count(const octseg_config *cfg, int32_t *n_tensors);
int32_t octseg_param_info(const octseg_config *cfg, int32_t index, char *name, int32_t name_cap,
                          int32_t *ndim, int64_t shape[4], int32_t *trainable);

/* ---- lifetime ------------------------------------------------------------------- *
 * Replaces `model_class(**config).build_model()` (reference training/training.py:243-260)
 * and the model object returned by tf.keras.models.load_model (common/utils.py:63-67). */
int32_t octseg_create(const octseg_config *cfg, int32_t device, int32_t precision, octseg_net **out);
int32_t octseg_destroy(octseg_net *net);

/* ---- weights (Keras get_weights()/set_weights() order, float32 host arrays) ------ */
int32_t octseg_set_param(octseg_net *net, int32_t index, const float *host, int64_t count);
int32_t octseg_get_param(octseg_net *net, int32_t index, float *host, int64_t count);

/* ---- inference ------------------------------------------------------------------ *
 * Replaces `loaded_model.predict(preprocess(x))`
 * (reference prediction/prediction.py:75-81, evaluation/evaluation.py:129-135, with
 * the x/255 preprocessing of models/unet.py:87-91 applied on the device).
 *   images : [n,h,w,input_channels], dtype OCTSEG_U8 / OCTSEG_F32 (raw 0..255) / OCTSEG_F32_PRE
 *   probs  : [n,h,w,num_classes] float32 softmax output, may be NULL
 *   labels : [n,h,w] uint8 argmax (first max on ties, as np.argmax), may be NULL
 * *_host: pageable or pinned host pointers, copies are part of the call: H2D, forward and D2H run as a chunked
 *   three-stream pipeline (pinned buffers reach PCIe speed; uint8 input takes the tensor-core stem).
 * *_device: device pointers on the handle's device; asynchronous on `stream`. */
int32_t octseg_predict_host(octseg_net *net, const void *images, int32_t dtype, int32_t n, int32_t h,
                            int32_t w, float *probs, uint8_t *labels);
int32_t octseg_predict_device(octseg_net *net, const void *images, int32_t dtype, int32_t n, int32_t h,
                              int32_t w, float *probs, uint8_t *labels, void *stream);
/* predict + argmax + boundary probability maps on the device: replaces
 * perform_argmax(bin=True) + convert_predictions_to_maps_semantic (reference common/utils.py:80-168),
 * which depend on the label map only.  maps: uint8 [n, K-1, h, w], or [n, K-1, w, h] when
 * `transposed` (the layout graph_search.segment_maps wants); labels may be NULL. */
int32_t octseg_predict_maps_host(octseg_net *net, const void *images, int32_t dtype, int32_t n, int32_t h,
                                 int32_t w, int32_t bg_ilm, int32_t bg_csi, int32_t transposed, uint8_t *labels,
                                 uint8_t *maps);
/* asynchronous pair of octseg_predict_maps_host for pipelining consecutive batches (reference loop: one predict per
 * image, prediction/prediction.py:70-81 -- here whole batches in flight): submit() only enqueues the H2D copies, the
 * forward passes and the D2H copies and returns a ticket; wait(ticket) completes that call.  Host buffers must be
 * PINNED (cudaHostAlloc / torch pin_memory) and stay valid until the matching wait.  Up to two calls may be in
 * flight: submit(i+1) before wait(i) overlaps the H2D of batch i+1 with the forward and the D2H of batch i. */
int32_t octseg_predict_maps_submit(octseg_net *net, const void *images, int32_t dtype, int32_t n, int32_t h, int32_t w,
                                   int32_t bg_ilm, int32_t bg_csi, int32_t transposed, uint8_t *labels, uint8_t *maps,
                                   int32_t *ticket);
int32_t octseg_predict_wait(octseg_net *net, int32_t ticket);
/* validation pass on the device: replaces model.predict over the validation Sequence +
 * the Dice monitor metrics (reference common/custom_metrics.py:19-77) + the validation loss of
 * weighted_categorical_crossentropy (common/custom_losses.py:27-35).  labels: true class ids u8 [n,h,w].
 *   counts   : int64 [n][K][3] = per image and class |{y == c and p_c > 0.5}|, |{p_c > 0.5}|, |{y == c}|
 *   loss_sums: double [n] = per-image sum over pixels of -w_y * log(clip(p_y / sum p, 1e-7, 1 - 1e-7))
 * class_weights: [K] or NULL (= all ones). */
int32_t octseg_evaluate_host(octseg_net *net, const void *images, int32_t dtype, const uint8_t *labels, int32_t n,
                             int32_t h, int32_t w, const float *class_weights, int64_t *counts, double *loss_sums);
/* wait for all work queued on the handle's stream */
int32_t octseg_synchronize(octseg_net *net);

/* ---- training ------------------------------------------------------------------- *
 * Replaces one step of `model.fit` after `model.compile(optimizer, loss)`
 * (reference training/training.py:262-266, 401-407): forward with batch-statistics
 * BatchNorm and Dropout(0.5), weighted categorical cross-entropy
 * (common/custom_losses.py:27-35, mean over all pixels of the GLOBAL batch),
 * backward, gradient all-reduce across ranks (MirroredStrategy, training.py:185),
 * Keras-Adam update, BN moving-statistics update. */
typedef struct octseg_train_config {
  float   learning_rate;     /* 1e-3  */
  float   beta_1, beta_2;    /* 0.9, 0.999 */
  float   epsilon;           /* 1e-7 (outside the bias correction, Keras optimizer_v2) */
  float   dropout_rate;      /* 0.5; 0 disables */
  uint64_t dropout_seed;
  int32_t global_batch;      /* samples over all ranks (loss denominator) */
} octseg_train_config;

int32_t octseg_train_begin(octseg_net *net, const octseg_train_config *tc,
                           const float *class_weights /* [num_classes] */);
/* NCCL communicator for data-parallel training: `unique_id` is the 128-byte
 * ncclUniqueId produced by octseg_comm_unique_id on rank 0 and broadcast by the host. */
int32_t octseg_comm_unique_id(uint8_t id_out[128]);
int32_t octseg_comm_init(octseg_net *net, const uint8_t unique_id[128], int32_t rank, int32_t world);
/* images: [n,h,w,Cin] (this rank's shard), labels: [n,h,w] uint8 class ids,
 * dropout_mask: NULL or [n, h/2^P, w/2^P, C_mid] uint8 {0,1} (injected for parity tests).
 * loss_out receives this rank's contribution sum(per_pixel)/(global_batch*h*w).
 * From the second call with identical arguments (same buffers, shape, stream) the step is replayed from a captured
 * CUDA graph; the step counter, Adam's bias-corrected rate and the dropout stream position live on the device. */
int32_t octseg_train_step_host(octseg_net *net, const void *images, int32_t dtype, const uint8_t *labels,
                               int32_t n, int32_t h, int32_t w, const uint8_t *dropout_mask,
                               float *loss_out);
int32_t octseg_train_step_device(octseg_net *net, const void *images, int32_t dtype, const uint8_t *labels,
                                 int32_t n, int32_t h, int32_t w, const uint8_t *dropout_mask,
                                 float *loss_out_device, void *stream);
/* optimizer state, for checkpoints that resume like the reference's ModelCheckpoint files (training.py:319-326):
 * which = 0 Adam m, 1 Adam v (parameter order and shapes); set != 0 writes, else reads */
int32_t octseg_opt_state(octseg_net *net, int32_t set, int32_t which, int32_t index, float *host, int64_t count);
int32_t octseg_opt_iterations(octseg_net *net, int32_t set, int64_t *iterations);
/* flat float32 gradient of the last step, Keras trainable-weight order (tests) */
int32_t octseg_get_grad(octseg_net *net, int32_t index, float *host, int64_t count);

/* ---- boundary extraction (host code) ------------------------------ *
 * Replaces graph_search.segment_maps (reference min_path_processing/graph_search.py:519-572 and
 * the Dijkstra at :5-105, graph at :108-225) with identical tie-breaking.
 *   maps_t  : [n_maps][width][height] uint8 boundary maps, transposed as the reference callers do
 *             (prediction/prediction.py:134-135)
 *   rows_out: [n_maps][width] uint16 boundary row per column
 *   n_threads: maps are independent; <=1 = serial */
int32_t octseg_min_path_segment(const uint8_t *maps_t, int32_t n_maps, int32_t width, int32_t height,
                                uint16_t *rows_out, int32_t n_threads);

/* ---- introspection for bench / tests -------------------------------------------- */
/* number of kernels this library launched on the handle since creation */
int64_t octseg_launch_count(octseg_net *net);
/* per-block CUDA-event timing of forward passes (bench roofline): when enabled every conv block
 * (and its pool / the head) is bracketed by events on the launching stream */
int32_t octseg_set_profiling(octseg_net *net, int32_t enable);
/* milliseconds per block of the LAST forward micro-batch; n_blocks receives the block count */
int32_t octseg_get_block_times(octseg_net *net, float *ms, int32_t cap, int32_t *n_blocks);
/* 1 if conv layer `conv_index` runs on the tcgen05 path at (h,w) in the current mode */
int32_t octseg_layer_uses_tensor_core(octseg_net *net, int32_t conv_index, int32_t h, int32_t w);
/* run ONE conv block (index in Keras order) on caller data, for per-layer parity tests:
 *   in  : float32 NHWC [n,h,w,cin] host;  out: float32 NHWC host [n,h_out,w_out,cout]
 *   path: 0 = CUDA-core kernel, 1 = tcgen05 kernel (bf16 mode only) */
int32_t octseg_debug_conv_block(octseg_net *net, int32_t conv_index, int32_t path, const float *in,
                                int32_t n, int32_t h, int32_t w, float *out, float *ms_out);

/* run the BACKWARD kernels of one conv block (index 1 .. last conv block) on caller data, for per-layer parity tests
 * against float64 math (SURVEY section 4 "unit (GPU)"): the weight-gradient kernel and the data-gradient convolution the
 * train step launches in the handle's precision (bf16: tensor cores; fp32: CUDA cores).  Needs octseg_train_begin.
 *   a_in: float32 NHWC [n,h,w,cin] input activation;  dz: float32 NHWC [n,h_out,w_out,cout] gradient wrt the conv output
 *   dW: [kh,kw,cin,cout], db: [cout], d_in: float32 NHWC [n,h,w,cin] */
int32_t octseg_debug_backward_block(octseg_net *net, int32_t conv_index, const float *a_in, const float *dz, int32_t n,
                                    int32_t h, int32_t w, float *dW, float *db, float *d_in);

#ifdef __cplusplus
}
#endif
#endif /* OCTSEG_H_ */
