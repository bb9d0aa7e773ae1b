// Internal network object behind the C ABI.
#pragma once
#include <map>
#include <string>
#include <vector>

#include "../../include/octseg.h"
#include "common.cuh"
#include "conv_tc.cuh"

namespace octseg {

struct BlockSpec {
  int index = 0;
  int role = 0;  // 0 enc, 1 mid, 2 up, 3 dec, 4 head
  int level = 0;
  int kh = 0, kw = 0, cin = 0, cout = 0;
  bool has_bn = true, pool_after = false, ups = false, dropout_after = false;
  int concat_level = -1;
  int conv_j = 0;  // position inside its group of conv_layers
  // parameter table indices (Keras order): kernel, bias, gamma, beta, mean, var
  int p_kernel = -1, p_bias = -1, p_gamma = -1, p_beta = -1, p_mean = -1, p_var = -1;
};

struct ParamSpec {
  std::string name;
  int ndim = 0;
  int64_t shape[4] = {0, 0, 0, 0};
  int64_t count = 0;
  int64_t offset = 0;   // floats into the flat buffer (16-float aligned)
  bool trainable = true;
  int block = 0;
};

int build_graph(const octseg_config &cfg, std::vector<BlockSpec> *blocks, std::vector<ParamSpec> *params,
                int64_t *total_floats);

// Per-block device state
struct BlockState {
  float *scale = nullptr, *shift = nullptr;   // folded BN (inference) [cout]
  bool geo_ok = false;
  TcGeometry geo{};
  __nv_bfloat16 *wpack = nullptr;
  size_t wpack_elems = 0;
  // inference-only row-pair variant of narrow 3x3 layers (cout 8 / 16): see tc_rowpair_weights
  bool geo2_ok = false;
  TcGeometry geo2{};
  __nv_bfloat16 *wpack2 = nullptr;
  size_t wpack2_elems = 0;
  float *rep_scale = nullptr, *rep_shift = nullptr;   // TC stem only: per-GEMM-column scale/255 and shift
  // fp32 mode on tensor cores: error-compensated fp16 pairs (tc_make_geometry_split)
  bool geo_s_ok = false;
  TcGeometry geo_s{};
  __nv_bfloat16 *wpack_s = nullptr;
  size_t wpack_s_elems = 0;
  // ... and its row-pair variant for the 3x3 layers with 8 / 16 output channels (tensor-pipe-bound in split mode)
  bool geo_s2_ok = false;
  TcGeometry geo_s2{};
  __nv_bfloat16 *wpack_s2 = nullptr;
  size_t wpack_s2_elems = 0;
};

// Workspace views for one (n,h,w)
struct BlockIO {
  void *in = nullptr;  int in_planes_total = 0, in_plane0 = 0, in_planes = 0, in_h = 0, in_w = 0;
  void *out = nullptr; int out_planes_total = 0, out_plane0 = 0, out_planes = 0, out_h = 0, out_w = 0;
  void *pool = nullptr; int pool_h = 0, pool_w = 0;   // pooled copy (encoder last block)
  bool use_tc = false;
  bool pool_fused = false;   // 2x2 max-pool written by the conv epilogue
  bool head_fused = false;   // 1x1 conv + softmax computed in this block's epilogue
  void *stem_in = nullptr;   // TC stem: the uint8 image widened to the activation type ([N][H][W/8][8])
  TcPlan plan;
};

}  // namespace octseg

struct octseg_net {
  octseg_config cfg{};
  int device = 0;
  int precision = 0;
  cudaStream_t stream = nullptr;
  std::vector<octseg::BlockSpec> blocks;
  std::vector<octseg::ParamSpec> params;
  int64_t total_floats = 0;
  float *d_params = nullptr;          // flat fp32 master weights (Keras order)
  std::vector<float> h_params;        // host mirror
  bool disable_rowpair = false;       // env OCTSEG_DISABLE_ROWPAIR=1 (experiments / tests)
  bool host_stale = false;            // device params changed (training) since last mirror
  bool derived_dirty = true;          // folded BN / packed weights need a rebuild
  std::vector<octseg::BlockState> bstate;
  // workspace
  int ws_n = 0, ws_h = 0, ws_w = 0;
  void *ws = nullptr;
  size_t ws_bytes = 0;
  std::vector<octseg::BlockIO> io;
  // host-call staging: two slots, so that the H2D of one call overlaps the forward / D2H of the previous one
  // (octseg_predict_maps_submit / octseg_predict_wait); the synchronous entry points use them alternately too
  struct HostSlot {
    void *d_img = nullptr; size_t d_img_bytes = 0;
    float *d_probs = nullptr; size_t d_probs_bytes = 0;
    uint8_t *d_labels = nullptr; size_t d_labels_bytes = 0;
    uint8_t *d_maps = nullptr; size_t d_maps_bytes = 0;
    std::vector<cudaEvent_t> ev;          // per chunk: H2D done, forward done
    cudaEvent_t ev_start = nullptr, ev_done = nullptr, ev_status = nullptr;
    int *h_status = nullptr;              // pinned snapshot of the status words after this call's last forward
    bool busy = false;
    // arguments of the call in flight (re-run on the CUDA-core path if the fp16-pair range overflowed)
    const void *images = nullptr; int dtype = 0, n = 0, h = 0, w = 0;
    float *probs = nullptr; uint8_t *labels = nullptr, *maps = nullptr; int bg_ilm = 0, bg_csi = 0, transposed = 0;
    bool pipelined = false;
  } slot[2];
  int next_slot = 0;
  void *d_eval = nullptr; size_t d_eval_bytes = 0;      // octseg_evaluate_host: true labels, counts, loss sums, class weights
  int *d_status = nullptr;            // [0] pipeline time-out code, [1] fp16-pair range overflow (split mode)
  int *h_status = nullptr;            // pinned, 2 ints
  // fp32 mode: 0 = tensor cores on error-compensated fp16 pairs where the shape allows (default), 1 = CUDA cores
  // (env OCTSEG_FP32_PATH=cuda, or set for good once an activation left the fp16 range)
  cudaStream_t last_train_stream = nullptr;   // caller stream of the last octseg_train_step_device (parameters are written there)
  cudaEvent_t ev_derived = nullptr;           // folded BN / packed weights ready (recorded on `stream`)
  int fp32_path = 0;
  bool ws_split = false;              // the planned workspace uses the split layout
  int64_t launches = 0;
  bool disable_tc = false;
  bool disable_fusion = false;
  int microbatch = 0;
  cudaStream_t copy_in = nullptr, copy_out = nullptr;   // host-API pipeline: H2D / D2H beside the compute stream
  bool profiling = false;
  std::vector<cudaEvent_t> prof_events;   // blocks+1 events
  bool prof_valid = false;
  // training state lives in train.cu
  void *train = nullptr;
};
