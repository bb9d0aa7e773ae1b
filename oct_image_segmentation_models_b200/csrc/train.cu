// Training half of the C ABI: one synchronous data-parallel step
//   forward (batch-statistics BN, dropout) -> weighted CE -> backward -> NCCL all-reduce ->
//   Keras-Adam, mirroring model.fit under MirroredStrategy (reference training/training.py:185,
//   262-266, 401-407; loss math common/custom_losses.py:27-35).
#include <dlfcn.h>

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <cstdio>

#include "conv_tc.cuh"
#include "kernels.cuh"
#include "net.cuh"
#include "train_kernels.cuh"

using namespace octseg;

namespace octseg {
int ensure_workspace(octseg_net *net, int n, int h, int w);
int check_status(octseg_net *net);

// ---- NCCL, resolved at run time from the process' libnccl (torch ships one) ------------------
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
struct NcclApi {
  void *lib = nullptr;
  int (*GetUniqueId)(ncclUniqueId *) = nullptr;
  int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;

static int load_nccl() {
  if (g_nccl.lib) return 0;
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char *n : names) {
    g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.lib) break;
  }
  if (!g_nccl.lib) { set_error(std::string("cannot dlopen libnccl.so.2: ") + dlerror()); return 1; }
  g_nccl.GetUniqueId = (int (*)(ncclUniqueId *))dlsym(g_nccl.lib, "ncclGetUniqueId");
  g_nccl.CommInitRank = (int (*)(ncclComm_t *, int, ncclUniqueId, int))dlsym(g_nccl.lib, "ncclCommInitRank");
  g_nccl.AllReduce = (int (*)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t))dlsym(g_nccl.lib, "ncclAllReduce");
  g_nccl.CommDestroy = (int (*)(ncclComm_t))dlsym(g_nccl.lib, "ncclCommDestroy");
  g_nccl.GetErrorString = (const char *(*)(int))dlsym(g_nccl.lib, "ncclGetErrorString");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy) {
    set_error("libnccl is missing required symbols");
    return 1;
  }
  return 0;
}
#define OCTSEG_NCCL(expr)                                                                  \
  do {                                                                                     \
    int _r = (expr);                                                                       \
    if (_r != 0) {                                                                         \
      set_error(std::string(#expr) + ": " +                                               \
                (g_nccl.GetErrorString ? g_nccl.GetErrorString(_r) : "nccl error"));      \
      return 1;                                                                            \
    }                                                                                      \
  } while (0)

// ---- per-block training tensors ------------------------------------------------------------------
struct TrainBlock {
  void *z = nullptr;        // pre-BN conv output [n][cout/8][h][w][8]
  void *a = nullptr;        // post BN+ReLU(+dropout) activation (may alias a concat-buffer slice)
  int a_planes_total = 0, a_plane0 = 0;
  void *pooled = nullptr;   // 2x2 max-pooled activation
  void *dz = nullptr;       // gradient wrt z
  float *mean = nullptr, *invstd = nullptr, *scale = nullptr, *shift = nullptr;   // [cout]
  float *w_t = nullptr;     // transformed weights for the data gradient
  int h = 0, w = 0;         // output grid
  // tensor-core path (bf16 mode): forward z-conv and data-gradient conv through conv_tc_kernel
  bool tc_fwd = false, tc_dgrad = false;
  bool stats_fused = false;   // the forward conv's epilogue accumulates the BatchNorm batch statistics
  TcGeometry geo_dgrad{};
  bool geo_dgrad_ok = false;
  __nv_bfloat16 *wpack_dgrad = nullptr;
  // row-pair variants (conv_tc.cuh: tc_rowpair_weights) for narrow 3x3 layers
  TcGeometry geo_dgrad2{};
  bool geo_dgrad2_ok = false, fwd_pair = false, dgrad_pair = false;
  bool dgrad_s2d = false;         // up-conv with <= 32 output channels: data gradient on the low-res grid (tc_make_geometry_s2d)
  __nv_bfloat16 *wpack_dgrad2 = nullptr;
  TcPlan plan_fwd, plan_dgrad;
};

struct TrainState {
  octseg_train_config tc{};
  float *d_class_w = nullptr;
  float *d_grads = nullptr, *d_m = nullptr, *d_v = nullptr;
  float *d_ones = nullptr;          // [max cout] of 1.0f: conv epilogue scale in training
  float *d_zeros = nullptr;         // [max cout] of 0.0f: conv epilogue shift of the data gradient
  double *d_sums = nullptr;         // [blocks][2 phases][2 * max cout]: all zeroed by one memset per step
  double *d_loss = nullptr;
  double *h_loss = nullptr;         // pinned
  float *d_stem_tmp = nullptr;      // [taps][8][cout] stem wgrad scratch
  long long step = 0;
  int max_cout = 0;
  // workspace
  int n = 0, h = 0, w = 0;
  void *ws = nullptr;
  size_t ws_bytes = 0;
  std::vector<TrainBlock> tb;
  std::vector<void *> dcat;         // gradient wrt each level's concat buffer [n][2f/8][h][w][8]
  std::vector<void *> gA, gB;       // ping-pong gradient buffers per level (f channels... sized 2f)
  StepState *d_state = nullptr;     // step counter + bias-corrected learning rate, advanced on the device
  // CUDA graph of the whole step (captured after one eager step with the same arguments)
  cudaGraphExec_t graph_exec = nullptr;
  const void *g_img = nullptr, *g_lab = nullptr, *g_mask = nullptr;
  float *g_loss = nullptr;
  int g_n = 0, g_h = 0, g_w = 0, g_dtype = -1;
  cudaStream_t g_stream = nullptr;
  bool warm = false;
  long long launches_per_step = 0;
  TcPackJob *d_pack_jobs = nullptr; // all tensor-core weight images of the net, packed in one launch per step
  int n_pack_jobs = 0;
  void *img_blocked = nullptr;      // stem input as a 1-plane blocked tensor
  float *d_dlog = nullptr;          // [n*h*w][K] dlogits of heads with more than 16 input channels (two-pass head)
  size_t dlog_floats = 0;
  float *d_wg_scratch = nullptr;    // split-K partial sums of the tcgen05 weight gradient (wgrad_tc.cu)
  size_t wg_scratch_floats = 0;
  void *up_scratch = nullptr;       // materialised x2-upsampled input of an up-conv (weight-gradient stream)
  void *mask = nullptr;             // dropout multiplier tensor at the bottleneck
  void *d_img = nullptr; size_t d_img_bytes = 0;
  uint8_t *d_labels = nullptr; size_t d_labels_bytes = 0;
  uint8_t *d_mask_in = nullptr; size_t d_mask_bytes = 0;
  // weight gradients run on a second stream, concurrently with the rest of the backward chain
  cudaStream_t wg_stream = nullptr;
  std::vector<cudaEvent_t> ev_dz;   // per block: dz ready
  cudaEvent_t ev_wg_done = nullptr, ev_step_start = nullptr;
  // gradient all-reduce in two reverse-order buckets on its own stream: the decoder-side bucket (bottleneck, decoder,
  // head: the tail of the flat buffer, complete first) is reduced while the encoder's backward pass still runs
  cudaStream_t comm_stream = nullptr;
  bool tail_sent = false;           // the decoder-side bucket of the current step has been issued
  cudaEvent_t ev_tail_main = nullptr, ev_tail_wg = nullptr, ev_head_main = nullptr, ev_comm_done = nullptr;
  // communicator
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
};

// optional per-phase CUDA-event profile of one train step (env OCTSEG_TRAIN_PROFILE=1)
struct PhaseProf {
  bool on = false;
  cudaStream_t st = nullptr;
  std::vector<std::pair<int, std::pair<cudaEvent_t, cudaEvent_t>>> spans;
  int cur = -1;
  int tag = 0, cur_tag = 0;      // block index for per-layer detail
  bool detail = false;
  cudaEvent_t cur_start = nullptr;
  void begin(int phase) {
    if (!on) return;
    end();
    cur = phase;
    cur_tag = tag;
    cudaEventCreate(&cur_start);
    cudaEventRecord(cur_start, st);
  }
  void end() {
    if (!on || cur < 0) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, st);
    spans.push_back({cur + 100 * cur_tag, {cur_start, e}});
    cur = -1;
  }
  void report() {
    if (!on) return;
    end();
    cudaStreamSynchronize(st);
    static const char *names[] = {"conv_fwd", "bn_fwd", "pool_fwd", "head_loss", "pool_bwd", "bn_bwd", "wgrad", "dgrad",
                                  "allreduce", "adam", "misc"};
    double tot[11] = {0};
    for (auto &s : spans) {
      float ms = 0;
      cudaEventElapsedTime(&ms, s.second.first, s.second.second);
      tot[s.first % 100] += ms;
      if (detail && ms > 0.002f) fprintf(stderr, "[train detail] block %d %s %.3f ms\n", s.first / 100, names[s.first % 100], ms);
      cudaEventDestroy(s.second.first);
      cudaEventDestroy(s.second.second);
    }
    double sum = 0;
    for (double t : tot) sum += t;
    fprintf(stderr, "[train profile] total %.2f ms:", sum);
    for (int i = 0; i < 11; ++i) fprintf(stderr, " %s %.2f", names[i], tot[i]);
    fprintf(stderr, "\n");
    spans.clear();
  }
};
enum { PH_CONV = 0, PH_BNF, PH_POOLF, PH_HEAD, PH_POOLB, PH_BNB, PH_WGRAD, PH_DGRAD, PH_AR, PH_ADAM, PH_MISC };

// weight gradient: tensor-core kernel in bf16 mode, CUDA-core kernel in fp32 mode
template <typename T>
static int wgrad_dispatch(octseg_net *net, View<const T> a_in, View<const T> dz, int kh, int kw, int pt, int pl, int ups,
                          int cin, int cout, float *dW, float *db, cudaStream_t st) {
  // kernel launches of the chosen path (gpu_launches in bench.py): tcgen05 = kernel + split-K reduce + two bias stages,
  // row-walking kernel = 1, deep kernel = kernel + bias gradient, CUDA cores = 1
  if constexpr (sizeof(T) == 2) {
    if (!net->disable_tc) {
      TrainState *S = reinterpret_cast<TrainState *>(net->train);
      // >= 64 input channels, dense tensors: tcgen05 with MN-major operands and a deterministic split-K reduction
      static const bool tc_off = []() { const char *e = std::getenv("OCTSEG_WGRAD_TC"); return e && e[0] == '0'; }();
      if (!tc_off && S && S->d_wg_scratch && wgrad_tc_applicable(kh, kw, cin, cout, ups, dz.h, dz.w) &&
          a_in.h == dz.h && a_in.w == dz.w && a_in.img_stride == (long long)a_in.planes * a_in.h * a_in.w * 8 &&
          dz.img_stride == (long long)dz.planes * dz.h * dz.w * 8 &&
          (size_t)wgrad_tc_scratch_floats(kh, kw, cin, cout, dz.n, dz.h, dz.w) <= S->wg_scratch_floats) {
        net->launches += 4;
        return launch_wgrad_tc(a_in, dz, kh, kw, pt, pl, cin, cout, dW, db, S->d_wg_scratch, S->wg_scratch_floats,
                               net->d_status, st);
      }
      net->launches += wgrad_rows_applicable(kh, kw, cin, ups) ? 1 : 2;
      return launch_wgrad_mma(a_in, dz, kh, kw, pt, pl, ups, cin, cout, dW, db, net->d_status, st);
    }
  }
  ++net->launches;
  return launch_wgrad<T>(a_in, dz, kh, kw, pt, pl, ups, cin, cout, dW, db, st);
}

static TrainState *ts(octseg_net *net) { return reinterpret_cast<TrainState *>(net->train); }
static size_t esz(const octseg_net *net) { return net->precision == OCTSEG_BF16 ? 2 : 4; }

struct Bump2 {
  size_t off = 0;
  size_t take(size_t bytes) { size_t o = off; off += (bytes + 1023) / 1024 * 1024; return o; }
};

static int ensure_train_workspace(octseg_net *net, int n, int h, int w) {
  TrainState *S = ts(net);
  if (S->ws && S->n == n && S->h == h && S->w == w) return 0;
  const int P = net->cfg.pool_layers, s = net->cfg.start_neurons;
  if ((h % (1 << P)) || (w % (1 << P))) { set_error("image height/width must be multiples of 2^pool_layers"); return 1; }
  const size_t es = esz(net);
  auto bytes = [&](int ch, int lvl) { return (size_t)n * ch * (h >> lvl) * (w >> lvl) * es; };
  Bump2 bump;
  const size_t nb = net->blocks.size();
  std::vector<size_t> o_z(nb), o_a(nb), o_pool(nb), o_dz(nb);
  std::vector<size_t> o_cat(P), o_dcat(P), o_gA(P + 1), o_gB(P + 1);
  for (int l = 0; l < P; ++l) { o_cat[l] = bump.take(bytes(2 * (s << l), l)); o_dcat[l] = bump.take(bytes(2 * (s << l), l)); }
  for (int l = 0; l <= P; ++l) { o_gA[l] = bump.take(bytes(2 * (s << l), l)); o_gB[l] = bump.take(bytes(2 * (s << l), l)); }
  for (auto &b : net->blocks) {
    if (b.role == 4) continue;
    o_z[b.index] = bump.take(bytes(b.cout, b.level));
    o_dz[b.index] = bump.take(bytes(b.cout, b.level));
    const bool in_cat = (b.role == 0 && b.pool_after) || b.role == 2;
    o_a[b.index] = in_cat ? (size_t)-1 : bump.take(bytes(b.cout, b.level));
    o_pool[b.index] = b.pool_after ? bump.take(bytes(b.cout, b.level + 1)) : (size_t)-1;
  }
  size_t up_bytes = 1024;
  // x2-upsampled input of the up-convs whose weight gradient needs a real tensor (>= 64 input channels); the narrow ones
  // read the low-res tensor, and the data gradient stores 2x2 sums straight from its epilogue
  for (auto &b : net->blocks)
    if (b.ups && !wgrad_rows_applicable(b.kh, b.kw, b.cin, 1)) up_bytes = std::max(up_bytes, bytes(b.cin, b.level));
  const size_t o_up = bump.take(up_bytes);
  const size_t o_img = bump.take((size_t)n * h * w * 8 * es);
  const size_t o_mask = bump.take(bytes(s << P, P));
  if (bump.off > S->ws_bytes) {
    if (S->ws) OCTSEG_CUDA(cudaFree(S->ws));
    S->ws = nullptr;
    OCTSEG_CUDA(cudaMalloc(&S->ws, bump.off));
    S->ws_bytes = bump.off;
  }
  uint8_t *base = reinterpret_cast<uint8_t *>(S->ws);
  S->dcat.resize(P); S->gA.resize(P + 1); S->gB.resize(P + 1);
  for (int l = 0; l < P; ++l) S->dcat[l] = base + o_dcat[l];
  for (int l = 0; l <= P; ++l) { S->gA[l] = base + o_gA[l]; S->gB[l] = base + o_gB[l]; }
  for (auto &b : net->blocks) {
    if (b.role == 4) continue;
    TrainBlock &t = S->tb[b.index];
    t.h = h >> b.level; t.w = w >> b.level;
    t.z = base + o_z[b.index];
    t.dz = base + o_dz[b.index];
    const int f8 = b.cout / 8;
    if (b.role == 0 && b.pool_after) { t.a = base + o_cat[b.level]; t.a_planes_total = 2 * f8; t.a_plane0 = f8; }
    else if (b.role == 2) { t.a = base + o_cat[b.level]; t.a_planes_total = 2 * f8; t.a_plane0 = 0; }
    else { t.a = base + o_a[b.index]; t.a_planes_total = f8; t.a_plane0 = 0; }
    t.pooled = b.pool_after ? base + o_pool[b.index] : nullptr;
  }
  S->img_blocked = base + o_img;
  S->up_scratch = base + o_up;
  S->mask = base + o_mask;
  S->n = n; S->h = h; S->w = w;
  if (net->blocks.back().cin > 16 || (net->blocks.back().cin == 16 && net->cfg.num_classes > 8)) {
    const size_t need = (size_t)n * h * w * net->cfg.num_classes;
    if (need > S->dlog_floats) {
      if (S->d_dlog) OCTSEG_CUDA(cudaFree(S->d_dlog));
      S->d_dlog = nullptr;
      OCTSEG_CUDA(cudaMalloc(&S->d_dlog, need * sizeof(float)));
      S->dlog_floats = need;
    }
  }
  if (net->precision == OCTSEG_BF16 && !net->disable_tc) {
    size_t need = 0;
    for (auto &b : net->blocks) {
      if (b.role == 4 || b.index == 0) continue;
      // the up-conv's weight gradient runs on the materialised x2 tensor, i.e. on this block's OUTPUT grid
      need = std::max(need, wgrad_tc_scratch_floats(b.kh, b.kw, b.cin, b.cout, n, h >> b.level, w >> b.level));
    }
    if (need > S->wg_scratch_floats) {
      if (S->d_wg_scratch) OCTSEG_CUDA(cudaFree(S->d_wg_scratch));
      S->d_wg_scratch = nullptr;
      OCTSEG_CUDA(cudaMalloc(&S->d_wg_scratch, need * sizeof(float)));
      S->wg_scratch_floats = need;
    }
  }
  // ---- tensor-core plans (bf16): z = conv(in)+bias and d(in) = conv(dz, W^T flipped)
  for (auto &b : net->blocks) {
    if (b.role == 4) continue;
    TrainBlock &t = S->tb[b.index];
    t.tc_fwd = t.tc_dgrad = false;
    t.stats_fused = false;
    if (net->precision != OCTSEG_BF16 || net->disable_tc || b.index == 0) continue;
    const BlockState &bs = net->bstate[b.index];
    // input view of this block (same rules as block_input)
    const BlockSpec &pb = net->blocks[b.index - 1];
    const TrainBlock &pt = S->tb[pb.index];
    const __nv_bfloat16 *in_ptr;
    int in_h, in_w;
    bool dense_in = true;
    if (b.concat_level >= 0) { in_ptr = (const __nv_bfloat16 *)pt.a; in_h = pt.h; in_w = pt.w; }
    else if (pb.pool_after) { in_ptr = (const __nv_bfloat16 *)pt.pooled; in_h = pt.h / 2; in_w = pt.w / 2; }
    else { in_ptr = (const __nv_bfloat16 *)pt.a; in_h = pt.h; in_w = pt.w; dense_in = (pt.a_planes_total == pb.cout / 8); }
    if (bs.geo_ok && dense_in && tc_supported(b.kh, b.kw, b.cin, b.cout, b.ups, in_h, in_w)) {
      TcEpilogue epi;
      epi.relu = 0;
      epi.scale = S->d_ones;
      epi.shift = net->d_params + net->params[b.p_bias].offset;
      epi.out = make_view((__nv_bfloat16 *)t.z, n, b.cout / 8, 0, b.cout / 8, t.h, t.w);
      epi.stats = S->d_sums + ((size_t)b.index * 2 + 0) * 2 * S->max_cout;      // == sums_of(b.index, 0)
      t.stats_fused = true;
      t.fwd_pair = bs.geo2_ok && (in_h % 2) == 0 && in_h >= 2 * kTcTileH;
      if (tc_make_plan(t.fwd_pair ? bs.geo2 : bs.geo, in_ptr, n, in_h, in_w, t.fwd_pair ? bs.wpack2 : bs.wpack, epi,
                       net->d_status, &t.plan_fwd))
        return 1;
      t.tc_fwd = true;
    }
    if (t.geo_dgrad_ok && tc_supported(b.kh, b.kw, b.cout, b.cin, 0, t.dgrad_s2d ? t.h / 2 : t.h, t.dgrad_s2d ? t.w / 2 : t.w)) {
      // destination of the data gradient: fixed per block (see the buffer walk in train_step_t)
      t.tc_dgrad = true;
      t.dgrad_pair = t.geo_dgrad2_ok && !b.ups && (t.h % 2) == 0 && t.h >= 2 * kTcTileH;
    }
  }
  for (auto &b : net->blocks) S->tb[b.index].plan_dgrad.valid = false;
  std::vector<TcPackJob> jobs;
  for (auto &b : net->blocks) {
    if (b.role == 4) continue;
    TrainBlock &t = S->tb[b.index];
    if (t.tc_fwd) {
      TcPackJob j;
      const BlockState &bs = net->bstate[b.index];
      j.g = t.fwd_pair ? bs.geo2 : bs.geo; j.w = net->d_params + net->params[b.p_kernel].offset;
      j.out = t.fwd_pair ? bs.wpack2 : bs.wpack; j.total = (long long)(t.fwd_pair ? bs.wpack2_elems : bs.wpack_elems);
      j.transposed = 0;
      jobs.push_back(j);
    }
    if (t.tc_dgrad) {
      TcPackJob j;
      const TcGeometry &gd = t.dgrad_pair ? t.geo_dgrad2 : t.geo_dgrad;
      j.g = gd; j.w = net->d_params + net->params[b.p_kernel].offset; j.out = t.dgrad_pair ? t.wpack_dgrad2 : t.wpack_dgrad;
      j.total = (long long)gd.n_tiles_n * gd.cin_chunks * gd.ksteps * 2 * gd.n_cols * 8;
      j.transposed = 1;
      jobs.push_back(j);
    }
  }
  if (S->d_pack_jobs) OCTSEG_CUDA(cudaFree(S->d_pack_jobs));
  S->d_pack_jobs = nullptr;
  S->n_pack_jobs = (int)jobs.size();
  if (!jobs.empty()) {
    OCTSEG_CUDA(cudaMalloc(&S->d_pack_jobs, jobs.size() * sizeof(TcPackJob)));
    OCTSEG_CUDA(cudaMemcpy(S->d_pack_jobs, jobs.data(), jobs.size() * sizeof(TcPackJob), cudaMemcpyHostToDevice));
  }
  return 0;
}

// Input activation of block b (as a view) during training
template <typename T>
static View<const T> block_input(octseg_net *net, const BlockSpec &b, int n) {
  TrainState *S = ts(net);
  if (b.index == 0) return make_view((const T *)S->img_blocked, n, 1, 0, 1, S->h, S->w);
  const BlockSpec &pb = net->blocks[b.index - 1];
  const TrainBlock &pt = S->tb[pb.index];
  if (b.concat_level >= 0) {   // concat([up, skip]) buffer of this level == where the up block wrote
    return make_view((const T *)pt.a, n, pt.a_planes_total, 0, pt.a_planes_total, pt.h, pt.w);
  }
  if (pb.pool_after) return make_view((const T *)pt.pooled, n, pb.cout / 8, 0, pb.cout / 8, pt.h / 2, pt.w / 2);
  return make_view((const T *)pt.a, n, pt.a_planes_total, pt.a_plane0, pb.cout / 8, pt.h, pt.w);
}

// Gradient all-reduce (sum over ranks; the loss is already scaled by the GLOBAL batch), bucketed and overlapped:
// bucket 1 = parameters of the bottleneck, decoder and head blocks (the tail of the flat Keras-order buffer, whose
// gradients are complete when the backward pass enters the encoder), bucket 0 = the encoder blocks.  Both
// collectives run on `comm_stream`, forked from / joined to the step's streams with events, so they are part of
// the captured CUDA graph of the step.
static int64_t bucket_split(const octseg_net *net) {
  for (auto &b : net->blocks)
    if (b.role != 0) return b.index == 0 ? 0 : net->params[b.p_kernel].offset;
  return 0;
}
static int allreduce_tail_bucket(octseg_net *net, cudaStream_t st, cudaStream_t wst) {
  TrainState *S = ts(net);
  if (!(S->comm && S->world > 1)) return 0;
  const int64_t split = bucket_split(net);
  if (split <= 0 || split >= net->total_floats) return 0;
  OCTSEG_CUDA(cudaEventRecord(S->ev_tail_main, st));
  OCTSEG_CUDA(cudaStreamWaitEvent(S->comm_stream, S->ev_tail_main, 0));
  if (wst != st) {
    OCTSEG_CUDA(cudaEventRecord(S->ev_tail_wg, wst));
    OCTSEG_CUDA(cudaStreamWaitEvent(S->comm_stream, S->ev_tail_wg, 0));
  }
  OCTSEG_NCCL(g_nccl.AllReduce(S->d_grads + split, S->d_grads + split, (size_t)(net->total_floats - split), /*ncclFloat*/ 7,
                               /*ncclSum*/ 0, S->comm, S->comm_stream));
  return 0;
}
// `st` has already joined the weight-gradient stream
static int train_tail(octseg_net *net, cudaStream_t st) {
  TrainState *S = ts(net);
  if (S->comm && S->world > 1) {
    int64_t split = bucket_split(net);
    if (!S->tail_sent || split <= 0 || split >= net->total_floats) split = net->total_floats;     // single bucket
    OCTSEG_CUDA(cudaEventRecord(S->ev_head_main, st));
    OCTSEG_CUDA(cudaStreamWaitEvent(S->comm_stream, S->ev_head_main, 0));
    OCTSEG_NCCL(g_nccl.AllReduce(S->d_grads, S->d_grads, (size_t)split, /*ncclFloat*/ 7, /*ncclSum*/ 0, S->comm,
                                 S->comm_stream));
    OCTSEG_CUDA(cudaEventRecord(S->ev_comm_done, S->comm_stream));
    OCTSEG_CUDA(cudaStreamWaitEvent(st, S->ev_comm_done, 0));
  }
  if (launch_adam(net->d_params, S->d_grads, S->d_m, S->d_v, net->total_floats, S->d_state, S->tc.beta_1,
                  S->tc.beta_2, S->tc.epsilon, st))
    return 1;
  ++net->launches;
  return 0;
}

template <typename T>
static int train_step_t(octseg_net *net, const void *d_img, int dtype, const uint8_t *d_labels, int n, int h,
                        int w, const uint8_t *d_mask_in, cudaStream_t st, bool with_tail) {
  TrainState *S = ts(net);
  const float *P = net->d_params;
  float *G = S->d_grads;
  const int nblk = (int)net->blocks.size();
  const BlockSpec &head = net->blocks.back();
  if (launch_step_advance(S->d_state, S->tc.learning_rate, S->tc.beta_1, S->tc.beta_2, st)) return 1;
  ++net->launches;
  PhaseProf prof;
  prof.on = std::getenv("OCTSEG_TRAIN_PROFILE") != nullptr;
  prof.detail = prof.on && std::getenv("OCTSEG_TRAIN_PROFILE")[0] == '2';
  prof.st = st;
  prof.begin(PH_MISC);
  OCTSEG_CUDA(cudaMemsetAsync(G, 0, net->total_floats * sizeof(float), st));
  OCTSEG_CUDA(cudaMemsetAsync(S->d_loss, 0, sizeof(double), st));
  OCTSEG_CUDA(cudaMemsetAsync(S->d_sums, 0, net->blocks.size() * 2 * 2 * S->max_cout * sizeof(double), st));
  auto sums_of = [&](int block, int phase) { return S->d_sums + ((size_t)block * 2 + phase) * 2 * S->max_cout; };
  if (tc_pack_all_device(S->d_pack_jobs, S->n_pack_jobs, st)) return 1;
  ++net->launches;
  const bool dual = !prof.on && std::getenv("OCTSEG_NO_DUAL_STREAM") == nullptr;
  cudaStream_t wst = dual ? S->wg_stream : st;
  if (launch_image_to_blocked<T>(d_img, dtype, n, h, w, net->cfg.input_channels, (T *)S->img_blocked, st)) return 1;
  ++net->launches;
  const BlockSpec *mid_last = nullptr;
  for (auto &b : net->blocks) if (b.dropout_after) mid_last = &b;
  const bool use_dropout = mid_last && (d_mask_in || S->tc.dropout_rate > 0.f);
  if (use_dropout) {
    const TrainBlock &t = S->tb[mid_last->index];
    if (launch_dropout_mask<T>(d_mask_in, S->tc.dropout_seed, S->d_state,
                               S->tc.dropout_rate > 0.f ? S->tc.dropout_rate : 0.5f, n, mid_last->cout, t.h, t.w,
                               (T *)S->mask, st))
      return 1;
    ++net->launches;
  }
  // ------------------------------- forward -------------------------------
  for (auto &b : net->blocks) {
    if (b.role == 4) break;
    TrainBlock &t = S->tb[b.index];
    prof.tag = b.index;
    View<const T> in = block_input<T>(net, b, n);
    View<T> z = make_view((T *)t.z, n, b.cout / 8, 0, b.cout / 8, t.h, t.w);
    // z = conv(in) + bias  (epilogue scale = 1, shift = bias, no ReLU)
    prof.begin(PH_CONV);
    if (b.index == 0) {
      if (launch_conv_first<T>(d_img, dtype, n, h, w, b.cin, P + net->params[b.p_kernel].offset, b.kh, b.kw, b.cout,
                               S->d_ones, P + net->params[b.p_bias].offset, 0, z, st))
        return 1;
      ++net->launches;
    } else if (t.tc_fwd) {
      if (tc_launch(t.plan_fwd, st)) return 1;
      ++net->launches;
    } else {
      if (launch_conv_direct<T>(in, P + net->params[b.p_kernel].offset, b.kh, b.kw, b.cin, b.cout, b.ups ? 1 : 0, S->d_ones,
                                P + net->params[b.p_bias].offset, 0, z, st))
        return 1;
      ++net->launches;
    }
    View<const T> zc = make_view((const T *)t.z, n, b.cout / 8, 0, b.cout / 8, t.h, t.w);
    prof.begin(PH_BNF);
    if (!t.stats_fused) {      // the tensor-core conv's epilogue has already accumulated sum(z), sum(z^2)
      if (launch_bn_stats<T>(zc, sums_of(b.index, 0), st)) return 1;
      ++net->launches;
    }
    View<T> a = make_view((T *)t.a, n, t.a_planes_total, t.a_plane0, b.cout / 8, t.h, t.w);
    const T *mask = (use_dropout && b.dropout_after) ? (const T *)S->mask : nullptr;
    View<T> po{};
    po.ptr = nullptr;
    if (b.pool_after) po = make_view((T *)t.pooled, n, b.cout / 8, 0, b.cout / 8, t.h / 2, t.w / 2);
    // batch mean / variance, a = relu(bn(z)) (* dropout), the pooled copy and the moving-statistics update: one pass
    if (launch_bn_finalize_apply<T>(zc, sums_of(b.index, 0), (long long)n * t.h * t.w, 1e-3f, 0.99f,
                                    P + net->params[b.p_gamma].offset, P + net->params[b.p_beta].offset,
                                    net->d_params + net->params[b.p_mean].offset, net->d_params + net->params[b.p_var].offset,
                                    t.mean, t.invstd, t.scale, t.shift, mask, a, po, st))
      return 1;
    ++net->launches;
  }
  // ------------------------------- loss + head backward -------------------------------
  const BlockSpec &last = net->blocks[nblk - 2];
  const TrainBlock &tl = S->tb[last.index];
  prof.begin(PH_HEAD);
  {
    View<const T> a = make_view((const T *)tl.a, n, tl.a_planes_total, tl.a_plane0, last.cout / 8, tl.h, tl.w);
    View<T> da = make_view((T *)S->gA[0], n, last.cout / 8, 0, last.cout / 8, tl.h, tl.w);
    const float inv_den = 1.0f / ((float)S->tc.global_batch * (float)h * (float)w);
    if (launch_head_loss<T>(a, P + net->params[head.p_kernel].offset, P + net->params[head.p_bias].offset, head.cin,
                            head.cout, d_labels, S->d_class_w, inv_den, da, G + net->params[head.p_kernel].offset,
                            G + net->params[head.p_bias].offset, S->d_loss, S->d_dlog, st))
      return 1;
    ++net->launches;
  }
  // ------------------------------- backward -------------------------------
  // `g` = gradient wrt the output activation of the block being processed
  const void *g_ptr = S->gA[0];
  int g_planes_total = last.cout / 8, g_plane0 = 0;
  S->tail_sent = false;
  for (int bi = nblk - 2; bi >= 0; --bi) {
    const BlockSpec &b = net->blocks[bi];
    TrainBlock &t = S->tb[bi];
    prof.tag = bi;
    if (with_tail && !S->tail_sent && b.role == 0 && S->comm && S->world > 1) {     // every bottleneck / decoder / head gradient has been issued
      if (allreduce_tail_bucket(net, st, wst)) return 1;
      S->tail_sent = true;
    }
    const int f8 = b.cout / 8;
    View<const T> zc = make_view((const T *)t.z, n, f8, 0, f8, t.h, t.w);
    View<const T> da = make_view((const T *)g_ptr, n, g_planes_total, g_plane0, f8, t.h, t.w);
    if (b.pool_after) {
      prof.begin(PH_POOLB);
      // da = d(skip half of the concat gradient) + scatter(d pooled); g currently holds d pooled
      View<const T> ac = make_view((const T *)t.a, n, t.a_planes_total, t.a_plane0, f8, t.h, t.w);
      View<const T> dpool = make_view((const T *)g_ptr, n, g_planes_total, g_plane0, f8, t.h / 2, t.w / 2);
      View<const T> dskip = make_view((const T *)S->dcat[b.level], n, 2 * f8, f8, f8, t.h, t.w);
      View<T> tot = make_view((T *)S->gB[b.level], n, f8, 0, f8, t.h, t.w);
      // ... and the BatchNorm-backward reductions of this block while the totals are in registers
      if (launch_pool_bwd_add_bnred<T>(ac, dpool, dskip, tot, zc, t.mean, t.invstd, P + net->params[b.p_gamma].offset,
                                       P + net->params[b.p_beta].offset, sums_of(bi, 1), st))
        return 1;
      ++net->launches;
      da = make_view((const T *)S->gB[b.level], n, f8, 0, f8, t.h, t.w);
    }
    const T *mask = (use_dropout && b.dropout_after) ? (const T *)S->mask : nullptr;
    const float *gamma = P + net->params[b.p_gamma].offset, *beta = P + net->params[b.p_beta].offset;
    prof.begin(PH_BNB);
    if (!b.pool_after) {
      if (launch_bn_bwd_reduce<T>(da, zc, t.mean, t.invstd, gamma, beta, mask, sums_of(bi, 1), st)) return 1;
      ++net->launches;
    }
    View<T> dz = make_view((T *)t.dz, n, f8, 0, f8, t.h, t.w);
    if (launch_bn_bwd_apply<T>(da, zc, t.mean, t.invstd, gamma, beta, mask, sums_of(bi, 1), (long long)n * t.h * t.w, dz,
                               G + net->params[b.p_gamma].offset, G + net->params[b.p_beta].offset, st))
      return 1;
    ++net->launches;
    View<const T> dzc = make_view((const T *)t.dz, n, f8, 0, f8, t.h, t.w);
    View<const T> in = block_input<T>(net, b, n);
    const int pt = (b.kh - 1) / 2, pl = (b.kw - 1) / 2;
    prof.begin(PH_WGRAD);
    if (dual) {   // dW needs only dz and the saved input: run it beside the remaining backward chain
      OCTSEG_CUDA(cudaEventRecord(S->ev_dz[bi], st));
      OCTSEG_CUDA(cudaStreamWaitEvent(wst, S->ev_dz[bi], 0));
    }
    if (b.index == 0) {
      const int taps = b.kh * b.kw;
      OCTSEG_CUDA(cudaMemsetAsync(S->d_stem_tmp, 0, (size_t)taps * 8 * b.cout * sizeof(float), wst));
      if (wgrad_dispatch<T>(net, in, dzc, b.kh, b.kw, pt, pl, 0, 8, b.cout, S->d_stem_tmp,
                            G + net->params[b.p_bias].offset, wst))
        return 1;
      if (launch_stem_wgrad_extract(S->d_stem_tmp, taps, b.cin, b.cout, G + net->params[b.p_kernel].offset, wst)) return 1;
      ++net->launches;
      break;
    }
    if (b.ups && sizeof(T) == 2 && !net->disable_tc && !wgrad_rows_applicable(b.kh, b.kw, b.cin, 1)) {
      // tensor-core path wants a real tensor: materialise the x2-upsampled input once (1 write + 1 read
      // of a tensor the forward pass never stores) and run the plain 2x2 weight gradient on it
      View<T> up = make_view((T *)S->up_scratch, n, b.cin / 8, 0, b.cin / 8, t.h, t.w);
      if (launch_upsample2x<T>(in, up, wst)) return 1;
      ++net->launches;
      View<const T> upc = make_view((const T *)S->up_scratch, n, b.cin / 8, 0, b.cin / 8, t.h, t.w);
      if (wgrad_dispatch<T>(net, upc, dzc, b.kh, b.kw, pt, pl, 0, b.cin, b.cout, G + net->params[b.p_kernel].offset,
                            G + net->params[b.p_bias].offset, wst))
        return 1;
    } else if (wgrad_dispatch<T>(net, in, dzc, b.kh, b.kw, pt, pl, b.ups ? 1 : 0, b.cin, b.cout,
                                 G + net->params[b.p_kernel].offset, G + net->params[b.p_bias].offset, wst))
      return 1;
    // ---- data gradient wrt this block's input
    prof.begin(PH_DGRAD);
    const BlockSpec &pb = net->blocks[bi - 1];
    if (!b.ups) {
      if (!t.tc_dgrad && launch_flip_transpose(P + net->params[b.p_kernel].offset, b.kh, b.kw, b.cin, b.cout, t.w_t, st))
        return 1;
      void *dst;
      int dst_total;
      if (b.concat_level >= 0) { dst = S->dcat[b.level]; dst_total = b.cin / 8; }
      else { dst = (g_ptr == S->gA[b.level]) ? S->gB[b.level] : S->gA[b.level]; dst_total = b.cin / 8; }
      // destination grid = this block's input grid (pooled input has the same h,w as the output here)
      View<T> din = make_view((T *)dst, n, dst_total, 0, b.cin / 8, t.h, t.w);
      if (t.tc_dgrad) {
        TcEpilogue epi;
        epi.relu = 0; epi.scale = S->d_ones; epi.shift = S->d_zeros;
        epi.out = make_view((__nv_bfloat16 *)dst, n, dst_total, 0, b.cin / 8, t.h, t.w);
        // the destination buffer of a block never changes between steps: build the plan once
        if (!t.plan_dgrad.valid || t.plan_dgrad.p.out != epi.out.ptr) {
          if (tc_make_plan(t.dgrad_pair ? t.geo_dgrad2 : t.geo_dgrad, (const __nv_bfloat16 *)t.dz, n, t.h, t.w,
                           t.dgrad_pair ? t.wpack_dgrad2 : t.wpack_dgrad, epi, net->d_status, &t.plan_dgrad))
            return 1;
        }
        if (tc_launch(t.plan_dgrad, st)) return 1;
      } else if (launch_conv_direct_ex<T>(dzc, t.w_t, b.kh, b.kw, b.cout, b.cin, 0, 1, b.kh - 1 - pt, b.kw - 1 - pl,
                                          S->d_ones, S->d_zeros, 0, din, st))
        return 1;
      g_ptr = dst;
      if (b.concat_level >= 0) { g_planes_total = b.cin / 8; g_plane0 = 0; }   // next: the up block (planes [0,f/8))
      else { g_planes_total = b.cin / 8; g_plane0 = 0; }
    } else {
      void *dst = S->gA[b.level + 1];
      View<T> din = make_view((T *)dst, n, b.cin / 8, 0, b.cin / 8, t.h / 2, t.w / 2);
      if (t.tc_dgrad && t.dgrad_s2d) {
        // stride-2 3x3 conv over the four pixel parities of dz, straight on the low-res grid
        TcEpilogue epi;
        epi.relu = 0; epi.scale = S->d_ones; epi.shift = S->d_zeros;
        epi.out = make_view((__nv_bfloat16 *)dst, n, b.cin / 8, 0, b.cin / 8, t.h / 2, t.w / 2);
        if (!t.plan_dgrad.valid || t.plan_dgrad.p.out != epi.out.ptr) {
          if (tc_make_plan(t.geo_dgrad, (const __nv_bfloat16 *)t.dz, n, t.h / 2, t.w / 2, t.wpack_dgrad, epi, net->d_status,
                           &t.plan_dgrad))
            return 1;
        }
        if (tc_launch(t.plan_dgrad, st)) return 1;
      } else if (t.tc_dgrad) {
        // d(upsampled input) on the high-res grid with the flipped kernel; the epilogue stores only its 2x2 sums
        // (the adjoint of the nearest up-sampling), so the 4x larger tensor is neither written nor re-read
        TcEpilogue epi;
        epi.relu = 0; epi.scale = S->d_ones; epi.shift = S->d_zeros;
        epi.out = make_view((__nv_bfloat16 *)nullptr, n, b.cin / 8, 0, b.cin / 8, t.h, t.w);
        epi.pool_out = (__nv_bfloat16 *)dst;
        epi.pool_img_stride = din.img_stride;
        epi.pool_sum = 1;
        if (!t.plan_dgrad.valid || t.plan_dgrad.p.pool_out != epi.pool_out) {
          if (tc_make_plan(t.geo_dgrad, (const __nv_bfloat16 *)t.dz, n, t.h, t.w, t.wpack_dgrad, epi, net->d_status,
                           &t.plan_dgrad))
            return 1;
        }
        if (tc_launch(t.plan_dgrad, st)) return 1;
      } else {
        // up-conv: d(prev) on the low-res grid = stride-2 (kh+1)x(kw+1) conv over dz
        if (launch_upconv_dgrad_weights(P + net->params[b.p_kernel].offset, b.kh, b.kw, b.cin, b.cout, t.w_t, st)) return 1;
        if (launch_conv_direct_ex<T>(dzc, t.w_t, b.kh + 1, b.kw + 1, b.cout, b.cin, 0, 2, b.kh - 1 - pt, b.kw - 1 - pl,
                                     S->d_ones, S->d_zeros, 0, din, st))
          return 1;
      }
      g_ptr = dst;
      g_planes_total = b.cin / 8; g_plane0 = 0;
    }
    (void)pb;
    net->launches += t.tc_dgrad ? 1 : 2;        // data gradient: one tensor-core conv, or weight transform + direct conv
  }
  // ------------------------------- all-reduce + optimizer -------------------------------
  if (dual) {
    OCTSEG_CUDA(cudaEventRecord(S->ev_wg_done, wst));
    OCTSEG_CUDA(cudaStreamWaitEvent(st, S->ev_wg_done, 0));
  }
  if (with_tail) {
    prof.begin(PH_AR);
    if (train_tail(net, st)) return 1;
  }
  prof.report();
  net->host_stale = true;
  net->derived_dirty = true;
  return 0;
}


// ---- per-block debug entries (tests): the backward kernels of ONE conv block on caller data -------------------------
// a_in: NHWC fp32 [n,h,w,cin] (the block's input activation), dz: NHWC fp32 [n,oh,ow,cout] (gradient wrt the conv output;
// oh = 2h for the up-conv).  Runs exactly the launchers the train step uses in the handle's precision (bf16: tensor-core
// weight gradient + tcgen05 data gradient; fp32: CUDA cores) and returns dW [kh,kw,cin,cout], db [cout], d_in NHWC [n,h,w,cin].
template <typename T>
static int debug_backward_t(octseg_net *net, const BlockSpec &b, const float *a_in, const float *dz, int n, int h, int w,
                            float *dW_out, float *db_out, float *din_out) {
  TrainState *S = ts(net);
  TrainBlock &t = S->tb[b.index];
  cudaStream_t st = net->stream;
  const int oh = b.ups ? 2 * h : h, ow = b.ups ? 2 * w : w;
  const size_t in_elems = (size_t)n * b.cin * h * w, dz_elems = (size_t)n * b.cout * oh * ow;
  auto to_blocked = [&](const float *src, int c, int hh, int ww, std::vector<T> *dst) {
    dst->resize((size_t)n * c * hh * ww);
    for (int i = 0; i < n; ++i)
      for (int y = 0; y < hh; ++y)
        for (int x = 0; x < ww; ++x)
          for (int k = 0; k < c; ++k) {
            const float v = src[(((size_t)i * hh + y) * ww + x) * c + k];
            T o;
            if constexpr (sizeof(T) == 2) o = __float2bfloat16(v); else o = v;
            (*dst)[((((size_t)i * (c / 8) + k / 8) * hh + y) * ww + x) * 8 + (k & 7)] = o;
          }
  };
  std::vector<T> hin, hdz;
  to_blocked(a_in, b.cin, h, w, &hin);
  to_blocked(dz, b.cout, oh, ow, &hdz);
  T *d_in = nullptr, *d_dz = nullptr, *d_din = nullptr, *d_up = nullptr;
  float *d_dW = nullptr, *d_db = nullptr;
  const size_t w_count = (size_t)b.kh * b.kw * b.cin * b.cout;
  OCTSEG_CUDA(cudaMalloc(&d_in, in_elems * sizeof(T)));
  OCTSEG_CUDA(cudaMalloc(&d_dz, dz_elems * sizeof(T)));
  OCTSEG_CUDA(cudaMalloc(&d_din, in_elems * sizeof(T)));
  OCTSEG_CUDA(cudaMalloc(&d_dW, w_count * sizeof(float)));
  OCTSEG_CUDA(cudaMalloc(&d_db, b.cout * sizeof(float)));
  OCTSEG_CUDA(cudaMemcpyAsync(d_in, hin.data(), in_elems * sizeof(T), cudaMemcpyHostToDevice, st));
  OCTSEG_CUDA(cudaMemcpyAsync(d_dz, hdz.data(), dz_elems * sizeof(T), cudaMemcpyHostToDevice, st));
  OCTSEG_CUDA(cudaMemsetAsync(d_din, 0xFF, in_elems * sizeof(T), st));
  OCTSEG_CUDA(cudaMemsetAsync(d_dW, 0, w_count * sizeof(float), st));
  OCTSEG_CUDA(cudaMemsetAsync(d_db, 0, b.cout * sizeof(float), st));
  View<const T> in = make_view((const T *)d_in, n, b.cin / 8, 0, b.cin / 8, h, w);
  View<const T> dzc = make_view((const T *)d_dz, n, b.cout / 8, 0, b.cout / 8, oh, ow);
  const int pt = (b.kh - 1) / 2, pl = (b.kw - 1) / 2;
  const float *P = net->d_params;
  int rc = 0;
  if (sizeof(T) == 2 && !net->disable_tc) {     // scratch of the tcgen05 weight gradient for THIS shape
    const size_t need = wgrad_tc_scratch_floats(b.kh, b.kw, b.cin, b.cout, n, oh, ow);
    if (need > S->wg_scratch_floats) {
      OCTSEG_CUDA(cudaStreamSynchronize(st));
      if (S->d_wg_scratch) OCTSEG_CUDA(cudaFree(S->d_wg_scratch));
      S->d_wg_scratch = nullptr;
      OCTSEG_CUDA(cudaMalloc(&S->d_wg_scratch, need * sizeof(float)));
      S->wg_scratch_floats = need;
    }
  }
  // ---- weight gradient (same dispatch as train_step_t)
  if (b.ups && sizeof(T) == 2 && !net->disable_tc && !wgrad_rows_applicable(b.kh, b.kw, b.cin, 1)) {
    OCTSEG_CUDA(cudaMalloc(&d_up, (size_t)n * b.cin * oh * ow * sizeof(T)));
    View<T> up = make_view(d_up, n, b.cin / 8, 0, b.cin / 8, oh, ow);
    rc = launch_upsample2x<T>(in, up, st);
    View<const T> upc = make_view((const T *)d_up, n, b.cin / 8, 0, b.cin / 8, oh, ow);
    if (!rc) rc = wgrad_dispatch<T>(net, upc, dzc, b.kh, b.kw, pt, pl, 0, b.cin, b.cout, d_dW, d_db, st);
  } else {
    rc = wgrad_dispatch<T>(net, in, dzc, b.kh, b.kw, pt, pl, b.ups ? 1 : 0, b.cin, b.cout, d_dW, d_db, st);
  }
  // ---- data gradient
  const bool tc_dgrad = sizeof(T) == 2 && !net->disable_tc && t.geo_dgrad_ok &&
                        tc_supported(b.kh, b.kw, b.cout, b.cin, 0, t.dgrad_s2d ? h : oh, t.dgrad_s2d ? w : ow);
  if (!rc && tc_dgrad) {
    const bool pair = t.geo_dgrad2_ok && !b.ups && (oh % 2) == 0 && oh >= 2 * kTcTileH;
    const TcGeometry &gd = pair ? t.geo_dgrad2 : t.geo_dgrad;
    __nv_bfloat16 *wp = pair ? t.wpack_dgrad2 : t.wpack_dgrad;
    rc = tc_pack_weights_device(gd, P + net->params[b.p_kernel].offset, 1, wp, st);
    TcEpilogue epi;
    epi.relu = 0; epi.scale = S->d_ones; epi.shift = S->d_zeros;
    TcPlan plan;
    if (!b.ups) {
      epi.out = make_view((__nv_bfloat16 *)d_din, n, b.cin / 8, 0, b.cin / 8, h, w);
      if (!rc) rc = tc_make_plan(gd, (const __nv_bfloat16 *)d_dz, n, oh, ow, wp, epi, net->d_status, &plan);
      if (!rc) rc = tc_launch(plan, st);
    } else if (t.dgrad_s2d) {
      epi.out = make_view((__nv_bfloat16 *)d_din, n, b.cin / 8, 0, b.cin / 8, h, w);
      if (!rc) rc = tc_make_plan(gd, (const __nv_bfloat16 *)d_dz, n, h, w, wp, epi, net->d_status, &plan);
      if (!rc) rc = tc_launch(plan, st);
    } else {
      epi.out = make_view((__nv_bfloat16 *)nullptr, n, b.cin / 8, 0, b.cin / 8, oh, ow);
      epi.pool_out = (__nv_bfloat16 *)d_din;
      epi.pool_img_stride = (long long)b.cin * h * w;
      epi.pool_sum = 1;
      if (!rc) rc = tc_make_plan(gd, (const __nv_bfloat16 *)d_dz, n, oh, ow, wp, epi, net->d_status, &plan);
      if (!rc) rc = tc_launch(plan, st);
    }
    cudaStreamSynchronize(st);
    tc_release_plan(&plan);
  } else if (!rc) {
    View<T> din = make_view(d_din, n, b.cin / 8, 0, b.cin / 8, h, w);
    if (!b.ups) {
      rc = launch_flip_transpose(P + net->params[b.p_kernel].offset, b.kh, b.kw, b.cin, b.cout, t.w_t, st);
      if (!rc) rc = launch_conv_direct_ex<T>(dzc, t.w_t, b.kh, b.kw, b.cout, b.cin, 0, 1, b.kh - 1 - pt, b.kw - 1 - pl, S->d_ones,
                                             S->d_zeros, 0, din, st);
    } else {
      rc = launch_upconv_dgrad_weights(P + net->params[b.p_kernel].offset, b.kh, b.kw, b.cin, b.cout, t.w_t, st);
      if (!rc) rc = launch_conv_direct_ex<T>(dzc, t.w_t, b.kh + 1, b.kw + 1, b.cout, b.cin, 0, 2, b.kh - 1 - pt, b.kw - 1 - pl,
                                             S->d_ones, S->d_zeros, 0, din, st);
    }
  }
  if (!rc) rc = check_status(net);
  if (!rc) {
    std::vector<T> hd(in_elems);
    cudaMemcpy(hd.data(), d_din, in_elems * sizeof(T), cudaMemcpyDeviceToHost);
    cudaMemcpy(dW_out, d_dW, w_count * sizeof(float), cudaMemcpyDeviceToHost);
    cudaMemcpy(db_out, d_db, b.cout * sizeof(float), cudaMemcpyDeviceToHost);
    for (int i = 0; i < n; ++i)
      for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x)
          for (int k = 0; k < b.cin; ++k) {
            const T v = hd[((((size_t)i * (b.cin / 8) + k / 8) * h + y) * w + x) * 8 + (k & 7)];
            float f;
            if constexpr (sizeof(T) == 2) f = __bfloat162float(v); else f = v;
            din_out[(((size_t)i * h + y) * w + x) * b.cin + k] = f;
          }
  }
  cudaFree(d_in); cudaFree(d_dz); cudaFree(d_din); cudaFree(d_dW); cudaFree(d_db); cudaFree(d_up);
  return rc;
}

}  // namespace octseg

extern "C" int32_t octseg_debug_backward_block(octseg_net *net, int32_t conv_index, const float *a_in, const float *dz, int32_t n,
                                               int32_t h, int32_t w, float *dW, float *db, float *d_in) {
  if (!net || !a_in || !dz || !dW || !db || !d_in) { set_error("null argument"); return 1; }
  TrainState *S = ts(net);
  if (!S) { set_error("call octseg_train_begin first"); return 1; }
  if (conv_index <= 0 || conv_index >= (int)net->blocks.size() - 1) { set_error("bad conv index (1 .. last conv block)"); return 1; }
  if (net->precision == OCTSEG_FP16) { set_error("training kernels run in fp32 or bf16"); return 1; }
  OCTSEG_CUDA(cudaSetDevice(net->device));
  const BlockSpec &b = net->blocks[conv_index];
  return net->precision == OCTSEG_BF16 ? debug_backward_t<__nv_bfloat16>(net, b, a_in, dz, n, h, w, dW, db, d_in)
                                       : debug_backward_t<float>(net, b, a_in, dz, n, h, w, dW, db, d_in);
}

namespace octseg {
__global__ void store_loss_kernel(const double *src, float *dst) { *dst = (float)*src; }

}  // namespace octseg

extern "C" {

void octseg_train_free(octseg_net *net) {
  TrainState *S = ts(net);
  if (!S) return;
  // the captured step holds NCCL's persistent plans: release the graph before the communicator (ncclCommDestroy
  // otherwise waits for them forever)
  if (S->graph_exec) { cudaGraphExecDestroy(S->graph_exec); S->graph_exec = nullptr; }
  cudaDeviceSynchronize();
  if (S->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(S->comm);
  for (auto &e : S->ev_dz) cudaEventDestroy(e);
  if (S->ev_wg_done) cudaEventDestroy(S->ev_wg_done);
  if (S->ev_step_start) cudaEventDestroy(S->ev_step_start);
  if (S->wg_stream) cudaStreamDestroy(S->wg_stream);
  if (S->comm_stream) cudaStreamDestroy(S->comm_stream);
  for (cudaEvent_t e : {S->ev_tail_main, S->ev_tail_wg, S->ev_head_main, S->ev_comm_done}) if (e) cudaEventDestroy(e);
  for (auto &t : S->tb) { cudaFree(t.mean); cudaFree(t.invstd); cudaFree(t.scale); cudaFree(t.shift); cudaFree(t.w_t); cudaFree(t.wpack_dgrad); cudaFree(t.wpack_dgrad2); tc_release_plan(&t.plan_dgrad); }
  cudaFree(S->d_class_w); cudaFree(S->d_grads); cudaFree(S->d_m); cudaFree(S->d_v); cudaFree(S->d_ones); cudaFree(S->d_zeros);
  cudaFree(S->d_sums); cudaFree(S->d_loss); cudaFree(S->d_stem_tmp); cudaFree(S->ws); cudaFree(S->d_img);
  cudaFree(S->d_labels); cudaFree(S->d_mask_in); cudaFree(S->d_pack_jobs); cudaFree(S->d_state); cudaFree(S->d_wg_scratch); cudaFree(S->d_dlog);
  if (S->graph_exec) cudaGraphExecDestroy(S->graph_exec);
  if (S->h_loss) cudaFreeHost(S->h_loss);
  delete S;
  net->train = nullptr;
}

int32_t octseg_train_begin(octseg_net *net, const octseg_train_config *tc, const float *class_weights) {
  if (!net || !tc || !class_weights) { set_error("null argument"); return 1; }
  if (tc->global_batch <= 0) { set_error("global_batch must be positive"); return 1; }
  if (tc->dropout_rate < 0.f || tc->dropout_rate >= 1.f) { set_error("dropout_rate must be in [0,1)"); return 1; }
  if (net->precision == OCTSEG_FP16) { set_error("training runs in fp32 or bf16 mode (fp16 storage is inference-only)"); return 1; }
  OCTSEG_CUDA(cudaSetDevice(net->device));
  TrainState *S = ts(net);
  if (!S) {
    S = new TrainState();
    net->train = S;
    const size_t fb = net->total_floats * sizeof(float);
    OCTSEG_CUDA(cudaMalloc(&S->d_grads, fb));
    OCTSEG_CUDA(cudaMalloc(&S->d_m, fb));
    OCTSEG_CUDA(cudaMalloc(&S->d_v, fb));
    OCTSEG_CUDA(cudaMalloc(&S->d_class_w, net->cfg.num_classes * sizeof(float)));
    int maxc = 8;
    for (auto &b : net->blocks) maxc = std::max(maxc, std::max(b.cin, b.cout));
    S->max_cout = maxc;
    std::vector<float> ones(maxc, 1.0f);
    OCTSEG_CUDA(cudaMalloc(&S->d_ones, maxc * sizeof(float)));
    OCTSEG_CUDA(cudaMemcpy(S->d_ones, ones.data(), maxc * sizeof(float), cudaMemcpyHostToDevice));
    OCTSEG_CUDA(cudaMalloc(&S->d_zeros, maxc * sizeof(float)));
    OCTSEG_CUDA(cudaMemset(S->d_zeros, 0, maxc * sizeof(float)));
    OCTSEG_CUDA(cudaMalloc(&S->d_sums, net->blocks.size() * 2 * 2 * maxc * sizeof(double)));
    OCTSEG_CUDA(cudaMalloc(&S->d_loss, sizeof(double)));
    OCTSEG_CUDA(cudaMalloc(&S->d_state, sizeof(StepState)));
    OCTSEG_CUDA(cudaMallocHost(&S->h_loss, sizeof(double)));
    const BlockSpec &b0 = net->blocks[0];
    OCTSEG_CUDA(cudaMalloc(&S->d_stem_tmp, (size_t)b0.kh * b0.kw * 8 * b0.cout * sizeof(float)));
    S->tb.resize(net->blocks.size());
    OCTSEG_CUDA(cudaStreamCreateWithFlags(&S->wg_stream, cudaStreamNonBlocking));
    S->ev_dz.resize(net->blocks.size());
    for (auto &e : S->ev_dz) OCTSEG_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    OCTSEG_CUDA(cudaEventCreateWithFlags(&S->ev_wg_done, cudaEventDisableTiming));
    OCTSEG_CUDA(cudaEventCreateWithFlags(&S->ev_step_start, cudaEventDisableTiming));
    OCTSEG_CUDA(cudaStreamCreateWithFlags(&S->comm_stream, cudaStreamNonBlocking));
    for (cudaEvent_t *e : {&S->ev_tail_main, &S->ev_tail_wg, &S->ev_head_main, &S->ev_comm_done})
      OCTSEG_CUDA(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    for (auto &b : net->blocks) {
      if (b.role == 4) continue;
      TrainBlock &t = S->tb[b.index];
      OCTSEG_CUDA(cudaMalloc(&t.mean, b.cout * sizeof(float)));
      OCTSEG_CUDA(cudaMalloc(&t.invstd, b.cout * sizeof(float)));
      OCTSEG_CUDA(cudaMalloc(&t.scale, b.cout * sizeof(float)));
      OCTSEG_CUDA(cudaMalloc(&t.shift, b.cout * sizeof(float)));
      OCTSEG_CUDA(cudaMalloc(&t.w_t, (size_t)(b.kh + 1) * (b.kw + 1) * b.cin * b.cout * sizeof(float)));
      static const bool s2d_off = []() { const char *e = std::getenv("OCTSEG_DGRAD_S2D"); return e && e[0] == '0'; }();
      t.dgrad_s2d = net->precision == OCTSEG_BF16 && !net->disable_tc && !s2d_off && b.ups && b.kh == 2 && b.kw == 2 &&
                    tc_make_geometry_s2d(b.cout, b.cin, &t.geo_dgrad) == 0;
      if (net->precision == OCTSEG_BF16 && !net->disable_tc && b.index > 0 &&
          tc_supported(b.kh, b.kw, b.cout, b.cin, 0, kTcTileH, kTcTileW) &&
          (t.dgrad_s2d || tc_make_geometry(b.kh, b.kw, b.cout, b.cin, 0, &t.geo_dgrad, b.kh - 1 - (b.kh - 1) / 2,
                                           b.kw - 1 - (b.kw - 1) / 2) == 0)) {
        t.geo_dgrad_ok = true;
        const size_t elems = (size_t)t.geo_dgrad.n_tiles_n * t.geo_dgrad.cin_chunks * t.geo_dgrad.ksteps * 2 *
                             t.geo_dgrad.n_cols * 8;
        OCTSEG_CUDA(cudaMalloc(&t.wpack_dgrad, elems * 2));
        // data gradient of a 3x3 layer whose INPUT has 8 / 16 channels: row-pair variant
        if (!net->disable_rowpair && !b.ups && b.kh == 3 && b.kw == 3 && (b.cin == 8 || b.cin == 16) &&
            tc_make_geometry(4, 3, b.cout, 2 * b.cin, 0, &t.geo_dgrad2, 1, 1) == 0 && t.geo_dgrad2.n_tiles_n == 1) {
          t.geo_dgrad2_ok = true;
          t.geo_dgrad2.rows2 = 1;
          const size_t e2 = (size_t)t.geo_dgrad2.n_tiles_n * t.geo_dgrad2.cin_chunks * t.geo_dgrad2.ksteps * 2 *
                            t.geo_dgrad2.n_cols * 8;
          OCTSEG_CUDA(cudaMalloc(&t.wpack_dgrad2, e2 * 2));
        }
      }
    }
  }
  S->tc = *tc;
  S->step = 0;
  OCTSEG_CUDA(cudaMemset(S->d_state, 0, sizeof(StepState)));
  if (S->graph_exec) { cudaGraphExecDestroy(S->graph_exec); S->graph_exec = nullptr; }
  S->warm = false;
  OCTSEG_CUDA(cudaMemset(S->d_m, 0, net->total_floats * sizeof(float)));
  OCTSEG_CUDA(cudaMemset(S->d_v, 0, net->total_floats * sizeof(float)));
  OCTSEG_CUDA(cudaMemset(S->d_grads, 0, net->total_floats * sizeof(float)));
  OCTSEG_CUDA(cudaMemcpy(S->d_class_w, class_weights, net->cfg.num_classes * sizeof(float), cudaMemcpyHostToDevice));
  return 0;
}

int32_t octseg_comm_unique_id(uint8_t *id_out) {
  if (!id_out) { set_error("null argument"); return 1; }
  if (load_nccl()) return 1;
  ncclUniqueId id;
  OCTSEG_NCCL(g_nccl.GetUniqueId(&id));
  std::memcpy(id_out, &id, 128);
  return 0;
}

int32_t octseg_comm_init(octseg_net *net, const uint8_t *unique_id, int32_t rank, int32_t world) {
  if (!net || !unique_id) { set_error("null argument"); return 1; }
  TrainState *S = ts(net);
  if (!S) { set_error("call octseg_train_begin first"); return 1; }
  if (world < 1 || rank < 0 || rank >= world) { set_error("bad rank/world"); return 1; }
  OCTSEG_CUDA(cudaSetDevice(net->device));
  S->rank = rank; S->world = world;
  if (S->graph_exec) { cudaGraphExecDestroy(S->graph_exec); S->graph_exec = nullptr; }   // captured without the all-reduce
  S->warm = false;
  if (world == 1) return 0;
  if (load_nccl()) return 1;
  ncclUniqueId id;
  std::memcpy(&id, unique_id, 128);
  OCTSEG_NCCL(g_nccl.CommInitRank(&S->comm, world, id, rank));
  return 0;
}

int32_t octseg_train_step_device(octseg_net *net, const void *images, int32_t dtype, const uint8_t *labels,
                                 int32_t n, int32_t h, int32_t w, const uint8_t *dropout_mask,
                                 float *loss_out_device, void *stream) {
  if (!net || !images || !labels) { set_error("null argument"); return 1; }
  TrainState *S = ts(net);
  if (!S) { set_error("call octseg_train_begin first"); return 1; }
  if (n <= 0 || h <= 0 || w <= 0) { set_error("bad batch shape"); return 1; }
  OCTSEG_CUDA(cudaSetDevice(net->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : net->stream;
  net->last_train_stream = st;
  if (ensure_train_workspace(net, n, h, w)) return 1;
  auto run_eager = [&](bool with_tail) -> int {
    int rc = net->precision == OCTSEG_BF16
                 ? train_step_t<__nv_bfloat16>(net, images, dtype, labels, n, h, w, dropout_mask, st, with_tail)
                 : train_step_t<float>(net, images, dtype, labels, n, h, w, dropout_mask, st, with_tail);
    if (rc) return 1;
    // loss is accumulated in double on the device; pinned host copy for the host API,
    // float copy into the caller's device word for the device API
    OCTSEG_CUDA(cudaMemcpyAsync(S->h_loss, S->d_loss, sizeof(double), cudaMemcpyDeviceToHost, st));
    if (loss_out_device) {
      store_loss_kernel<<<1, 1, 0, st>>>(S->d_loss, loss_out_device);
      OCTSEG_CUDA(cudaGetLastError());
    }
    return 0;
  };
  // The whole step is stream-ordered and every per-step scalar lives on the device, so after one eager
  // step with the same arguments the step is captured into a CUDA graph and replayed.  The bucketed NCCL
  // all-reduces (on their own stream) and the Adam launch are captured with it; OCTSEG_TRAIN_GRAPH_NCCL=0
  // keeps them outside the graph (eager), OCTSEG_TRAIN_GRAPH=0 disables graphs altogether.
  const char *genv = std::getenv("OCTSEG_TRAIN_GRAPH");
  const bool graphs_on = !(genv && genv[0] == '0');
  const bool same = S->warm && S->g_img == images && S->g_lab == labels && S->g_mask == dropout_mask && S->g_n == n &&
                    S->g_h == h && S->g_w == w && S->g_dtype == dtype && S->g_stream == st && S->g_loss == loss_out_device;
  const bool profiling = std::getenv("OCTSEG_TRAIN_PROFILE") != nullptr;
  const char *gn = std::getenv("OCTSEG_TRAIN_GRAPH_NCCL");
  const bool tail_in_graph = !(S->comm && S->world > 1) || !(gn && gn[0] == '0');
  if (graphs_on && !profiling && same) {
    if (!S->graph_exec) {
      cudaGraph_t graph = nullptr;
      const long long l0 = net->launches;
      OCTSEG_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed));
      const int rc = run_eager(tail_in_graph);
      cudaError_t ce = cudaStreamEndCapture(st, &graph);
      if (rc || ce != cudaSuccess || !graph) {
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        if (rc) return 1;
        set_error(std::string("train-step graph capture failed: ") + cudaGetErrorString(ce));
        return 1;
      }
      S->launches_per_step = net->launches - l0;
      net->launches = l0;
      ce = cudaGraphInstantiate(&S->graph_exec, graph, 0);
      cudaGraphDestroy(graph);
      if (ce != cudaSuccess) { S->graph_exec = nullptr; set_error(std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ce)); return 1; }
    }
    OCTSEG_CUDA(cudaGraphLaunch(S->graph_exec, st));
    net->launches += S->launches_per_step;
    if (!tail_in_graph) { S->tail_sent = false; if (train_tail(net, st)) return 1; }
    ++S->step;
    net->host_stale = true;
    net->derived_dirty = true;
    return 0;
  }
  if (!same && S->graph_exec) { cudaGraphExecDestroy(S->graph_exec); S->graph_exec = nullptr; }
  if (run_eager(true)) return 1;
  ++S->step;
  S->warm = true;
  S->g_img = images; S->g_lab = labels; S->g_mask = dropout_mask; S->g_n = n; S->g_h = h; S->g_w = w; S->g_dtype = dtype;
  S->g_stream = st; S->g_loss = loss_out_device;
  return 0;
}

int32_t octseg_train_step_host(octseg_net *net, const void *images, int32_t dtype, const uint8_t *labels,
                               int32_t n, int32_t h, int32_t w, const uint8_t *dropout_mask, float *loss_out) {
  if (!net || !images || !labels) { set_error("null argument"); return 1; }
  TrainState *S = ts(net);
  if (!S) { set_error("call octseg_train_begin first"); return 1; }
  OCTSEG_CUDA(cudaSetDevice(net->device));
  const size_t img_bytes = (dtype == OCTSEG_U8 ? 1 : 4) * (size_t)n * h * w * net->cfg.input_channels;
  const size_t lab_bytes = (size_t)n * h * w;
  auto grow = [&](void **p, size_t *cap, size_t need) -> int {
    if (need <= *cap) return 0;
    if (*p) cudaFree(*p);
    *p = nullptr;
    if (cudaMalloc(p, need) != cudaSuccess) { set_error("cudaMalloc failed"); return 1; }
    *cap = need;
    return 0;
  };
  if (grow(&S->d_img, &S->d_img_bytes, img_bytes)) return 1;
  if (grow((void **)&S->d_labels, &S->d_labels_bytes, lab_bytes)) return 1;
  OCTSEG_CUDA(cudaMemcpyAsync(S->d_img, images, img_bytes, cudaMemcpyHostToDevice, net->stream));
  OCTSEG_CUDA(cudaMemcpyAsync(S->d_labels, labels, lab_bytes, cudaMemcpyHostToDevice, net->stream));
  const uint8_t *d_mask = nullptr;
  if (dropout_mask) {
    const int P = net->cfg.pool_layers;
    const size_t mb = (size_t)n * (h >> P) * (w >> P) * (net->cfg.start_neurons << P);
    if (grow((void **)&S->d_mask_in, &S->d_mask_bytes, mb)) return 1;
    OCTSEG_CUDA(cudaMemcpyAsync(S->d_mask_in, dropout_mask, mb, cudaMemcpyHostToDevice, net->stream));
    d_mask = S->d_mask_in;
  }
  if (octseg_train_step_device(net, S->d_img, dtype, S->d_labels, n, h, w, d_mask, nullptr, net->stream)) return 1;
  OCTSEG_CUDA(cudaStreamSynchronize(net->stream));
  if (loss_out) *loss_out = (float)*S->h_loss;
  return check_status(net);
}

// Optimizer state for checkpoints (the reference's ModelCheckpoint saves it with the model, training.py:319-326):
// which = 0: Adam first moment m, 1: second moment v; Keras weight order, same shapes as the parameters.
int32_t octseg_opt_state(octseg_net *net, int32_t set, int32_t which, int32_t index, float *host, int64_t count) {
  if (!net || !host) { set_error("null argument"); return 1; }
  TrainState *S = ts(net);
  if (!S) { set_error("call octseg_train_begin first"); return 1; }
  if (index < 0 || index >= (int)net->params.size()) { set_error("param index out of range"); return 1; }
  if (which != 0 && which != 1) { set_error("which must be 0 (m) or 1 (v)"); return 1; }
  const ParamSpec &p = net->params[index];
  if (count != p.count) { set_error("param " + p.name + ": element count mismatch"); return 1; }
  OCTSEG_CUDA(cudaSetDevice(net->device));
  OCTSEG_CUDA(cudaStreamSynchronize(net->stream));
  if (S->g_stream) OCTSEG_CUDA(cudaStreamSynchronize(S->g_stream));
  float *dev = (which == 0 ? S->d_m : S->d_v) + p.offset;
  if (set) OCTSEG_CUDA(cudaMemcpy(dev, host, count * sizeof(float), cudaMemcpyHostToDevice));
  else OCTSEG_CUDA(cudaMemcpy(host, dev, count * sizeof(float), cudaMemcpyDeviceToHost));
  return 0;
}

// optimizer iteration count (Keras `optimizer.iterations`): drives Adam's bias correction
int32_t octseg_opt_iterations(octseg_net *net, int32_t set, int64_t *iterations) {
  if (!net || !iterations) { set_error("null argument"); return 1; }
  TrainState *S = ts(net);
  if (!S) { set_error("call octseg_train_begin first"); return 1; }
  OCTSEG_CUDA(cudaSetDevice(net->device));
  OCTSEG_CUDA(cudaStreamSynchronize(net->stream));
  if (S->g_stream) OCTSEG_CUDA(cudaStreamSynchronize(S->g_stream));
  StepState h;
  OCTSEG_CUDA(cudaMemcpy(&h, S->d_state, sizeof(h), cudaMemcpyDeviceToHost));
  if (set) {
    if (*iterations < 0) { set_error("iterations must be >= 0"); return 1; }
    h.step = (unsigned long long)*iterations;
    OCTSEG_CUDA(cudaMemcpy(S->d_state, &h, sizeof(h), cudaMemcpyHostToDevice));
    S->step = *iterations;
  } else {
    *iterations = (int64_t)h.step;
  }
  return 0;
}

int32_t octseg_get_grad(octseg_net *net, int32_t index, float *host, int64_t count) {
  if (!net || !host) { set_error("null argument"); return 1; }
  TrainState *S = ts(net);
  if (!S) { set_error("call octseg_train_begin first"); return 1; }
  if (index < 0 || index >= (int)net->params.size()) { set_error("param index out of range"); return 1; }
  const ParamSpec &p = net->params[index];
  if (count != p.count) { set_error("param " + p.name + ": element count mismatch"); return 1; }
  OCTSEG_CUDA(cudaSetDevice(net->device));
  OCTSEG_CUDA(cudaStreamSynchronize(net->stream));
  OCTSEG_CUDA(cudaMemcpy(host, S->d_grads + p.offset, count * sizeof(float), cudaMemcpyDeviceToHost));
  return 0;
}

}  // extern "C"
