// C ABI (inference half) + network object.  See include/octseg.h for the contract.
#include "net.cuh"

#include <algorithm>
#include <type_traits>
#include <cstdlib>
#include <cstring>
#include <cstdio>

#include "kernels.cuh"

namespace octseg {

static thread_local std::string g_err;
void set_error(const std::string &msg) { g_err = msg; }

static const float kBnEps = 1e-3f;   // Keras BatchNormalization() default

// ---------------------------------------------------------------------------------
// graph (structure only) -- reference models/unet.py:106-153
// ---------------------------------------------------------------------------------
int build_graph(const octseg_config &c, std::vector<BlockSpec> *blocks, std::vector<ParamSpec> *params,
                int64_t *total_floats) {
  if (c.input_channels < 1 || c.num_classes < 1 || c.start_neurons < 1 || c.pool_layers < 0 ||
      c.conv_layers < 1 || c.enc_kh < 1 || c.enc_kw < 1 || c.dec_kh < 1 || c.dec_kw < 1) {
    set_error("invalid octseg_config");
    return 1;
  }
  blocks->clear();
  int idx = 0, cin = c.input_channels;
  const int P = c.pool_layers, L = c.conv_layers, s = c.start_neurons;
  auto add = [&](int role, int level, int kh, int kw, int ci, int co, int j) -> BlockSpec & {
    BlockSpec b;
    b.index = idx++; b.role = role; b.level = level; b.kh = kh; b.kw = kw; b.cin = ci; b.cout = co;
    b.conv_j = j;
    blocks->push_back(b);
    return blocks->back();
  };
  for (int i = 0; i < P; ++i) {
    const int f = s << i;
    for (int j = 0; j < L; ++j) {
      BlockSpec &b = add(0, i, c.enc_kh, c.enc_kw, cin, f, j);
      b.pool_after = (j == L - 1);
      cin = f;
    }
  }
  {
    const int f = s << P;
    for (int j = 0; j < L; ++j) {
      BlockSpec &b = add(1, P, c.enc_kh, c.enc_kw, cin, f, j);
      b.dropout_after = (j == L - 1);
      cin = f;
    }
  }
  for (int i = 0; i < P; ++i) {
    const int lvl = P - 1 - i, f = s << lvl;
    BlockSpec &u = add(2, lvl, c.dec_kh, c.dec_kw, cin, f, 0);
    u.ups = true;
    cin = 2 * f;
    for (int j = 0; j < L; ++j) {
      BlockSpec &b = add(3, lvl, c.enc_kh, c.enc_kw, cin, f, j);
      if (j == 0) b.concat_level = lvl;
      cin = f;
    }
  }
  {
    BlockSpec &h = add(4, 0, 1, 1, cin, c.num_classes, 0);
    h.has_bn = false;
  }
  params->clear();
  int64_t off = 0;
  auto addp = [&](const std::string &name, int block, bool trainable, int nd, int64_t a, int64_t b2,
                  int64_t c2, int64_t d) {
    ParamSpec p;
    p.name = name; p.ndim = nd; p.block = block; p.trainable = trainable;
    p.shape[0] = a; p.shape[1] = b2; p.shape[2] = c2; p.shape[3] = d;
    p.count = 1;
    for (int i = 0; i < nd; ++i) p.count *= p.shape[i];
    p.offset = off;
    off += (p.count + 15) / 16 * 16;
    params->push_back(p);
    return (int)params->size() - 1;
  };
  for (auto &b : *blocks) {
    const std::string sfx = b.index == 0 ? "" : "_" + std::to_string(b.index);
    b.p_kernel = addp("conv2d" + sfx + "/kernel:0", b.index, true, 4, b.kh, b.kw, b.cin, b.cout);
    b.p_bias = addp("conv2d" + sfx + "/bias:0", b.index, true, 1, b.cout, 0, 0, 0);
    if (b.has_bn) {
      const std::string bn = "batch_normalization" + sfx;
      b.p_gamma = addp(bn + "/gamma:0", b.index, true, 1, b.cout, 0, 0, 0);
      b.p_beta = addp(bn + "/beta:0", b.index, true, 1, b.cout, 0, 0, 0);
      b.p_mean = addp(bn + "/moving_mean:0", b.index, false, 1, b.cout, 0, 0, 0);
      b.p_var = addp(bn + "/moving_variance:0", b.index, false, 1, b.cout, 0, 0, 0);
    }
  }
  *total_floats = off;
  return 0;
}

static size_t elem_size(const octseg_net *net) { return net->precision == OCTSEG_FP32 ? 4 : 2; }

static int check_supported(const octseg_net *net) {
  for (auto &b : net->blocks) {
    if (b.role != 4 && b.cout % 8) { set_error("start_neurons must be a multiple of 8"); return 1; }
    if (b.index > 0 && b.cin % 8) { set_error("channel counts must be multiples of 8"); return 1; }
  }
  if (net->cfg.num_classes > 16) { set_error("num_classes > 16 not supported"); return 1; }
  return 0;
}

// ---------------------------------------------------------------------------------
// derived state: folded BN + packed tensor-core weights
// ---------------------------------------------------------------------------------
int sync_host_mirror(octseg_net *net) {
  if (!net->host_stale) return 0;
  // training writes the parameters on the caller's stream: finish it before reading them on ours
  if (net->last_train_stream && net->last_train_stream != net->stream) OCTSEG_CUDA(cudaStreamSynchronize(net->last_train_stream));
  OCTSEG_CUDA(cudaMemcpyAsync(net->h_params.data(), net->d_params, net->total_floats * sizeof(float),
                              cudaMemcpyDeviceToHost, net->stream));
  OCTSEG_CUDA(cudaStreamSynchronize(net->stream));
  net->host_stale = false;
  return 0;
}

int prepare_derived(octseg_net *net) {
  if (!net->derived_dirty) return 0;
  if (sync_host_mirror(net)) return 1;
  OCTSEG_CUDA(cudaGetLastError());
  for (auto &b : net->blocks) {
    BlockState &st = net->bstate[b.index];
    if (!b.has_bn) continue;
    const float *P = net->d_params;
    if (launch_bn_fold(P + net->params[b.p_bias].offset, P + net->params[b.p_gamma].offset,
                       P + net->params[b.p_beta].offset, P + net->params[b.p_mean].offset,
                       P + net->params[b.p_var].offset, kBnEps, b.cout, st.scale, st.shift, net->stream))
      return 1;
    ++net->launches;
    if (net->precision == OCTSEG_FP32 && st.geo_s_ok) {
      // fp32 mode on tensor cores: pre-scale by a power of two, split every (tap-folded) weight into an fp16
      // pair and lay the pairs out as [hi rows: W_hi | W_lo'] / [lo' rows: 0 | W_hi] (tc_pack_weights)
      const float *wk = net->h_params.data() + net->params[b.p_kernel].offset;
      const float ws = tc_split_weight_scale(wk, (size_t)net->params[b.p_kernel].count);
      if (ws != st.geo_s.wscale) { st.geo_s.wscale = ws; net->ws_n = 0; }   // plans carry 1/wscale: re-plan
      std::vector<uint16_t> packed;
      tc_pack_weights(st.geo_s, wk, &packed, 1);
      if (packed.size() != st.wpack_s_elems) { set_error("internal: wpack_s size"); return 1; }
      OCTSEG_CUDA(cudaMemcpyAsync(st.wpack_s, packed.data(), packed.size() * 2, cudaMemcpyHostToDevice, net->stream));
      OCTSEG_CUDA(cudaStreamSynchronize(net->stream));
      if (st.geo_s2_ok) {
        std::vector<float> pair_w;
        tc_rowpair_weights(wk, b.cin, b.cout, &pair_w);
        st.geo_s2.wscale = ws;                       // same maximum: the banded filter only re-arranges the taps
        tc_pack_weights(st.geo_s2, pair_w.data(), &packed, 1);
        if (packed.size() != st.wpack_s2_elems) { set_error("internal: wpack_s2 size"); return 1; }
        OCTSEG_CUDA(cudaMemcpyAsync(st.wpack_s2, packed.data(), packed.size() * 2, cudaMemcpyHostToDevice, net->stream));
        OCTSEG_CUDA(cudaStreamSynchronize(net->stream));
      }
    }
    if (net->precision != OCTSEG_FP32 && st.geo_ok) {
      std::vector<uint16_t> packed;
      const float *wk = net->h_params.data() + net->params[b.p_kernel].offset;
      std::vector<float> grouped;
      if (b.index == 0) {
        tc_stem_group_weights(wk, b.cout, &grouped);
        wk = grouped.data();
        if (launch_stem_rep(st.scale, st.shift, b.cout, st.rep_scale, st.rep_shift, net->stream)) return 1;
        ++net->launches;
      }
      tc_pack_weights(st.geo, wk, &packed,
                      net->precision == OCTSEG_FP16);
      if (packed.size() != st.wpack_elems) { set_error("internal: wpack size"); return 1; }
      OCTSEG_CUDA(cudaMemcpyAsync(st.wpack, packed.data(), packed.size() * 2, cudaMemcpyHostToDevice,
                                  net->stream));
      OCTSEG_CUDA(cudaStreamSynchronize(net->stream));   // `packed` is a stack-lifetime buffer
      if (st.geo2_ok) {
        std::vector<float> pair_w;
        tc_rowpair_weights(net->h_params.data() + net->params[b.p_kernel].offset, b.cin, b.cout, &pair_w);
        tc_pack_weights(st.geo2, pair_w.data(), &packed, net->precision == OCTSEG_FP16);
        if (packed.size() != st.wpack2_elems) { set_error("internal: wpack2 size"); return 1; }
        OCTSEG_CUDA(cudaMemcpyAsync(st.wpack2, packed.data(), packed.size() * 2, cudaMemcpyHostToDevice, net->stream));
        OCTSEG_CUDA(cudaStreamSynchronize(net->stream));
      }
    }
  }
  net->derived_dirty = false;
  return 0;
}

// ---------------------------------------------------------------------------------
// workspace planning
// ---------------------------------------------------------------------------------
struct Bump {
  size_t off = 0;
  size_t take(size_t bytes) { size_t o = off; off += (bytes + 1023) / 1024 * 1024; return o; }
};

// fp32 mode: can this shape run on the tensor cores (every conv block after the stem has a split plan, the head fuses)?
static bool split_applicable(const octseg_net *net, int h, int w) {
  if (net->precision != OCTSEG_FP32 || net->fp32_path != 0 || net->disable_tc || net->disable_fusion) return false;
  const BlockSpec &last = net->blocks.back();
  if (last.role != 4 || !tc_head_fusable(last.cout) || net->blocks.size() < 3) return false;
  for (auto &b : net->blocks) {
    if (b.index == 0 || b.role == 4) continue;
    const BlockState &st = net->bstate[b.index];
    if (!st.geo_s_ok) return false;
    int lh = h >> b.level, lw = w >> b.level;
    if (b.ups) { lh >>= 1; lw >>= 1; }
    if (!tc_supported(b.kh, b.kw, 2 * b.cin, 2 * b.cout, b.ups, lh, lw)) return false;
    if (b.index == last.index - 1 && st.geo_s.n_tiles_n != 1) return false;
  }
  return true;
}

int ensure_workspace(octseg_net *net, int n, int h, int w) {
  const bool split = split_applicable(net, h, w);
  if (net->ws && net->ws_n == n && net->ws_h == h && net->ws_w == w && net->ws_split == split) return 0;
  const int P = net->cfg.pool_layers, L = net->cfg.conv_layers, s = net->cfg.start_neurons;
  if ((h % (1 << P)) || (w % (1 << P))) {
    set_error("image height/width must be multiples of 2^pool_layers");
    return 1;
  }
  // split layout: fp16 elements, two physical planes (hi, lo') per logical 8-channel plane
  const size_t es = split ? 2 : elem_size(net);
  const int pm = split ? 2 : 1;
  auto bytes = [&](int ch, int lvl) { return (size_t)n * ch * pm * (h >> lvl) * (w >> lvl) * es; };
  Bump bump;
  std::vector<size_t> cat(P), pooled(P), encT0(P), encT1(P), decT0(P), decT1(P);
  for (int l = 0; l < P; ++l) {
    const int f = s << l;
    cat[l] = bump.take(bytes(2 * f, l));
    pooled[l] = bump.take(bytes(f, l + 1));
    encT0[l] = bump.take(bytes(f, l));
    encT1[l] = bump.take(bytes(f, l));
    decT0[l] = bump.take(bytes(f, l));
    decT1[l] = bump.take(bytes(f, l));
  }
  const int fm = s << P;
  size_t midT0 = bump.take(bytes(fm, P)), midT1 = bump.take(bytes(fm, P));
  const bool stem_tc = net->precision != OCTSEG_FP32 && !net->disable_tc && net->bstate[0].geo_ok && (w % 8) == 0 &&
                       tc_supported(3, 3, 8, 8 * s, 0, h, w / 8);
  size_t stem_in = stem_tc ? bump.take((size_t)n * h * w * es) : 0;
  if (bump.off > net->ws_bytes) {
    if (net->ws) OCTSEG_CUDA(cudaFree(net->ws));
    net->ws = nullptr;
    OCTSEG_CUDA(cudaMalloc(&net->ws, bump.off));
    net->ws_bytes = bump.off;
  }
  uint8_t *base = reinterpret_cast<uint8_t *>(net->ws);
  net->io.assign(net->blocks.size(), BlockIO());
  void *prev = nullptr;
  int prev_planes = 0, prev_h = h, prev_w = w;
  for (auto &b : net->blocks) {
    BlockIO &io = net->io[b.index];
    const int lh = h >> b.level, lw = w >> b.level;
    const int f = b.cout;
    const int fp = f / 8 * pm;           // physical planes of an f-channel tensor
    // ---- input
    if (b.index == 0) {
      io.in = nullptr; io.in_h = h; io.in_w = w;
    } else if (b.concat_level >= 0) {
      io.in = base + cat[b.level]; io.in_planes_total = io.in_planes = b.cin / 8 * pm; io.in_plane0 = 0;
      io.in_h = lh; io.in_w = lw;
    } else {
      io.in = prev; io.in_planes_total = io.in_planes = prev_planes; io.in_plane0 = 0;
      io.in_h = prev_h; io.in_w = prev_w;
    }
    // ---- output
    io.out_h = lh; io.out_w = lw;
    if (b.role == 4) {
      io.out = nullptr;
    } else if (b.role == 0) {
      if (b.pool_after) {
        io.out = base + cat[b.level]; io.out_planes_total = 2 * fp; io.out_plane0 = fp; io.out_planes = fp;
        io.pool = base + pooled[b.level]; io.pool_h = lh / 2; io.pool_w = lw / 2;
      } else {
        io.out = base + ((b.conv_j & 1) ? encT1[b.level] : encT0[b.level]);
        io.out_planes_total = io.out_planes = fp; io.out_plane0 = 0;
      }
    } else if (b.role == 1) {
      io.out = base + ((b.conv_j & 1) ? midT1 : midT0);
      io.out_planes_total = io.out_planes = fp; io.out_plane0 = 0;
    } else if (b.role == 2) {
      io.out = base + cat[b.level]; io.out_planes_total = 2 * fp; io.out_plane0 = 0; io.out_planes = fp;
    } else {
      io.out = base + ((b.conv_j & 1) ? decT1[b.level] : decT0[b.level]);
      io.out_planes_total = io.out_planes = fp; io.out_plane0 = 0;
    }
    // what the next block sees
    if (io.pool) { prev = io.pool; prev_planes = fp; prev_h = io.pool_h; prev_w = io.pool_w; }
    else { prev = io.out; prev_planes = fp; prev_h = lh; prev_w = lw; }
    // ---- tensor-core plan
    io.use_tc = false;
    if (split && b.index > 0 && b.role != 4) {
      const BlockState &bst = net->bstate[b.index];
      TcEpilogue epi;
      epi.fp16 = 1;
      epi.static_weights = 1;
      epi.overflow = net->d_status + 1;
      epi.scale = bst.scale; epi.shift = bst.shift;
      epi.out = make_view(reinterpret_cast<__nv_bfloat16 *>(io.out), n, io.out_planes_total, io.out_plane0,
                          io.out_planes, io.out_h, io.out_w);
      if (io.pool) {
        epi.pool_out = reinterpret_cast<__nv_bfloat16 *>(io.pool);
        epi.pool_img_stride = (long long)io.out_planes * io.pool_h * io.pool_w * 8;
        io.pool_fused = true;
      }
      const BlockSpec &last = net->blocks.back();
      if (b.index == last.index - 1) {
        epi.head_w = net->d_params + net->params[last.p_kernel].offset;
        epi.head_b = net->d_params + net->params[last.p_bias].offset;
        epi.head_k = last.cout;
        io.head_fused = true;
      }
      // row pairs (12 instead of 18 K steps per two output rows) where the tensor-core operand path is the bound;
      // the fused head is epilogue-bound and keeps single rows
      const bool pair = bst.geo_s2_ok && (io.in_h % 2) == 0 && io.in_h >= 2 * kTcTileH && !epi.head_w;
      if (tc_make_plan(pair ? bst.geo_s2 : bst.geo_s, reinterpret_cast<const __nv_bfloat16 *>(io.in), n, io.in_h, io.in_w,
                       pair ? bst.wpack_s2 : bst.wpack_s, epi, net->d_status, &io.plan))
        return 1;
      io.use_tc = true;
      continue;
    }
    if (b.index == 0 && stem_tc) {
      TcEpilogue epi;
      epi.fp16 = net->precision == OCTSEG_FP16;
      epi.static_weights = 1;
      epi.scale = net->bstate[0].rep_scale; epi.shift = net->bstate[0].rep_shift;
      epi.out = make_view(reinterpret_cast<__nv_bfloat16 *>(io.out), n, io.out_planes_total, io.out_plane0,
                          io.out_planes, io.out_h, io.out_w);
      epi.out.w = w / 8;                      // the epilogue indexes GEMM rows = pixel groups
      io.stem_in = base + stem_in;
      if (tc_make_plan(net->bstate[0].geo, reinterpret_cast<const __nv_bfloat16 *>(io.stem_in), n, h, w / 8,
                       net->bstate[0].wpack, epi, net->d_status, &io.plan))
        return 1;
      io.use_tc = true;
    }
    if (net->precision != OCTSEG_FP32 && !net->disable_tc && b.index > 0 && b.role != 4 &&
        net->bstate[b.index].geo_ok && tc_supported(b.kh, b.kw, b.cin, b.cout, b.ups, io.in_h, io.in_w)) {
      TcEpilogue epi;
      epi.fp16 = net->precision == OCTSEG_FP16;
      epi.static_weights = 1;
      epi.scale = net->bstate[b.index].scale; epi.shift = net->bstate[b.index].shift;
      epi.out = make_view(reinterpret_cast<__nv_bfloat16 *>(io.out), n, io.out_planes_total, io.out_plane0,
                          io.out_planes, io.out_h, io.out_w);
      if (io.pool && !net->disable_fusion) {
        epi.pool_out = reinterpret_cast<__nv_bfloat16 *>(io.pool);
        epi.pool_img_stride = (long long)io.out_planes * io.pool_h * io.pool_w * 8;
        io.pool_fused = true;
      }
      const BlockSpec &last = net->blocks.back();
      if (!net->disable_fusion && b.index == last.index - 1 && last.role == 4 &&
          net->bstate[b.index].geo.n_tiles_n == 1 && tc_head_fusable(last.cout)) {
        epi.head_w = net->d_params + net->params[last.p_kernel].offset;
        epi.head_b = net->d_params + net->params[last.p_bias].offset;
        epi.head_k = last.cout;
        io.head_fused = true;
      }
      const BlockState &bst = net->bstate[b.index];
      const bool pair = bst.geo2_ok && (io.in_h % 2) == 0 && io.in_h >= 2 * kTcTileH && !epi.head_w;   // the fused head is epilogue-bound: fewer MMAs do not help it
      if (tc_make_plan(pair ? bst.geo2 : bst.geo, reinterpret_cast<const __nv_bfloat16 *>(io.in), n, io.in_h,
                       io.in_w, pair ? bst.wpack2 : bst.wpack, epi, net->d_status, &io.plan))
        return 1;
      io.use_tc = true;
    }
  }
  net->ws_n = n; net->ws_h = h; net->ws_w = w; net->ws_split = split;
  return 0;
}

// ---------------------------------------------------------------------------------
// forward (inference)
// ---------------------------------------------------------------------------------
template <typename T>
static int forward_t(octseg_net *net, const void *d_img, int dtype, int n, int h, int w, float *d_probs,
                     uint8_t *d_labels, cudaStream_t st) {
  constexpr bool kSplit = std::is_same<T, SplitHalf>::value;   // fp32 mode on tensor cores: every block after the stem has a plan
  const float *P = net->d_params;
  const bool prof = net->profiling && net->prof_events.size() == net->blocks.size() + 1;
  if (prof) cudaEventRecord(net->prof_events[0], st);
  for (auto &b : net->blocks) {
    BlockIO &io = net->io[b.index];
    BlockState &bs = net->bstate[b.index];
    if (b.index > 0 && prof) cudaEventRecord(net->prof_events[b.index], st);
    if (b.role == 4) {
      if (net->io[b.index - 1].head_fused) {   // already produced by the last conv's epilogue
        if (prof) { cudaEventRecord(net->prof_events[net->blocks.size()], st); net->prof_valid = true; }
        continue;
      }
      if constexpr (kSplit) { set_error("internal: split plan without a fused head"); return 1; }
      else {
      View<const T> in = make_view(reinterpret_cast<const T *>(io.in), n, io.in_planes_total, io.in_plane0,
                                   io.in_planes, io.in_h, io.in_w);
      if (launch_head<T>(in, P + net->params[b.p_kernel].offset, P + net->params[b.p_bias].offset, b.cin,
                         b.cout, d_probs, d_labels, st))
        return 1;
      }
      ++net->launches;
      if (prof) { cudaEventRecord(net->prof_events[net->blocks.size()], st); net->prof_valid = true; }
      continue;
    }
    View<T> out = make_view(reinterpret_cast<T *>(io.out), n, io.out_planes_total, io.out_plane0,
                            io.out_planes, io.out_h, io.out_w);
    if (!kSplit && b.index == 0 && io.use_tc && dtype == OCTSEG_U8 && ((uintptr_t)d_img % 16) == 0) {
      // tensor-core stem: widen the image to 16 bits (exact), then one conv_tc launch over pixel groups
      if constexpr (!kSplit) {
      if (launch_u8_to_act<T>(reinterpret_cast<const uint8_t *>(d_img), (long long)n * h * w,
                              reinterpret_cast<T *>(io.stem_in), st))
        return 1;
      }
      ++net->launches;
      if (tc_launch(io.plan, st)) return 1;
    } else if (b.index == 0) {
      if (launch_conv_first<T>(d_img, dtype, n, h, w, b.cin, P + net->params[b.p_kernel].offset, b.kh, b.kw,
                               b.cout, bs.scale, bs.shift, 1, out, st))
        return 1;
    } else if (io.use_tc) {
      if (io.head_fused) {
        TcPlan pl = io.plan;
        pl.p.probs = d_probs; pl.p.labels = d_labels;
        if (tc_launch(pl, st)) return 1;
      } else if (tc_launch(io.plan, st)) return 1;
    } else if constexpr (kSplit) {
      set_error("internal: split plan without a tensor-core block");
      return 1;
    } else {
      View<const T> in = make_view(reinterpret_cast<const T *>(io.in), n, io.in_planes_total, io.in_plane0,
                                   io.in_planes, io.in_h, io.in_w);
      if (launch_conv_direct<T>(in, P + net->params[b.p_kernel].offset, b.kh, b.kw, b.cin, b.cout,
                                b.ups ? 1 : 0, bs.scale, bs.shift, 1, out, st))
        return 1;
    }
    ++net->launches;
    {
    if (io.pool && !io.pool_fused) {      // split mode: only the stem of a conv_layers == 1 net gets here (FFMA stem, no fused pool)
      View<const T> pin = make_view(reinterpret_cast<const T *>(io.out), n, io.out_planes_total,
                                    io.out_plane0, io.out_planes, io.out_h, io.out_w);
      View<T> pout = make_view(reinterpret_cast<T *>(io.pool), n, io.out_planes, 0, io.out_planes, io.pool_h,
                               io.pool_w);
      if (launch_maxpool2<T>(pin, pout, st)) return 1;
      ++net->launches;
    }
    }
  }
  return 0;
}

int forward(octseg_net *net, const void *d_img, int dtype, int n, int h, int w, float *d_probs,
            uint8_t *d_labels, cudaStream_t st) {
  const bool rebuilt = net->derived_dirty;
  if (prepare_derived(net)) return 1;
  if (rebuilt && st != net->stream) {
    // folded BN tables / packed weights were (re)built on the handle's stream: a caller stream must wait for them
    OCTSEG_CUDA(cudaEventRecord(net->ev_derived, net->stream));
    OCTSEG_CUDA(cudaStreamWaitEvent(st, net->ev_derived, 0));
  }
  if (ensure_workspace(net, n, h, w)) return 1;
  if (net->precision == OCTSEG_BF16)
    return forward_t<__nv_bfloat16>(net, d_img, dtype, n, h, w, d_probs, d_labels, st);
  if (net->precision == OCTSEG_FP16)
    return forward_t<__half>(net, d_img, dtype, n, h, w, d_probs, d_labels, st);
  if (net->ws_split) return forward_t<SplitHalf>(net, d_img, dtype, n, h, w, d_probs, d_labels, st);
  return forward_t<float>(net, d_img, dtype, n, h, w, d_probs, d_labels, st);
}

// 0 = ok, 1 = error, 2 = an activation left the fp16-pair range of the fp32 tensor-core path: the results of the
// work since the last check are invalid and the handle has switched to the CUDA-core fp32 path for good
int check_status(octseg_net *net) {
  OCTSEG_CUDA(cudaMemcpyAsync(net->h_status, net->d_status, 2 * sizeof(int), cudaMemcpyDeviceToHost, net->stream));
  OCTSEG_CUDA(cudaStreamSynchronize(net->stream));
  if (net->h_status[0] != 0) {
    set_error("tensor-core conv pipeline timed out (code " + std::to_string(net->h_status[0]) + ")");
    OCTSEG_CUDA(cudaMemsetAsync(net->d_status, 0, 2 * sizeof(int), net->stream));
    return 1;
  }
  if (net->h_status[1] != 0) {
    set_error("fp32 tensor-core path: an activation exceeded the fp16-pair range (|a| > 65000); results of this "
              "call are invalid, the handle now uses the CUDA-core fp32 path");
    OCTSEG_CUDA(cudaMemsetAsync(net->d_status, 0, 2 * sizeof(int), net->stream));
    net->fp32_path = 1;
    return 2;
  }
  return 0;
}

static int grow(void **p, size_t *cap, size_t need) {
  if (need <= *cap) return 0;
  if (*p) OCTSEG_CUDA(cudaFree(*p));
  *p = nullptr;
  OCTSEG_CUDA(cudaMalloc(p, need));
  *cap = need;
  return 0;
}

static int pick_microbatch(const octseg_net *net, int n, int h, int w) {
  if (net->microbatch > 0) return std::min(n, net->microbatch);
  // activation footprint per image (all planned buffers) must stay under ~32 GB
  const int P = net->cfg.pool_layers, s = net->cfg.start_neurons;
  double per_img = 0;
  for (int l = 0; l <= P; ++l) per_img += 7.0 * (s << l) * (double)(h >> l) * (w >> l);
  per_img *= (net->precision == OCTSEG_FP32 ? 4 : 2);
  int mb = (int)std::max(1.0, std::min((double)n, 32e9 / per_img));
  return mb;
}

}  // namespace octseg

using namespace octseg;

extern "C" void octseg_train_free(octseg_net *net);

// ===================================================================================
// C ABI
// ===================================================================================
extern "C" {

int32_t octseg_version(void) { return 100; }
const char *octseg_last_error(void) { return g_err.c_str(); }

int32_t octseg_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int32_t octseg_param_count(const octseg_config *cfg, int32_t *n_tensors) {
  std::vector<BlockSpec> b; std::vector<ParamSpec> p; int64_t t;
  if (!cfg || !n_tensors) { set_error("null argument"); return 1; }
  if (build_graph(*cfg, &b, &p, &t)) return 1;
  *n_tensors = (int32_t)p.size();
  return 0;
}

int32_t octseg_param_info(const octseg_config *cfg, int32_t index, char *name, int32_t name_cap,
                          int32_t *ndim, int64_t shape[4], int32_t *trainable) {
  std::vector<BlockSpec> b; std::vector<ParamSpec> p; int64_t t;
  if (!cfg) { set_error("null argument"); return 1; }
  if (build_graph(*cfg, &b, &p, &t)) return 1;
  if (index < 0 || index >= (int)p.size()) { set_error("param index out of range"); return 1; }
  const ParamSpec &ps = p[index];
  if (name && name_cap > 0) { std::strncpy(name, ps.name.c_str(), name_cap - 1); name[name_cap - 1] = 0; }
  if (ndim) *ndim = ps.ndim;
  if (shape) for (int i = 0; i < 4; ++i) shape[i] = ps.shape[i];
  if (trainable) *trainable = ps.trainable ? 1 : 0;
  return 0;
}

static int32_t create_impl(const octseg_config *cfg, int32_t device, int32_t precision, octseg_net *net) {
  net->cfg = *cfg; net->device = device; net->precision = precision;
  if (build_graph(*cfg, &net->blocks, &net->params, &net->total_floats) || check_supported(net)) return 1;
  const char *e = std::getenv("OCTSEG_DISABLE_TC");
  net->disable_tc = e && e[0] == '1';
  { const char *rp = std::getenv("OCTSEG_DISABLE_ROWPAIR"); net->disable_rowpair = rp && rp[0] == '1'; }
  e = std::getenv("OCTSEG_DISABLE_FUSION");
  net->disable_fusion = e && e[0] == '1';
  e = std::getenv("OCTSEG_MICROBATCH");
  net->microbatch = e ? std::atoi(e) : 0;
  OCTSEG_CUDA(cudaStreamCreateWithFlags(&net->stream, cudaStreamNonBlocking));
  OCTSEG_CUDA(cudaEventCreateWithFlags(&net->ev_derived, cudaEventDisableTiming));
  OCTSEG_CUDA(cudaMalloc(&net->d_params, net->total_floats * sizeof(float)));
  OCTSEG_CUDA(cudaMemset(net->d_params, 0, net->total_floats * sizeof(float)));
  net->h_params.assign(net->total_floats, 0.f);
  OCTSEG_CUDA(cudaMalloc(&net->d_status, 2 * sizeof(int)));
  OCTSEG_CUDA(cudaMemset(net->d_status, 0, 2 * sizeof(int)));
  OCTSEG_CUDA(cudaMallocHost(&net->h_status, 2 * sizeof(int)));
  { const char *fp = std::getenv("OCTSEG_FP32_PATH"); net->fp32_path = (fp && (fp[0] == 'c' || fp[0] == 'C')) ? 1 : 0; }
  if (init_preprocess_lut()) return 1;
  OCTSEG_CUDA(cudaGetLastError());
  net->bstate.resize(net->blocks.size());
  for (auto &b : net->blocks) {
    BlockState &st = net->bstate[b.index];
    if (b.has_bn) {
      OCTSEG_CUDA(cudaMalloc(&st.scale, b.cout * sizeof(float)));
      OCTSEG_CUDA(cudaMalloc(&st.shift, b.cout * sizeof(float)));
    }
    if (precision != OCTSEG_FP32 && b.index > 0 && b.role != 4 && b.cin % 8 == 0 &&
        tc_supported(b.kh, b.kw, b.cin, b.cout, b.ups, kTcTileH, kTcTileW) &&
        tc_make_geometry(b.kh, b.kw, b.cin, b.cout, b.ups ? 1 : 0, &st.geo) == 0) {
      st.geo_ok = true;
      st.wpack_elems = (size_t)st.geo.n_tiles_n * st.geo.cin_chunks * st.geo.ksteps * 2 * st.geo.n_cols * 8;
      OCTSEG_CUDA(cudaMalloc(&st.wpack, st.wpack_elems * 2));
    }
    if (precision == OCTSEG_FP32 && b.index > 0 && b.role != 4 && b.cin % 8 == 0 &&
        tc_supported(b.kh, b.kw, 2 * b.cin, 2 * b.cout, b.ups, kTcTileH, kTcTileW) &&
        tc_make_geometry_split(b.kh, b.kw, b.cin, b.cout, b.ups ? 1 : 0, &st.geo_s) == 0) {
      st.geo_s_ok = true;
      st.wpack_s_elems = (size_t)st.geo_s.n_tiles_n * st.geo_s.cin_chunks * st.geo_s.ksteps * 2 * st.geo_s.n_cols * 8;
      OCTSEG_CUDA(cudaMalloc(&st.wpack_s, st.wpack_s_elems * 2));
      if (!net->disable_rowpair && !b.ups && b.kh == 3 && b.kw == 3 && (b.cout == 8 || b.cout == 16) &&
          tc_make_geometry_split(4, 3, b.cin, 2 * b.cout, 0, &st.geo_s2, 1, 1) == 0 && st.geo_s2.n_tiles_n == 1) {
        st.geo_s2_ok = true;
        st.geo_s2.rows2 = 1;
        st.wpack_s2_elems = (size_t)st.geo_s2.n_tiles_n * st.geo_s2.cin_chunks * st.geo_s2.ksteps * 2 * st.geo_s2.n_cols * 8;
        OCTSEG_CUDA(cudaMalloc(&st.wpack_s2, st.wpack_s2_elems * 2));
      }
    }
    // row-pair variant (inference): 3x3 layers with 8 or 16 output channels are bound by the smem reads of
    // the A operand; computing two output rows per GEMM row cuts those by a third
    if (st.geo_ok && !net->disable_rowpair && !b.ups && b.kh == 3 && b.kw == 3 && (b.cout == 8 || b.cout == 16) &&
        tc_make_geometry(4, 3, b.cin, 2 * b.cout, 0, &st.geo2, 1, 1) == 0 && st.geo2.n_tiles_n == 1) {
      st.geo2_ok = true;
      st.geo2.rows2 = 1;
      st.wpack2_elems = (size_t)st.geo2.n_tiles_n * st.geo2.cin_chunks * st.geo2.ksteps * 2 * st.geo2.n_cols * 8;
      OCTSEG_CUDA(cudaMalloc(&st.wpack2, st.wpack2_elems * 2));
    }
    OCTSEG_CUDA(cudaGetLastError());
    // tensor-core stem: 3x3, one input channel; GEMM row = 8 adjacent pixels (K = 8 per tap),
    // GEMM columns = 8 pixels x cout channels (a banded weight matrix, see stem_group_weights)
    if (precision != OCTSEG_FP32 && !net->disable_tc && b.index == 0 && b.kh == 3 && b.kw == 3 && b.cin == 1 &&
        b.cout <= 32 && tc_make_geometry(3, 3, 8, 8 * b.cout, 0, &st.geo) == 0) {
      st.geo_ok = true;
      st.geo.stem_groups = 1;
      st.wpack_elems = (size_t)st.geo.n_tiles_n * st.geo.cin_chunks * st.geo.ksteps * 2 * st.geo.n_cols * 8;
      OCTSEG_CUDA(cudaMalloc(&st.wpack, st.wpack_elems * 2));
      OCTSEG_CUDA(cudaMalloc(&st.rep_scale, 8 * b.cout * sizeof(float)));
      OCTSEG_CUDA(cudaMalloc(&st.rep_shift, 8 * b.cout * sizeof(float)));
    }
  }
  return 0;
}


int32_t octseg_destroy(octseg_net *net);

int32_t octseg_create(const octseg_config *cfg, int32_t device, int32_t precision, octseg_net **out) {
  if (!cfg || !out) { set_error("null argument"); return 1; }
  if (precision != OCTSEG_FP32 && precision != OCTSEG_BF16 && precision != OCTSEG_FP16) { set_error("bad precision"); return 1; }
  int ndev = octseg_device_count();
  if (ndev <= 0) { set_error("no CUDA device: liboctseg has no CPU fallback"); return 1; }
  if (device < 0 || device >= ndev) { set_error("device index out of range"); return 1; }
  OCTSEG_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  OCTSEG_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) { set_error("liboctseg is built for sm_100a (B200) only"); return 1; }
  octseg_net *net = new octseg_net();
  if (create_impl(cfg, device, precision, net)) {
    // every early return above leaves a partly built handle: release whatever exists (the error string is kept)
    const std::string err = g_err;
    octseg_destroy(net);
    g_err = err;
    return 1;
  }
  OCTSEG_CUDA(cudaGetLastError());
  *out = net;
  return 0;
}

int32_t octseg_destroy(octseg_net *net) {
  if (!net) return 0;
  cudaSetDevice(net->device);
  if (net->stream) cudaStreamSynchronize(net->stream);
  octseg_train_free(net);
  for (auto &e : net->prof_events) cudaEventDestroy(e);
  for (auto &S : net->slot) {
    for (auto &e : S.ev) cudaEventDestroy(e);
    if (S.ev_start) cudaEventDestroy(S.ev_start);
    if (S.ev_done) cudaEventDestroy(S.ev_done);
    if (S.ev_status) cudaEventDestroy(S.ev_status);
    if (S.h_status) cudaFreeHost(S.h_status);
    cudaFree(S.d_img); cudaFree(S.d_probs); cudaFree(S.d_labels); cudaFree(S.d_maps);
  }
  if (net->copy_in) cudaStreamDestroy(net->copy_in);
  if (net->copy_out) cudaStreamDestroy(net->copy_out);
  for (auto &st : net->bstate) { cudaFree(st.scale); cudaFree(st.shift); cudaFree(st.wpack); cudaFree(st.rep_scale); cudaFree(st.rep_shift); cudaFree(st.wpack2); cudaFree(st.wpack_s); cudaFree(st.wpack_s2); }
  cudaFree(net->d_params); cudaFree(net->ws); cudaFree(net->d_eval); cudaFree(net->d_status);
  if (net->h_status) cudaFreeHost(net->h_status);
  if (net->ev_derived) cudaEventDestroy(net->ev_derived);
  if (net->stream) cudaStreamDestroy(net->stream);
  delete net;
  return 0;
}

int32_t octseg_set_param(octseg_net *net, int32_t index, const float *host, int64_t count) {
  if (!net || !host) { set_error("null argument"); return 1; }
  if (index < 0 || index >= (int)net->params.size()) { set_error("param index out of range"); return 1; }
  const ParamSpec &p = net->params[index];
  if (count != p.count) { set_error("param " + p.name + ": element count mismatch"); return 1; }
  OCTSEG_CUDA(cudaSetDevice(net->device));
  if (sync_host_mirror(net)) return 1;
  std::memcpy(net->h_params.data() + p.offset, host, count * sizeof(float));
  OCTSEG_CUDA(cudaMemcpyAsync(net->d_params + p.offset, net->h_params.data() + p.offset,
                              count * sizeof(float), cudaMemcpyHostToDevice, net->stream));
  OCTSEG_CUDA(cudaStreamSynchronize(net->stream));
  net->derived_dirty = true;
  return 0;
}

int32_t octseg_get_param(octseg_net *net, int32_t index, float *host, int64_t count) {
  if (!net || !host) { set_error("null argument"); return 1; }
  if (index < 0 || index >= (int)net->params.size()) { set_error("param index out of range"); return 1; }
  const ParamSpec &p = net->params[index];
  if (count != p.count) { set_error("param " + p.name + ": element count mismatch"); return 1; }
  OCTSEG_CUDA(cudaSetDevice(net->device));
  if (sync_host_mirror(net)) return 1;
  std::memcpy(host, net->h_params.data() + p.offset, count * sizeof(float));
  return 0;
}

int32_t octseg_predict_device(octseg_net *net, const void *images, int32_t dtype, int32_t n, int32_t h,
                              int32_t w, float *probs, uint8_t *labels, void *stream) {
  if (!net || !images) { set_error("null argument"); return 1; }
  if (n <= 0 || h <= 0 || w <= 0) { set_error("bad image batch shape"); return 1; }
  if (dtype != OCTSEG_U8 && dtype != OCTSEG_F32 && dtype != OCTSEG_F32_PRE) { set_error("bad image dtype"); return 1; }
  OCTSEG_CUDA(cudaSetDevice(net->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : net->stream;
  const int mb = pick_microbatch(net, n, h, w);
  const size_t img_elem = (dtype == OCTSEG_U8 ? 1 : 4) * (size_t)h * w * net->cfg.input_channels;
  for (int i0 = 0; i0 < n; i0 += mb) {
    // keep micro-batches equal-sized where possible so the workspace plan is reused
    const int cur = std::min(mb, n - i0);
    const uint8_t *img = reinterpret_cast<const uint8_t *>(images) + (size_t)i0 * img_elem;
    float *pr = probs ? probs + (size_t)i0 * h * w * net->cfg.num_classes : nullptr;
    uint8_t *lb = labels ? labels + (size_t)i0 * h * w : nullptr;
    if (forward(net, img, dtype, cur, h, w, pr, lb, st)) return 1;
  }
  return 0;
}

// Chunked three-stage pipeline shared by the host-buffer entry points: H2D of chunk i+1, forward (+ boundary
// maps) of chunk i and D2H of chunk i-1 overlap on three streams (PCIe is full duplex).  Any of
// probs / labels / maps may be null; maps need labels on the device but not necessarily on the host.
// The call uses one of two staging slots and only ENQUEUES work; predict_wait() completes it.  Two calls in flight
// (submit i+1 before waiting for i) overlap the H2D of batch i+1 with the forward and D2H of batch i.
static int predict_wait(octseg_net *net, int si);
__global__ void snapshot_status_kernel(const int *__restrict__ dev, int *host_pinned) {
  if (threadIdx.x < 2) host_pinned[threadIdx.x] = dev[threadIdx.x];
  __threadfence_system();
}

static int predict_enqueue(octseg_net *net, int si, const void *images, int32_t dtype, int32_t n, int32_t h, int32_t w,
                           float *probs, uint8_t *labels, uint8_t *maps, int bg_ilm, int bg_csi, int transposed,
                           bool pipelined = false) {
  if (n <= 0 || h <= 0 || w <= 0) { set_error("bad image batch shape"); return 1; }
  if (dtype != OCTSEG_U8 && dtype != OCTSEG_F32 && dtype != OCTSEG_F32_PRE) { set_error("bad image dtype"); return 1; }
  OCTSEG_CUDA(cudaSetDevice(net->device));
  OCTSEG_CUDA(cudaGetLastError());                        // an earlier unchecked failure must not be blamed on this call
  octseg_net::HostSlot &S = net->slot[si];
  if (S.busy && predict_wait(net, si)) return 1;          // the slot's previous call must have drained
  const int K = net->cfg.num_classes;
  const size_t img_per = (dtype == OCTSEG_U8 ? 1 : 4) * (size_t)h * w * net->cfg.input_channels;
  const size_t pr_per = (size_t)h * w * K * sizeof(float), lb_per = (size_t)h * w, mp_per = lb_per * (K - 1);
  const bool need_labels = labels || maps;
  if (grow(&S.d_img, &S.d_img_bytes, img_per * n)) return 1;
  if (probs && grow(reinterpret_cast<void **>(&S.d_probs), &S.d_probs_bytes, pr_per * n)) return 1;
  if (need_labels && grow(reinterpret_cast<void **>(&S.d_labels), &S.d_labels_bytes, lb_per * n)) return 1;
  if (maps && grow(reinterpret_cast<void **>(&S.d_maps), &S.d_maps_bytes, mp_per * n)) return 1;
  // Chunk size: total ~ H2D(chunk) + sum of forwards + D2H(last chunk), and a forward costs a fixed ~0.2 ms
  // (22 launches) plus ~0.022 ms per 512x512 B-scan.  With fp32 probabilities going back (4*K bytes per
  // pixel) the call is D2H-bound and small chunks start the return traffic early; with only label / boundary
  // maps (1 + K-1 bytes per pixel) the forwards dominate and fewer, larger chunks (~22 B-scans) win (measured on B200).
  // A pipelined caller (submit / wait) overlaps whole calls instead, so the batch is forwarded in one piece: a chunked
  // forward pays the fixed per-launch cost three times (3 x 0.68 ms instead of 1.17 ms per 64 B-scans).
  int chunk;
  if (net->microbatch > 0) chunk = std::min(n, net->microbatch);
  else if (pipelined) chunk = n;
  else if (probs) chunk = std::min(n, 8);
  else if (n < 16) chunk = n;
  else { const int nc = std::max(2, (n + 21) / 22); chunk = (n + nc - 1) / nc; }
  chunk = std::max(1, std::min(chunk, pick_microbatch(net, n, h, w)));   // workspace memory bound
  if (!net->copy_in) {
    OCTSEG_CUDA(cudaStreamCreateWithFlags(&net->copy_in, cudaStreamNonBlocking));
    OCTSEG_CUDA(cudaStreamCreateWithFlags(&net->copy_out, cudaStreamNonBlocking));
  }
  if (!S.ev_start) {
    OCTSEG_CUDA(cudaEventCreateWithFlags(&S.ev_start, cudaEventDisableTiming));
    OCTSEG_CUDA(cudaEventCreateWithFlags(&S.ev_done, cudaEventDisableTiming));
    OCTSEG_CUDA(cudaEventCreateWithFlags(&S.ev_status, cudaEventDisableTiming));
    OCTSEG_CUDA(cudaMallocHost(&S.h_status, 2 * sizeof(int)));
  }
  const int n_chunks = (n + chunk - 1) / chunk;
  if ((int)S.ev.size() < 2 * n_chunks) {
    const size_t old = S.ev.size();
    S.ev.resize(2 * n_chunks);
    for (size_t i = old; i < S.ev.size(); ++i) OCTSEG_CUDA(cudaEventCreateWithFlags(&S.ev[i], cudaEventDisableTiming));
  }
  // The upload does not wait for earlier work on the compute stream (that would serialise H2D(i+1) behind forward(i)):
  // it touches only this slot's image buffer, and the slot was drained above.
  for (int c = 0; c < n_chunks; ++c) {
    const int i0 = c * chunk, cur = std::min(chunk, n - i0);
    cudaEvent_t ev_in = S.ev[2 * c], ev_fw = S.ev[2 * c + 1];
    uint8_t *dimg = reinterpret_cast<uint8_t *>(S.d_img) + (size_t)i0 * img_per;
    OCTSEG_CUDA(cudaMemcpyAsync(dimg, reinterpret_cast<const uint8_t *>(images) + (size_t)i0 * img_per,
                                (size_t)cur * img_per, cudaMemcpyHostToDevice, net->copy_in));
    OCTSEG_CUDA(cudaEventRecord(ev_in, net->copy_in));
    OCTSEG_CUDA(cudaStreamWaitEvent(net->stream, ev_in, 0));
    float *dpr = probs ? S.d_probs + (size_t)i0 * h * w * K : nullptr;
    uint8_t *dlb = need_labels ? S.d_labels + (size_t)i0 * lb_per : nullptr;
    uint8_t *dmp = maps ? S.d_maps + (size_t)i0 * mp_per : nullptr;
    if (forward(net, dimg, dtype, cur, h, w, dpr, dlb, net->stream)) return 1;
    if (maps) {
      if (launch_boundary_maps(dlb, cur, h, w, K, bg_ilm, bg_csi, transposed, dmp, net->stream)) return 1;
      ++net->launches;
    }
    OCTSEG_CUDA(cudaEventRecord(ev_fw, net->stream));
    OCTSEG_CUDA(cudaStreamWaitEvent(net->copy_out, ev_fw, 0));
    if (probs)
      OCTSEG_CUDA(cudaMemcpyAsync(probs + (size_t)i0 * h * w * K, dpr, (size_t)cur * pr_per, cudaMemcpyDeviceToHost,
                                  net->copy_out));
    if (labels)
      OCTSEG_CUDA(cudaMemcpyAsync(labels + (size_t)i0 * lb_per, dlb, (size_t)cur * lb_per, cudaMemcpyDeviceToHost,
                                  net->copy_out));
    if (maps)
      OCTSEG_CUDA(cudaMemcpyAsync(maps + (size_t)i0 * mp_per, dmp, (size_t)cur * mp_per, cudaMemcpyDeviceToHost,
                                  net->copy_out));
  }
  // Status words as they stand after THIS call's last forward (later calls queued behind it are not waited for).
  // Written by a one-thread kernel straight into pinned host memory: a cudaMemcpyAsync here would queue behind the
  // big D2H copies of the previous call on the copy engine and stall the compute stream with it.
  snapshot_status_kernel<<<1, 32, 0, net->stream>>>(net->d_status, S.h_status);
  OCTSEG_CUDA(cudaGetLastError());
  OCTSEG_CUDA(cudaEventRecord(S.ev_status, net->stream));
  OCTSEG_CUDA(cudaEventRecord(S.ev_done, net->copy_out));
  S.busy = true;
  S.images = images; S.dtype = dtype; S.n = n; S.h = h; S.w = w; S.probs = probs; S.labels = labels; S.maps = maps;
  S.bg_ilm = bg_ilm; S.bg_csi = bg_csi; S.transposed = transposed; S.pipelined = pipelined;
  return 0;
}

static int predict_wait(octseg_net *net, int si) {
  octseg_net::HostSlot &S = net->slot[si];
  if (!S.busy) return 0;
  OCTSEG_CUDA(cudaSetDevice(net->device));
  OCTSEG_CUDA(cudaEventSynchronize(S.ev_status));
  OCTSEG_CUDA(cudaEventSynchronize(S.ev_done));
  S.busy = false;
  if (S.h_status[0] != 0) {
    set_error("tensor-core conv pipeline timed out (code " + std::to_string(S.h_status[0]) + ")");
    OCTSEG_CUDA(cudaMemsetAsync(net->d_status, 0, 2 * sizeof(int), net->stream));
    return 1;
  }
  if (S.h_status[1] != 0) {
    // fp16-pair range overflow of the fp32 tensor-core path: switch the handle to the CUDA-core fp32 path for good
    // and run this call again (anything else in flight is re-run by its own wait)
    OCTSEG_CUDA(cudaStreamSynchronize(net->stream));
    OCTSEG_CUDA(cudaMemsetAsync(net->d_status, 0, 2 * sizeof(int), net->stream));
    net->fp32_path = 1;
    if (predict_enqueue(net, si, S.images, S.dtype, S.n, S.h, S.w, S.probs, S.labels, S.maps, S.bg_ilm, S.bg_csi, S.transposed,
                        S.pipelined))
      return 1;
    return predict_wait(net, si);
  }
  return 0;
}

static int predict_pipeline(octseg_net *net, const void *images, int32_t dtype, int32_t n, int32_t h, int32_t w,
                            float *probs, uint8_t *labels, uint8_t *maps, int bg_ilm, int bg_csi, int transposed) {
  const int si = net->next_slot;
  net->next_slot ^= 1;
  if (predict_enqueue(net, si, images, dtype, n, h, w, probs, labels, maps, bg_ilm, bg_csi, transposed)) return 1;
  return predict_wait(net, si);
}

int32_t octseg_predict_host(octseg_net *net, const void *images, int32_t dtype, int32_t n, int32_t h,
                            int32_t w, float *probs, uint8_t *labels) {
  if (!net || !images) { set_error("null argument"); return 1; }
  return predict_pipeline(net, images, dtype, n, h, w, probs, labels, nullptr, 0, 0, 0);
}

int32_t octseg_predict_maps_host(octseg_net *net, const void *images, int32_t dtype, int32_t n, int32_t h,
                                 int32_t w, int32_t bg_ilm, int32_t bg_csi, int32_t transposed, uint8_t *labels,
                                 uint8_t *maps) {
  if (!net || !images || !maps) { set_error("null argument"); return 1; }
  return predict_pipeline(net, images, dtype, n, h, w, nullptr, labels, maps, bg_ilm, bg_csi, transposed);
}

// Validation pass on the device: forward + per-image Dice counts + weighted-CE sums; only 24*K + 8 bytes per image come back.
int32_t octseg_evaluate_host(octseg_net *net, const void *images, int32_t dtype, const uint8_t *labels, int32_t n,
                             int32_t h, int32_t w, const float *class_weights, int64_t *counts, double *loss_sums) {
  if (!net || !images || !labels || !counts || !loss_sums) { set_error("null argument"); return 1; }
  if (n <= 0 || h <= 0 || w <= 0) { set_error("bad image batch shape"); return 1; }
  if (dtype != OCTSEG_U8 && dtype != OCTSEG_F32 && dtype != OCTSEG_F32_PRE) { set_error("bad image dtype"); return 1; }
  OCTSEG_CUDA(cudaSetDevice(net->device));
  const int K = net->cfg.num_classes;
  if (K < 2 || K > 16) { set_error("evaluate: num_classes must be 2..16"); return 1; }
  const size_t img_per = (dtype == OCTSEG_U8 ? 1 : 4) * (size_t)h * w * net->cfg.input_channels;
  const size_t lb_per = (size_t)h * w;
  const int chunk = std::max(1, std::min(std::min(n, 32), pick_microbatch(net, n, h, w)));
  for (int si = 0; si < 2; ++si)
    if (predict_wait(net, si)) return 1;
  octseg_net::HostSlot &S0 = net->slot[0];
  if (grow(&S0.d_img, &S0.d_img_bytes, img_per * chunk)) return 1;
  if (grow(reinterpret_cast<void **>(&S0.d_probs), &S0.d_probs_bytes, lb_per * K * sizeof(float) * chunk)) return 1;
  const size_t off_cnt = (lb_per * chunk + 255) & ~(size_t)255, off_loss = off_cnt + (size_t)n * 3 * K * 8;
  const size_t off_cw = off_loss + (size_t)n * 8, total = off_cw + 64 * sizeof(float);
  if (grow(&net->d_eval, &net->d_eval_bytes, total)) return 1;
  uint8_t *eb = reinterpret_cast<uint8_t *>(net->d_eval);
  unsigned long long *d_cnt = reinterpret_cast<unsigned long long *>(eb + off_cnt);
  double *d_loss = reinterpret_cast<double *>(eb + off_loss);
  float *d_cw = reinterpret_cast<float *>(eb + off_cw);
  OCTSEG_CUDA(cudaMemsetAsync(eb + off_cnt, 0, off_cw - off_cnt, net->stream));
  if (class_weights)
    OCTSEG_CUDA(cudaMemcpyAsync(d_cw, class_weights, K * sizeof(float), cudaMemcpyHostToDevice, net->stream));
  for (int i0 = 0; i0 < n; i0 += chunk) {
    const int cur = std::min(chunk, n - i0);
    OCTSEG_CUDA(cudaMemcpyAsync(S0.d_img, reinterpret_cast<const uint8_t *>(images) + (size_t)i0 * img_per, cur * img_per,
                                cudaMemcpyHostToDevice, net->stream));
    OCTSEG_CUDA(cudaMemcpyAsync(eb, labels + (size_t)i0 * lb_per, cur * lb_per, cudaMemcpyHostToDevice, net->stream));
    if (forward(net, S0.d_img, dtype, cur, h, w, S0.d_probs, nullptr, net->stream)) return 1;
    if (launch_eval_counts(S0.d_probs, eb, cur, h, w, K, class_weights ? d_cw : nullptr, d_cnt + (size_t)i0 * 3 * K,
                           d_loss + i0, net->stream))
      return 1;
    ++net->launches;
  }
  static_assert(sizeof(unsigned long long) == sizeof(int64_t), "count width");
  OCTSEG_CUDA(cudaMemcpyAsync(counts, d_cnt, (size_t)n * 3 * K * 8, cudaMemcpyDeviceToHost, net->stream));
  OCTSEG_CUDA(cudaMemcpyAsync(loss_sums, d_loss, (size_t)n * 8, cudaMemcpyDeviceToHost, net->stream));
  const int rc = check_status(net);
  if (rc == 2) return octseg_evaluate_host(net, images, dtype, labels, n, h, w, class_weights, counts, loss_sums);
  return rc;
}

// Asynchronous pair of octseg_predict_maps_host for pipelining consecutive batches: submit() only enqueues (host
// buffers must be PINNED and stay valid until the matching wait()); at most two calls may be in flight.
int32_t octseg_predict_maps_submit(octseg_net *net, const void *images, int32_t dtype, int32_t n, int32_t h, int32_t w,
                                   int32_t bg_ilm, int32_t bg_csi, int32_t transposed, uint8_t *labels, uint8_t *maps,
                                   int32_t *ticket) {
  if (!net || !images || !maps || !ticket) { set_error("null argument"); return 1; }
  const int si = net->next_slot;
  net->next_slot ^= 1;
  if (predict_enqueue(net, si, images, dtype, n, h, w, nullptr, labels, maps, bg_ilm, bg_csi, transposed, true)) return 1;
  *ticket = si;
  return 0;
}

int32_t octseg_predict_wait(octseg_net *net, int32_t ticket) {
  if (!net || ticket < 0 || ticket > 1) { set_error("bad ticket"); return 1; }
  return predict_wait(net, ticket);
}

int32_t octseg_synchronize(octseg_net *net) {
  if (!net) { set_error("null argument"); return 1; }
  OCTSEG_CUDA(cudaSetDevice(net->device));
  for (int si = 0; si < 2; ++si)
    if (predict_wait(net, si)) return 1;
  return check_status(net);
}

int64_t octseg_launch_count(octseg_net *net) { return net ? net->launches : 0; }

int32_t octseg_set_profiling(octseg_net *net, int32_t enable) {
  if (!net) { set_error("null argument"); return 1; }
  OCTSEG_CUDA(cudaSetDevice(net->device));
  if (enable && net->prof_events.empty()) {
    net->prof_events.resize(net->blocks.size() + 1);
    for (auto &e : net->prof_events) OCTSEG_CUDA(cudaEventCreate(&e));
  }
  net->profiling = enable != 0;
  net->prof_valid = false;
  return 0;
}

int32_t octseg_get_block_times(octseg_net *net, float *ms, int32_t cap, int32_t *n_blocks) {
  if (!net || !ms) { set_error("null argument"); return 1; }
  if (!net->prof_valid) { set_error("no profiled forward pass yet"); return 1; }
  OCTSEG_CUDA(cudaSetDevice(net->device));
  OCTSEG_CUDA(cudaEventSynchronize(net->prof_events.back()));
  const int nb = (int)net->blocks.size();
  if (n_blocks) *n_blocks = nb;
  for (int i = 0; i < nb && i < cap; ++i)
    OCTSEG_CUDA(cudaEventElapsedTime(&ms[i], net->prof_events[i], net->prof_events[i + 1]));
  return 0;
}

int32_t octseg_layer_uses_tensor_core(octseg_net *net, int32_t conv_index, int32_t h, int32_t w) {
  if (!net || conv_index < 0 || conv_index >= (int)net->blocks.size()) return 0;
  const BlockSpec &b = net->blocks[conv_index];
  if (net->disable_tc || b.role == 4) return 0;
  if (net->precision == OCTSEG_FP32) return (b.index > 0 && split_applicable(net, h, w)) ? 1 : 0;
  if (!net->bstate[b.index].geo_ok) return 0;
  if (b.index == 0)      // tensor-core stem over pixel groups (uint8 input only)
    return ((w % 8) == 0 && tc_supported(3, 3, 8, 8 * b.cout, 0, h, w / 8)) ? 1 : 0;
  int lh = h >> b.level, lw = w >> b.level;
  if (b.ups) { lh >>= 1; lw >>= 1; }
  return tc_supported(b.kh, b.kw, b.cin, b.cout, b.ups, lh, lw) ? 1 : 0;
}

// ---- per-block debug entry (tests): NHWC fp32 host in/out ---------------------------
int32_t octseg_debug_conv_block(octseg_net *net, int32_t conv_index, int32_t path, const float *in,
                                int32_t n, int32_t h, int32_t w, float *out, float *ms_out) {
  if (!net || !in || !out) { set_error("null argument"); return 1; }
  if (conv_index <= 0 || conv_index >= (int)net->blocks.size()) { set_error("bad conv index"); return 1; }
  const BlockSpec &b = net->blocks[conv_index];
  if (b.role == 4) { set_error("head is not a conv block"); return 1; }
  OCTSEG_CUDA(cudaSetDevice(net->device));
  if (prepare_derived(net)) return 1;
  const int oh = b.ups ? 2 * h : h, ow = b.ups ? 2 * w : w;
  // fp32 mode, path 1: the tensor-core path on error-compensated fp16 pairs (two physical planes per logical plane)
  const bool split = (path == 1 && net->precision == OCTSEG_FP32);
  const size_t es = split ? 4 : elem_size(net);       // bytes per LOGICAL element
  const size_t in_elems = (size_t)n * b.cin * h * w, out_elems = (size_t)n * b.cout * oh * ow;
  // host re-layout NHWC -> [N][C/8][H][W][8]
  std::vector<float> blk(in_elems);
  for (int i = 0; i < n; ++i)
    for (int y = 0; y < h; ++y)
      for (int x = 0; x < w; ++x)
        for (int c = 0; c < b.cin; ++c)
          blk[((((size_t)i * (b.cin / 8) + c / 8) * h + y) * w + x) * 8 + (c & 7)] =
              in[(((size_t)i * h + y) * w + x) * b.cin + c];
  void *d_in = nullptr, *d_out = nullptr;
  OCTSEG_CUDA(cudaMalloc(&d_in, in_elems * es));
  OCTSEG_CUDA(cudaMalloc(&d_out, out_elems * es));
  OCTSEG_CUDA(cudaMemset(d_out, 0xFF, out_elems * es));   // poison: unwritten outputs show up as NaN
  std::vector<uint16_t> h16;
  if (split) {
    h16.resize(2 * in_elems);
    const size_t pe = (size_t)h * w * 8;
    for (size_t pl = 0; pl < (size_t)n * (b.cin / 8); ++pl)
      for (size_t i = 0; i < pe; ++i) {
        const float a = blk[pl * pe + i];
        const __half hi = __float2half_rn(a);
        const __half lo = __float2half_rn((a - __half2float(hi)) * 2048.f);
        std::memcpy(&h16[(2 * pl) * pe + i], &hi, 2);
        std::memcpy(&h16[(2 * pl + 1) * pe + i], &lo, 2);
      }
    OCTSEG_CUDA(cudaMemcpy(d_in, h16.data(), in_elems * 4, cudaMemcpyHostToDevice));
  } else if (net->precision != OCTSEG_FP32) {
    h16.resize(in_elems);
    for (size_t i = 0; i < in_elems; ++i) {
      if (net->precision == OCTSEG_BF16) { __nv_bfloat16 v = __float2bfloat16(blk[i]); std::memcpy(&h16[i], &v, 2); }
      else { __half v = __float2half(blk[i]); std::memcpy(&h16[i], &v, 2); }
    }
    OCTSEG_CUDA(cudaMemcpy(d_in, h16.data(), in_elems * 2, cudaMemcpyHostToDevice));
  } else {
    OCTSEG_CUDA(cudaMemcpy(d_in, blk.data(), in_elems * 4, cudaMemcpyHostToDevice));
  }
  const BlockState &bs = net->bstate[b.index];
  const float *P = net->d_params;
  cudaEvent_t e0, e1;
  OCTSEG_CUDA(cudaEventCreate(&e0));
  OCTSEG_CUDA(cudaEventCreate(&e1));
  int rc = 0;
  const int reps = ms_out ? 5 : 1;
  TcPlan plan;
  if (split) {
    if (!bs.geo_s_ok || !tc_supported(b.kh, b.kw, 2 * b.cin, 2 * b.cout, b.ups, h, w)) {
      set_error("tensor-core (split fp32) path not available for this block/shape");
      rc = 1;
    } else {
      TcEpilogue epi;
      epi.fp16 = 1;
      epi.static_weights = 1;
      epi.overflow = net->d_status + 1;
      epi.scale = bs.scale; epi.shift = bs.shift;
      epi.out = make_view(reinterpret_cast<__nv_bfloat16 *>(d_out), n, b.cout / 4, 0, b.cout / 4, oh, ow);
      rc = tc_make_plan(bs.geo_s, reinterpret_cast<const __nv_bfloat16 *>(d_in), n, h, w, bs.wpack_s, epi,
                        net->d_status, &plan);
    }
  } else if (path == 1) {
    if (!bs.geo_ok || !tc_supported(b.kh, b.kw, b.cin, b.cout, b.ups, h, w)) {
      set_error("tensor-core path not available for this block/shape");
      rc = 1;
    } else {
      TcEpilogue epi;
      epi.fp16 = net->precision == OCTSEG_FP16;
      epi.static_weights = 1;
      epi.scale = bs.scale; epi.shift = bs.shift;
      epi.out = make_view(reinterpret_cast<__nv_bfloat16 *>(d_out), n, b.cout / 8, 0, b.cout / 8, oh, ow);
      rc = tc_make_plan(bs.geo, reinterpret_cast<const __nv_bfloat16 *>(d_in), n, h, w, bs.wpack, epi,
                        net->d_status, &plan);
    }
  }
  long long *d_dbg = nullptr;
  if (path == 1 && rc == 0 && std::getenv("OCTSEG_TC_DEBUG")) {
    cudaMalloc(&d_dbg, 64);
    cudaMemset(d_dbg, 0, 64);
    plan.p.dbg = d_dbg;
    fprintf(stderr, "[tc] block %d: mt %dx%d tiles %d grid %d a_stages %d (%u B) b_res %d b_stages %d (%u B) n_cols %d ksteps %d chunks %d smem %zu\n",
            conv_index, plan.p.mt_x, plan.p.mt_y, plan.p.num_tiles, plan.grid, plan.p.a_stages, plan.p.a_stage_bytes,
            plan.p.b_resident, plan.p.b_stages, plan.p.b_stage_bytes, plan.p.n_cols, plan.p.ksteps, plan.p.cin_chunks,
            plan.smem_bytes);
  }
  for (int rep = 0; rc == 0 && rep < reps; ++rep) {
    if (rep == reps - 1) cudaEventRecord(e0, net->stream);
    if (path == 1) {
      rc = tc_launch(plan, net->stream);
    } else if (net->precision == OCTSEG_FP16) {
      View<const __half> iv = make_view(reinterpret_cast<const __half *>(d_in), n, b.cin / 8, 0, b.cin / 8, h, w);
      View<__half> ov = make_view(reinterpret_cast<__half *>(d_out), n, b.cout / 8, 0, b.cout / 8, oh, ow);
      rc = launch_conv_direct<__half>(iv, P + net->params[b.p_kernel].offset, b.kh, b.kw, b.cin, b.cout,
                                      b.ups ? 1 : 0, bs.scale, bs.shift, 1, ov, net->stream);
    } else if (net->precision == OCTSEG_BF16) {
      View<const __nv_bfloat16> iv = make_view(reinterpret_cast<const __nv_bfloat16 *>(d_in), n, b.cin / 8, 0, b.cin / 8, h, w);
      View<__nv_bfloat16> ov = make_view(reinterpret_cast<__nv_bfloat16 *>(d_out), n, b.cout / 8, 0, b.cout / 8, oh, ow);
      rc = launch_conv_direct<__nv_bfloat16>(iv, P + net->params[b.p_kernel].offset, b.kh, b.kw, b.cin, b.cout,
                                             b.ups ? 1 : 0, bs.scale, bs.shift, 1, ov, net->stream);
    } else {
      View<const float> iv = make_view(reinterpret_cast<const float *>(d_in), n, b.cin / 8, 0, b.cin / 8, h, w);
      View<float> ov = make_view(reinterpret_cast<float *>(d_out), n, b.cout / 8, 0, b.cout / 8, oh, ow);
      rc = launch_conv_direct<float>(iv, P + net->params[b.p_kernel].offset, b.kh, b.kw, b.cin, b.cout,
                                     b.ups ? 1 : 0, bs.scale, bs.shift, 1, ov, net->stream);
    }
    ++net->launches;
    if (rep == reps - 1) cudaEventRecord(e1, net->stream);
  }
  if (rc == 0) rc = check_status(net);
  if (rc == 0 && ms_out) cudaEventElapsedTime(ms_out, e0, e1);
  if (d_dbg) {
    long long h[8];
    cudaMemcpy(h, d_dbg, 64, cudaMemcpyDeviceToHost);
    fprintf(stderr, "[tc] block0 cycles: producer wait_a_empty %lld of %lld | issuer0 wait_acc_empty %lld wait_a_full %lld wait_b_full %lld of %lld | epilogue wait_acc_full %lld of %lld\n",
            h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7]);
    cudaFree(d_dbg);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (rc == 0) {
    std::vector<float> ob(out_elems);
    if (split) {
      std::vector<uint16_t> o16(2 * out_elems);
      cudaMemcpy(o16.data(), d_out, out_elems * 4, cudaMemcpyDeviceToHost);
      const size_t pe = (size_t)oh * ow * 8;
      for (size_t pl = 0; pl < (size_t)n * (b.cout / 8); ++pl)
        for (size_t i = 0; i < pe; ++i) {
          __half hi, lo;
          std::memcpy(&hi, &o16[(2 * pl) * pe + i], 2);
          std::memcpy(&lo, &o16[(2 * pl + 1) * pe + i], 2);
          ob[pl * pe + i] = __half2float(hi) + __half2float(lo) * 4.8828125e-4f;
        }
    } else if (net->precision != OCTSEG_FP32) {
      std::vector<uint16_t> o16(out_elems);
      cudaMemcpy(o16.data(), d_out, out_elems * 2, cudaMemcpyDeviceToHost);
      for (size_t i = 0; i < out_elems; ++i) {
        if (net->precision == OCTSEG_BF16) { uint32_t u = (uint32_t)o16[i] << 16; std::memcpy(&ob[i], &u, 4); }
        else { __half hv; std::memcpy(&hv, &o16[i], 2); ob[i] = __half2float(hv); }
      }
    } else {
      cudaMemcpy(ob.data(), d_out, out_elems * 4, cudaMemcpyDeviceToHost);
    }
    for (int i = 0; i < n; ++i)
      for (int y = 0; y < oh; ++y)
        for (int x = 0; x < ow; ++x)
          for (int c = 0; c < b.cout; ++c)
            out[(((size_t)i * oh + y) * ow + x) * b.cout + c] =
                ob[((((size_t)i * (b.cout / 8) + c / 8) * oh + y) * ow + x) * 8 + (c & 7)];
  }
  cudaFree(d_in);
  cudaFree(d_out);
  return rc;
}

}  // extern "C"
