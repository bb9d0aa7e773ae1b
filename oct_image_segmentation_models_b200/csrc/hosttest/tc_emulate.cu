// Host-only emulation of conv_tc_kernel's address arithmetic: builds the smem halo tile the
// TMA box would deliver, reads it back through the (a_off, LBO, SBO) descriptor table and the
// packed-weight image exactly as the UMMA would, and compares with a direct convolution.
// Validates geometry + weight packing + epilogue indexing without a GPU.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>
#include "../conv_tc.cuh"

using namespace octseg;
namespace octseg { void set_error(const std::string &m) { fprintf(stderr, "error: %s\n", m.c_str()); } }

static float bf2f(uint16_t h) { uint32_t u = (uint32_t)h << 16; float f; memcpy(&f, &u, 4); return f; }
static uint16_t f2bf(float f) { uint32_t u; memcpy(&u, &f, 4); u += 0x7fffu + ((u >> 16) & 1u); return (uint16_t)(u >> 16); }

// rowpair: the inference / training row-pair variant of a 3x3 layer (two output rows per GEMM row, banded 4x3
// filter from tc_rowpair_weights, descriptor SBO = 2 x row pitch)
static int run(int kh, int kw, int cin, int cout, int ups, int n, int h, int w, int rowpair = 0) {
  TcGeometry g;
  if (rowpair) {
    if (tc_make_geometry(4, 3, cin, 2 * cout, 0, &g, 1, 1)) { printf("geometry failed\n"); return 1; }
    g.rows2 = 1;
  } else
  if (tc_make_geometry(kh, kw, cin, cout, ups, &g)) { printf("geometry failed\n"); return 1; }
  TcConvParams p; size_t smem;
  if (tc_fill_params(g, n, h, w, &p, &smem)) return 1;
  std::mt19937 rng(1);
  std::uniform_real_distribution<float> U(-1, 1);
  std::vector<float> wt((size_t)kh * kw * cin * cout);
  for (auto &v : wt) v = U(rng);
  std::vector<uint16_t> x((size_t)n * cin * h * w);   // blocked [n][cg][h][w][8]
  for (auto &v : x) v = f2bf(U(rng));
  std::vector<uint16_t> wp;
  if (rowpair) {
    std::vector<float> banded;
    tc_rowpair_weights(wt.data(), cin, cout, &banded);
    tc_pack_weights(g, banded.data(), &wp);
  } else
  tc_pack_weights(g, wt.data(), &wp);
  const int RM = p.row_mul;
  const int oh = ups ? 2 * h : h, ow = ups ? 2 * w : w, cg = cin / 8;
  std::vector<double> out((size_t)n * oh * ow * cout, 1e30), ref((size_t)n * oh * ow * cout, 0);
  // reference (double, bf16 inputs, fp32 weights)
  const int pt = (kh - 1) / 2, pl = (kw - 1) / 2;
  for (int b = 0; b < n; ++b) for (int y = 0; y < oh; ++y) for (int xx = 0; xx < ow; ++xx)
    for (int co = 0; co < cout; ++co) {
      double acc = 0;
      for (int a = 0; a < kh; ++a) for (int c = 0; c < kw; ++c) {
        int iy = y + a - pt, ix = xx + c - pl;
        if (iy < 0 || iy >= oh || ix < 0 || ix >= ow) continue;
        if (ups) { iy >>= 1; ix >>= 1; }
        for (int ci = 0; ci < cin; ++ci)
          acc += (double)bf2f(x[((((size_t)b * cg + ci / 8) * h + iy) * w + ix) * 8 + (ci & 7)]) *
                 wt[(((size_t)a * kw + c) * cin + ci) * cout + co];
      }
      ref[(((size_t)b * oh + y) * ow + xx) * cout + co] = acc;
    }
  // emulated kernel
  std::vector<uint8_t> stage(p.a_stage_bytes + 4096, 0xFF);
  for (int tile = 0; tile < p.num_tiles; ++tile) {
    int n_tile = tile % p.n_tiles_n, t = tile / p.n_tiles_n;
    int tx = t % p.tiles_x; t /= p.tiles_x; int ty = t % p.tiles_y; int img = t / p.tiles_y;
    const int MT = p.mt_x * p.mt_y;
    std::vector<double> D((size_t)MT * 128 * p.n_cols, 0.0);
    for (int ch = 0; ch < p.cin_chunks; ++ch) {
      // TMA box: dims (W*8, H, CG, N), start ((tx*8-pad_x)*8, ty*16-pad_y, ch*CGC, img)
      uint16_t *s16 = reinterpret_cast<uint16_t *>(stage.data());
      for (int pc = 0; pc < p.planes_per_chunk; ++pc) for (int r = 0; r < p.box_h; ++r) for (int e = 0; e < p.box_w * 8; ++e) {
        int gx = (tx * p.mt_x * kTcTileW - p.pad_x) * 8 + e, gy = ty * p.mt_y * kTcTileH * RM - p.pad_y + r, gp = ch * p.planes_per_chunk + pc;
        uint16_t v = 0;
        if (gx >= 0 && gx < w * 8 && gy >= 0 && gy < h) v = x[(((size_t)img * cg + gp) * h + gy) * w * 8 + gx];
        s16[((size_t)pc * p.box_h + r) * p.box_w * 8 + e] = v;
      }
      for (int ks = 0; ks < p.ksteps; ++ks) {
        const uint16_t *bbase = wp.data() + ((size_t)(n_tile * p.cin_chunks + ch) * p.ksteps + ks) * 2 * p.n_cols * 8;
        for (int t = 0; t < MT; ++t) for (int m = 0; m < 128; ++m) for (int k = 0; k < 16; ++k) {
          const int iy = t / p.mt_x, ix = t % p.mt_x;
          size_t aoff = p.a_off[ks] + (size_t)iy * kTcTileH * RM * p.box_w * 16 + (size_t)ix * 128 +
                        (size_t)(k / 8) * p.a_lbo[ks] + (size_t)(m / 8) * RM * p.box_w * 16 + (m % 8) * 16 + (k % 8) * 2;
          if (p.a_desc_lo[ks] != ((p.a_off[ks] >> 4) | ((p.a_lbo[ks] >> 4) << 16))) { printf("a_desc_lo mismatch\n"); return 1; }
          if (aoff + 2 > p.a_stage_bytes) { printf("A read out of stage: ks %d m %d k %d off %zu\n", ks, m, k, aoff); return 1; }
          float av = bf2f(*reinterpret_cast<uint16_t *>(stage.data() + aoff));
          for (int nn = 0; nn < p.n_cols; ++nn) {
            size_t boff = (size_t)(k / 8) * p.n_cols * 16 + (size_t)(nn / 8) * 128 + (nn % 8) * 16 + (k % 8) * 2;
            D[((size_t)t * 128 + m) * p.n_cols + nn] += (double)av * bf2f(bbase[boff / 2]);
          }
        }
      }
    }
    for (int t = 0; t < MT; ++t) for (int m = 0; m < 128; ++m) {
      const int iy = t / p.mt_x, ix = t % p.mt_x;
      int r = m >> 3, px = m & 7, y = (ty * p.mt_y + iy) * kTcTileH + r, xx = (tx * p.mt_x + ix) * kTcTileW + px;
      if (y * RM >= h || xx >= w) continue;
      for (int j = 0; j < p.n_cols; ++j) {
        int col = n_tile * p.n_cols + j;
        if (col >= p.cols_valid) break;
        int co, oy, ox;
        if (RM == 2) { int par = col / p.scale_mod; co = col - par * p.scale_mod; oy = 2 * y + par; ox = xx; if (oy >= h) continue; }
        else if (p.mode == 0) { co = col; oy = y; ox = xx; }
        else { int par = col / p.cout; co = col - par * p.cout; oy = 2 * y + (par >> 1); ox = 2 * xx + (par & 1); }
        out[(((size_t)img * oh + oy) * ow + ox) * cout + co] = D[((size_t)t * 128 + m) * p.n_cols + j];
      }
    }
  }
  double maxerr = 0, maxref = 0;
  for (size_t i = 0; i < ref.size(); ++i) { maxerr = std::max(maxerr, std::fabs(out[i] - ref[i])); maxref = std::max(maxref, std::fabs(ref[i])); }
  printf("k%dx%d cin %d cout %d ups %d %dx%dx%d: mt %dx%d res %d ksteps %d bgroup %d n_cols %d n_tiles %d chunks %d a_st %d b_st %d smem %zu | max err %.4g (ref max %.3g) %s\n",
         kh, kw, cin, cout, ups, n, h, w, p.mt_x, p.mt_y, p.b_resident, p.ksteps, p.bgroup, p.n_cols, p.n_tiles_n, p.cin_chunks, p.a_stages, p.b_stages, smem,
         maxerr, maxref, maxerr < 0.02 * maxref ? "OK" : "MISMATCH");
  return maxerr < 0.02 * maxref ? 0 : 1;
}

// Tensor-core stem: the uint8 image widened to bf16 is read as [n][1][h][w/8][8] (8 adjacent pixels = the 8 input
// channels of a GEMM row), banded weights from tc_stem_group_weights, output columns [plane][pixel][8 channels]
// = the blocked activation layout.  Checked against a direct 3x3 / 1-channel convolution.
static int run_stem(int cout, int n, int h, int w) {
  TcGeometry g;
  if (tc_make_geometry(3, 3, 8, 8 * cout, 0, &g)) { printf("geometry failed\n"); return 1; }
  g.stem_groups = 1;
  const int wg = w / 8;
  TcConvParams p; size_t smem;
  if (tc_fill_params(g, n, h, wg, &p, &smem)) return 1;
  std::mt19937 rng(7);
  std::uniform_real_distribution<float> U(-1, 1);
  std::vector<float> wt((size_t)9 * cout);
  for (auto &v : wt) v = U(rng);
  std::vector<uint16_t> x((size_t)n * h * w);            // pixel values 0..255, exact in bf16
  for (auto &v : x) v = f2bf((float)(rng() % 256));
  std::vector<float> banded;
  tc_stem_group_weights(wt.data(), cout, &banded);
  std::vector<uint16_t> wp;
  tc_pack_weights(g, banded.data(), &wp);
  std::vector<double> out((size_t)n * h * w * cout, 1e30), ref((size_t)n * h * w * cout, 0);
  for (int b = 0; b < n; ++b) for (int y = 0; y < h; ++y) for (int xx = 0; xx < w; ++xx) for (int co = 0; co < cout; ++co) {
    double acc = 0;
    for (int a = 0; a < 3; ++a) for (int c = 0; c < 3; ++c) {
      const int iy = y + a - 1, ix = xx + c - 1;
      if (iy < 0 || iy >= h || ix < 0 || ix >= w) continue;
      acc += (double)bf2f(x[((size_t)b * h + iy) * w + ix]) * bf2f(f2bf(wt[(a * 3 + c) * cout + co]));
    }
    ref[(((size_t)b * h + y) * w + xx) * cout + co] = acc;
  }
  std::vector<uint8_t> stage(p.a_stage_bytes + 4096, 0xFF);
  for (int tile = 0; tile < p.num_tiles; ++tile) {
    int n_tile = tile % p.n_tiles_n, t = tile / p.n_tiles_n;
    int tx = t % p.tiles_x; t /= p.tiles_x; int ty = t % p.tiles_y; int img = t / p.tiles_y;
    const int MT = p.mt_x * p.mt_y;
    std::vector<double> D((size_t)MT * 128 * p.n_cols, 0.0);
    uint16_t *s16 = reinterpret_cast<uint16_t *>(stage.data());
    for (int r = 0; r < p.box_h; ++r) for (int e = 0; e < p.box_w * 8; ++e) {
      const int gx = (tx * p.mt_x * kTcTileW - p.pad_x) * 8 + e, gy = ty * p.mt_y * kTcTileH - p.pad_y + r;
      uint16_t v = 0;
      if (gx >= 0 && gx < wg * 8 && gy >= 0 && gy < h) v = x[((size_t)img * h + gy) * w + gx];
      s16[(size_t)r * p.box_w * 8 + e] = v;
    }
    for (int ks = 0; ks < p.ksteps; ++ks) {
      const uint16_t *bbase = wp.data() + ((size_t)n_tile * p.ksteps + ks) * 2 * p.n_cols * 8;
      for (int tt = 0; tt < MT; ++tt) for (int m = 0; m < 128; ++m) for (int k = 0; k < 16; ++k) {
        const int iy = tt / p.mt_x, ix = tt % p.mt_x;
        size_t aoff = p.a_off[ks] + (size_t)iy * kTcTileH * p.box_w * 16 + (size_t)ix * 128 + (size_t)(k / 8) * p.a_lbo[ks] +
                      (size_t)(m / 8) * p.box_w * 16 + (m % 8) * 16 + (k % 8) * 2;
        if (aoff + 2 > p.a_stage_bytes) { printf("stem A read out of stage\n"); return 1; }
        const float av = bf2f(*reinterpret_cast<uint16_t *>(stage.data() + aoff));
        for (int nn = 0; nn < p.n_cols; ++nn) {
          size_t boff = (size_t)(k / 8) * p.n_cols * 16 + (size_t)(nn / 8) * 128 + (nn % 8) * 16 + (k % 8) * 2;
          D[((size_t)tt * 128 + m) * p.n_cols + nn] += (double)av * bf2f(bbase[boff / 2]);
        }
      }
    }
    for (int tt = 0; tt < MT; ++tt) for (int m = 0; m < 128; ++m) {
      const int iy = tt / p.mt_x, ix = tt % p.mt_x;
      const int r = m >> 3, px = m & 7, y = (ty * p.mt_y + iy) * kTcTileH + r, gx = (tx * p.mt_x + ix) * kTcTileW + px;
      if (y >= h || gx >= wg) continue;
      for (int j = 0; j < p.n_cols; ++j) {
        const int col = n_tile * p.n_cols + j;
        if (col >= p.cols_valid) break;
        const int plane = col >> 6, pix = (col >> 3) & 7, c8 = col & 7;      // epilogue mode 3 mapping
        out[(((size_t)img * h + y) * w + gx * 8 + pix) * cout + plane * 8 + c8] = D[((size_t)tt * 128 + m) * p.n_cols + j];
      }
    }
  }
  double maxerr = 0, maxref = 0;
  for (size_t i = 0; i < ref.size(); ++i) { maxerr = std::max(maxerr, std::fabs(out[i] - ref[i])); maxref = std::max(maxref, std::fabs(ref[i])); }
  printf("stem groups cout %d %dx%dx%d: mt %dx%d ksteps %d n_cols %d n_tiles %d smem %zu | max err %.4g (ref max %.3g) %s\n", cout, n, h, w,
         p.mt_x, p.mt_y, p.ksteps, p.n_cols, p.n_tiles_n, smem, maxerr, maxref, maxerr < 1e-6 * maxref + 1e-9 ? "OK" : "MISMATCH");
  return maxerr < 1e-6 * maxref + 1e-9 ? 0 : 1;
}

// fp32-accurate split mode: fp32 activations stored as fp16 (hi, lo') plane pairs, split-packed weights, epilogue
// value = (main + corr * 2^-11) / wscale.  Checked against a double-precision convolution of the fp32 data.
#include <cuda_fp16.h>
static float h2f(uint16_t b) { __half_raw r; r.x = b; return __half2float(__half(r)); }
static uint16_t f2h16(float f) { return static_cast<__half_raw>(__float2half_rn(f)).x; }
static int run_split(int kh, int kw, int cin, int cout, int ups, int n, int h, int w, int rowpair = 0) {
  TcGeometry g;
  if (rowpair) {
    if (tc_make_geometry_split(4, 3, cin, 2 * cout, 0, &g, 1, 1)) { printf("split row-pair geometry failed\n"); return 1; }
    g.rows2 = 1;
  } else
  if (tc_make_geometry_split(kh, kw, cin, cout, ups, &g)) { printf("split geometry failed\n"); return 1; }
  TcConvParams p; size_t smem;
  std::mt19937 rng(3);
  std::uniform_real_distribution<float> U(-1, 1);
  std::vector<float> wt((size_t)kh * kw * cin * cout);
  for (auto &v : wt) v = 0.05f * U(rng);
  g.wscale = tc_split_weight_scale(wt.data(), wt.size());
  if (tc_fill_params(g, n, h, w, &p, &smem)) return 1;
  const int cg = cin / 8, cgp = 2 * cg;
  std::vector<float> xf((size_t)n * cin * h * w);            // logical blocked [n][cg][h][w][8]
  for (auto &v : xf) v = std::max(0.f, 3.f * U(rng) + 0.5f);
  std::vector<uint16_t> x((size_t)n * cgp * h * w * 8);      // physical: plane 2p = hi, 2p+1 = lo'
  for (int b = 0; b < n; ++b) for (int pl = 0; pl < cg; ++pl) for (size_t i = 0; i < (size_t)h * w * 8; ++i) {
    const float a = xf[((size_t)b * cg + pl) * h * w * 8 + i];
    const uint16_t hi = f2h16(a);
    x[((size_t)b * cgp + 2 * pl) * h * w * 8 + i] = hi;
    x[((size_t)b * cgp + 2 * pl + 1) * h * w * 8 + i] = f2h16((a - h2f(hi)) * 2048.f);
  }
  std::vector<uint16_t> wp;
  if (rowpair) {
    std::vector<float> banded;
    tc_rowpair_weights(wt.data(), cin, cout, &banded);
    tc_pack_weights(g, banded.data(), &wp, 1);
  } else
  tc_pack_weights(g, wt.data(), &wp, 1);
  const int RM = p.row_mul;
  const int oh = ups ? 2 * h : h, ow = ups ? 2 * w : w;
  std::vector<double> out((size_t)n * oh * ow * cout, 1e30), ref((size_t)n * oh * ow * cout, 0), mag((size_t)n * oh * ow * cout, 0);
  const int pt = (kh - 1) / 2, pl_ = (kw - 1) / 2;
  for (int b = 0; b < n; ++b) for (int y = 0; y < oh; ++y) for (int xx = 0; xx < ow; ++xx)
    for (int co = 0; co < cout; ++co) {
      double acc = 0, m = 0;
      for (int a = 0; a < kh; ++a) for (int c = 0; c < kw; ++c) {
        int iy = y + a - pt, ix = xx + c - pl_;
        if (iy < 0 || iy >= oh || ix < 0 || ix >= ow) continue;
        if (ups) { iy >>= 1; ix >>= 1; }
        for (int ci = 0; ci < cin; ++ci) {
          const double t = (double)xf[((((size_t)b * cg + ci / 8) * h + iy) * w + ix) * 8 + (ci & 7)] * wt[(((size_t)a * kw + c) * cin + ci) * cout + co];
          acc += t; m += std::fabs(t);
        }
      }
      ref[(((size_t)b * oh + y) * ow + xx) * cout + co] = acc;
      mag[(((size_t)b * oh + y) * ow + xx) * cout + co] = m;
    }
  std::vector<uint8_t> stage(p.a_stage_bytes + 4096, 0xFF);
  for (int tile = 0; tile < p.num_tiles; ++tile) {
    int n_tile = tile % p.n_tiles_n, t = tile / p.n_tiles_n;
    int tx = t % p.tiles_x; t /= p.tiles_x; int ty = t % p.tiles_y; int img = t / p.tiles_y;
    const int MT = p.mt_x * p.mt_y;
    std::vector<double> D((size_t)MT * 128 * p.n_cols, 0.0);
    for (int ch = 0; ch < p.cin_chunks; ++ch) {
      uint16_t *s16 = reinterpret_cast<uint16_t *>(stage.data());
      for (int pc = 0; pc < p.planes_per_chunk; ++pc) for (int r = 0; r < p.box_h; ++r) for (int e = 0; e < p.box_w * 8; ++e) {
        int gx = (tx * p.mt_x * kTcTileW - p.pad_x) * 8 + e, gy = ty * p.mt_y * kTcTileH * RM - p.pad_y + r, gp = ch * p.planes_per_chunk + pc;
        uint16_t v = 0;
        if (gx >= 0 && gx < w * 8 && gy >= 0 && gy < h) v = x[(((size_t)img * cgp + gp) * h + gy) * w * 8 + gx];
        s16[((size_t)pc * p.box_h + r) * p.box_w * 8 + e] = v;
      }
      for (int ks = 0; ks < p.ksteps; ++ks) {
        const uint16_t *bbase = wp.data() + ((size_t)(n_tile * p.cin_chunks + ch) * p.ksteps + ks) * 2 * p.n_cols * 8;
        for (int tt = 0; tt < MT; ++tt) for (int m = 0; m < 128; ++m) for (int k = 0; k < 16; ++k) {
          const int iy = tt / p.mt_x, ix = tt % p.mt_x;
          size_t aoff = p.a_off[ks] + (size_t)iy * kTcTileH * RM * p.box_w * 16 + (size_t)ix * 128 +
                        (size_t)(k / 8) * p.a_lbo[ks] + (size_t)(m / 8) * RM * p.box_w * 16 + (m % 8) * 16 + (k % 8) * 2;
          if (aoff + 2 > p.a_stage_bytes) { printf("split A read out of stage\n"); return 1; }
          const float av = h2f(*reinterpret_cast<uint16_t *>(stage.data() + aoff));
          if (av == 0.f) continue;
          for (int nn = 0; nn < p.n_cols; ++nn) {
            size_t boff = (size_t)(k / 8) * p.n_cols * 16 + (size_t)(nn / 8) * 128 + (nn % 8) * 16 + (k % 8) * 2;
            D[((size_t)tt * 128 + m) * p.n_cols + nn] += (double)av * h2f(bbase[boff / 2]);
          }
        }
      }
    }
    for (int tt = 0; tt < MT; ++tt) for (int m = 0; m < 128; ++m) {
      const int iy = tt / p.mt_x, ix = tt % p.mt_x;
      int r = m >> 3, px = m & 7, y = (ty * p.mt_y + iy) * kTcTileH + r, xx = (tx * p.mt_x + ix) * kTcTileW + px;
      if (y * RM >= h || xx >= w) continue;
      const int groups = std::min(p.n_cols, p.cols_valid - n_tile * p.n_cols) >> 4;
      for (int j = 0; j < groups; ++j) {
        const int colp = n_tile * p.n_cols + j * 16;       // the kernel's EPI 7 mapping
        int co0 = (colp >> 4) << 3, oy = y, ox = xx;
        if (p.mode == 1) { const int par = colp / p.cout; co0 = ((colp - par * p.cout) >> 4) << 3; oy = 2 * y + (par >> 1); ox = 2 * xx + (par & 1); }
        if (RM == 2) { const int cg8 = p.scale_mod >> 3, par = j / cg8; co0 = (j - par * cg8) * 8; oy = 2 * y + par; if (oy >= h) continue; }
        for (int k = 0; k < 8; ++k) {
          const float mainv = (float)D[((size_t)tt * 128 + m) * p.n_cols + j * 16 + k], corr = (float)D[((size_t)tt * 128 + m) * p.n_cols + j * 16 + 8 + k];
          const float val = std::fmaf(corr, 4.8828125e-4f, mainv) * p.scale_mul;
          out[(((size_t)img * oh + oy) * ow + ox) * cout + co0 + k] = val;
        }
      }
    }
  }
  double maxrel = 0;
  for (size_t i = 0; i < ref.size(); ++i) maxrel = std::max(maxrel, std::fabs(out[i] - ref[i]) / std::max(mag[i], 1e-30));
  const bool ok = maxrel < 2e-6 && p.scale_mod == cout && p.row_mul == (rowpair ? 2 : 1);
  printf("split k%dx%d cin %d cout %d ups %d %dx%dx%d: mt %dx%d res %d ksteps %d n_cols %d n_tiles %d chunks %d a_st %d smem %zu wscale %g | max err / sum|terms| %.3g %s\n",
         kh, kw, cin, cout, ups, n, h, w, p.mt_x, p.mt_y, p.b_resident, p.ksteps, p.n_cols, p.n_tiles_n, p.cin_chunks, p.a_stages, smem,
         g.wscale, maxrel, ok ? "OK" : "MISMATCH");
  return ok ? 0 : 1;
}

// Data gradient of the up-conv on the low-res grid (tc_make_geometry_s2d): the stage holds the four pixel parities of the
// high-res dz as sub-tiles [parity][plane][row][px] (what the four stride-2 tensor maps deliver), every K half is a tap-shifted
// read of one sub-tile.  Checked against the adjoint of z = conv2x2(upsample2x(a)) written out directly.
static int run_s2d(int cz, int cout, int n, int h, int w) {      // (h, w) = low-res grid
  TcGeometry g;
  if (tc_make_geometry_s2d(cz, cout, &g)) { printf("s2d geometry failed\n"); return 1; }
  TcConvParams p; size_t smem;
  if (tc_fill_params(g, n, h, w, &p, &smem)) return 1;
  const int H = 2 * h, W = 2 * w, P = cz / 8;
  std::mt19937 rng(3);
  std::uniform_real_distribution<float> U(-1, 1);
  std::vector<float> wt((size_t)4 * cout * cz);                 // forward kernel [2][2][cin_up = cout][cout_up = cz]
  for (auto &v : wt) v = U(rng);
  std::vector<uint16_t> dz((size_t)n * cz * H * W);              // blocked [n][P][H][W][8]
  for (auto &v : dz) v = f2bf(U(rng));
  std::vector<uint16_t> wp;
  tc_pack_weights(g, wt.data(), &wp);
  std::vector<double> out((size_t)n * h * w * cout, 1e30), ref((size_t)n * h * w * cout, 0);
  // reference: d(a)[Y][X][ci] = sum over (y, x, ky, kx) with (y + ky) >> 1 == Y, (x + kx) >> 1 == X of W[ky][kx][ci][co] dz[y][x][co]
  for (int b = 0; b < n; ++b) for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) for (int ky = 0; ky < 2; ++ky) for (int kx = 0; kx < 2; ++kx) {
    const int Y = (y + ky) >> 1, X = (x + kx) >> 1;
    if (Y >= h || X >= w) continue;                             // the forward read zero padding there
    for (int co = 0; co < cz; ++co) {
      const double d = bf2f(dz[((((size_t)b * P + co / 8) * H + y) * W + x) * 8 + (co & 7)]);
      for (int ci = 0; ci < cout; ++ci)
        ref[(((size_t)b * h + Y) * w + X) * cout + ci] += d * bf2f(f2bf(wt[(((size_t)ky * 2 + kx) * cout + ci) * cz + co]));
    }
  }
  std::vector<uint8_t> stage(p.a_stage_bytes + 4096, 0xFF);
  if (p.a_tx_bytes != (uint32_t)(4 * P * p.box_w * p.box_h * 16)) { printf("s2d: a_tx_bytes\n"); return 1; }
  for (int tile = 0; tile < p.num_tiles; ++tile) {
    int n_tile = tile % p.n_tiles_n, t = tile / p.n_tiles_n;
    int tx = t % p.tiles_x; t /= p.tiles_x; int ty = t % p.tiles_y; int img = t / p.tiles_y;
    const int MT = p.mt_x * p.mt_y;
    std::vector<double> D((size_t)MT * 128 * p.n_cols, 0.0);
    for (int par = 0; par < 4; ++par) {                         // the four TMA loads: parity map = base + (py * W + px) pixels
      const int py = par >> 1, px = par & 1;
      uint16_t *s16 = reinterpret_cast<uint16_t *>(stage.data() + (size_t)par * p.s2d_part_bytes);
      for (int pc = 0; pc < P; ++pc) for (int r = 0; r < p.box_h; ++r) for (int c = 0; c < p.box_w; ++c) for (int e = 0; e < 8; ++e) {
        const int X = tx * p.mt_x * kTcTileW - p.pad_x + c, Y = ty * p.mt_y * kTcTileH - p.pad_y + r;
        uint16_t v = 0;
        if (X >= 0 && X < w && Y >= 0 && Y < h) v = dz[((((size_t)img * P + pc) * H + 2 * Y + py) * W + 2 * X + px) * 8 + e];
        s16[(((size_t)pc * p.box_h + r) * p.box_w + c) * 8 + e] = v;
      }
    }
    for (int ks = 0; ks < p.ksteps; ++ks) {
      const uint16_t *bbase = wp.data() + ((size_t)n_tile * p.ksteps + ks) * 2 * p.n_cols * 8;
      for (int t2 = 0; t2 < MT; ++t2) for (int m = 0; m < 128; ++m) for (int k = 0; k < 16; ++k) {
        const int iy = t2 / p.mt_x, ix = t2 % p.mt_x;
        size_t aoff = p.a_off[ks] + (size_t)iy * kTcTileH * p.box_w * 16 + (size_t)ix * 128 + (size_t)(k / 8) * p.a_lbo[ks] +
                      (size_t)(m / 8) * p.box_w * 16 + (m % 8) * 16 + (k % 8) * 2;
        if (aoff + 2 > p.a_stage_bytes) { printf("s2d A read out of stage\n"); return 1; }
        const float av = bf2f(*reinterpret_cast<uint16_t *>(stage.data() + aoff));
        for (int nn = 0; nn < p.n_cols; ++nn) {
          size_t boff = (size_t)(k / 8) * p.n_cols * 16 + (size_t)(nn / 8) * 128 + (nn % 8) * 16 + (k % 8) * 2;
          D[((size_t)t2 * 128 + m) * p.n_cols + nn] += (double)av * bf2f(bbase[boff / 2]);
        }
      }
    }
    for (int t2 = 0; t2 < MT; ++t2) for (int m = 0; m < 128; ++m) {
      const int iy = t2 / p.mt_x, ix = t2 % p.mt_x;
      const int y = (ty * p.mt_y + iy) * kTcTileH + (m >> 3), xx = (tx * p.mt_x + ix) * kTcTileW + (m & 7);
      if (y >= h || xx >= w) continue;
      for (int j = 0; j < p.n_cols; ++j) {
        const int col = n_tile * p.n_cols + j;
        if (col >= p.cols_valid) break;
        out[(((size_t)img * h + y) * w + xx) * cout + col] = D[((size_t)t2 * 128 + m) * p.n_cols + j];
      }
    }
  }
  double maxerr = 0, maxref = 0;
  for (size_t i = 0; i < ref.size(); ++i) { maxerr = std::max(maxerr, std::fabs(out[i] - ref[i])); maxref = std::max(maxref, std::fabs(ref[i])); }
  printf("s2d cz %d cout %d %dx%dx%d: mt %dx%d ksteps %d n_cols %d a_st %d part %u | max err %.4g (ref max %.3g) %s\n", cz, cout, n, h, w,
         p.mt_x, p.mt_y, p.ksteps, p.n_cols, p.a_stages, p.s2d_part_bytes, maxerr, maxref, maxerr < 0.02 * maxref ? "OK" : "MISMATCH");
  return maxerr < 0.02 * maxref ? 0 : 1;
}

int main() {
  int bad = 0;
  bad += run_s2d(8, 16, 2, 32, 24);       // up3 of the default net: 9 single-plane halves (one dummy)
  bad += run_s2d(8, 16, 3, 64, 64);       // super-tiles
  bad += run_s2d(16, 32, 1, 20, 12);      // plane pairs, ragged tiles
  bad += run_s2d(32, 64, 1, 16, 16);
  bad += run_split(3, 3, 8, 8, 0, 2, 32, 24);
  bad += run_split(3, 3, 16, 8, 0, 1, 20, 12);
  bad += run_split(3, 3, 8, 16, 0, 40, 64, 64);
  bad += run_split(2, 2, 16, 8, 1, 2, 32, 24);
  bad += run_split(3, 3, 64, 64, 0, 1, 16, 16);
  bad += run_split(3, 3, 128, 128, 0, 1, 16, 8);
  bad += run_split(2, 2, 128, 64, 1, 1, 16, 8);
  bad += run_split(3, 3, 256, 256, 0, 1, 16, 8);   // physical 512 columns: two n-tiles
  bad += run_split(3, 3, 8, 8, 0, 2, 32, 24, 1);   // split mode x row pairs
  bad += run_split(3, 3, 16, 8, 0, 1, 64, 40, 1);
  bad += run_split(3, 3, 8, 16, 0, 1, 34, 16, 1);  // ragged: 34 rows in 32-row tiles (32 -> 16 would need 48 K steps: no pair plan)
  bad += run_split(3, 3, 16, 16, 0, 40, 64, 64, 1);
  bad += run_stem(8, 2, 32, 128);
  bad += run_stem(8, 1, 40, 136);     // ragged group count
  bad += run_stem(16, 1, 16, 64);
  bad += run_stem(32, 1, 16, 128);    // 256 columns
  bad += run(3, 3, 8, 8, 0, 2, 32, 24);
  bad += run(3, 3, 8, 16, 0, 1, 16, 16);
  bad += run(3, 3, 16, 16, 0, 1, 32, 16);
  bad += run(3, 3, 32, 64, 0, 1, 16, 8);
  bad += run(3, 3, 128, 128, 0, 1, 16, 8);
  bad += run(3, 3, 128, 64, 0, 1, 16, 16);
  bad += run(2, 2, 128, 64, 1, 1, 16, 8);
  bad += run(2, 2, 16, 8, 1, 2, 32, 24);
  bad += run(3, 3, 16, 8, 0, 1, 20, 12);   // ragged tiles
  bad += run(3, 3, 8, 8, 0, 40, 64, 64);    // super-tiles 8x2
  bad += run(3, 3, 16, 8, 0, 80, 32, 64);   // super-tiles
  bad += run(2, 2, 16, 8, 1, 80, 32, 32);   // up-conv super-tiles
  bad += run(3, 3, 64, 64, 0, 40, 32, 32);
  bad += run(3, 3, 512, 512, 0, 1, 16, 8); // wide net: n-tiles + chunks
  bad += run(2, 2, 256, 128, 1, 1, 16, 8); // wide up-conv: 512 columns
  // row pairs (inference + training plans of the 8 / 16-channel 3x3 layers)
  bad += run(3, 3, 8, 8, 0, 2, 32, 24, 1);
  bad += run(3, 3, 16, 8, 0, 1, 64, 40, 1);
  bad += run(3, 3, 32, 16, 0, 1, 34, 16, 1);   // ragged: 34 rows in 32-row tiles
  bad += run(3, 3, 16, 16, 0, 40, 64, 64, 1);  // super-tiles
  bad += run(3, 3, 8, 16, 0, 3, 96, 72, 1);
  printf(bad ? "FAILED %d\n" : "ALL OK\n", bad);
  return bad;
}
