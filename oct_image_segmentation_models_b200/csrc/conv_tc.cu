// tcgen05 implicit-GEMM convolution for sm_100a.
//
// GEMM view of one conv block (reference models/unet.py:26-29, 41-44):
//   D[pixel, cout] = sum_{tap, cin} X[pixel + tap, cin] * W[tap, cin, cout]
//   M = 128 output pixels (an 8 px x 16 row tile), N = cout (padded to 16), K = taps*cin.
//
// B200-first design (this is not how cuDNN/CUTLASS stage a convolution):
//   * activations live in HBM as [N][C/8][H][W][8]; ONE TMA box load brings the
//     (8+kw-1) x (16+kh-1) halo tile of every 8-channel plane of the chunk into smem
//     as [plane][row][px][8ch] -- TMA zero-fills out-of-image coordinates, which IS
//     Keras "same" padding;
//   * that layout is already the canonical no-swizzle K-major UMMA operand: 8 pixels x
//     16 B form a core matrix, SBO = tile row pitch, LBO = plane pitch.  Every filter
//     tap is the same smem tile viewed through a descriptor whose start address is
//     shifted by (dy*pitch + dx*16) bytes, so the halo tile is read from L2/HBM once
//     and never duplicated in smem (no im2col);
//   * for Cin = 8 one K=16 MMA covers two taps (LBO = distance between the taps);
//   * the x2 nearest up-sampling in front of the decoder 2x2 conv is folded into the
//     weights: the four output parities are four column groups of one GEMM on the
//     LOW-res tile and the epilogue scatters them (pixel shuffle) straight into the
//     up-conv's plane range of the concat buffer;
//   * accumulators sit in TMEM (double buffered), the epilogue applies the folded
//     BatchNorm scale/shift + ReLU and stores one 16 B vector per pixel per plane;
//   * the data gradient of that up-conv is a stride-2 3x3 conv over dz: the producer fetches the
//     four pixel parities of dz as four sub-tiles through stride-2 5-D tensor maps, after which every
//     tap is again a shifted descriptor on one of the sub-tiles (tc_make_geometry_s2d).
//
// Warp roles (512 threads, 1 CTA/SM, persistent over super-tiles):
//   warp 0        : TMEM allocator; lane 0 = TMA producer (A halo tiles + packed weights)
//   warps 1..3    : lane 0 of each = MMA issuer for a third of the super-tile's M-tiles
//   warps 4..15   : epilogue, three warpgroups (TMEM lane quarter = warp % 4)
#include "conv_tc.cuh"
#include "tc_ptx.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include <cuda_fp16.h>

namespace octseg {

// 16-bit storage helpers (bf16 or fp16 selected per launch; warp-uniform branch)
__device__ __forceinline__ uint32_t pack2(float a, float b, int fp16) {
  if (fp16) { __half2 h = __floats2half2_rn(a, b); return *reinterpret_cast<uint32_t *>(&h); }
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t *>(&h);
}
__device__ __forceinline__ uint32_t max2(uint32_t a, uint32_t b, int fp16) {
  if (fp16) { __half2 r = __hmax2(*reinterpret_cast<__half2 *>(&a), *reinterpret_cast<__half2 *>(&b)); return *reinterpret_cast<uint32_t *>(&r); }
  __nv_bfloat162 r = __hmax2(*reinterpret_cast<__nv_bfloat162 *>(&a), *reinterpret_cast<__nv_bfloat162 *>(&b));
  return *reinterpret_cast<uint32_t *>(&r);
}

// Folded BN + ReLU + 16-bit pack of two accumulator values: one packed f32x2 FMA, one convert, one packed
// 16-bit max against `floor2` (0 | 0 with ReLU, the most negative finite pair without).  ReLU after the
// rounding equals ReLU before it (rounding is monotonic and keeps 0).
__device__ __forceinline__ uint32_t bn_relu_pack2(uint32_t v0, uint32_t v1, float s0, float s1, float h0, float h1,
                                                 uint32_t floor2, int fp16) {
  const float2 y = __ffma2_rn(make_float2(__uint_as_float(v0), __uint_as_float(v1)), make_float2(s0, s1), make_float2(h0, h1));
  return max2(pack2(y.x, y.y, fp16), floor2, fp16);
}

// K-major, no-swizzle shared-memory matrix descriptor (sm_100 "version 1").
//   core matrix = 8 rows x 16 B, rows 16 B apart; SBO = bytes between 8-row groups,
//   LBO = bytes between the two 8-element K halves of one K=16 step.
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}

constexpr int kMaxStages = 12;

struct __align__(8) TcBarriers {
  uint64_t a_full[kMaxStages], a_empty[kMaxStages];
  uint64_t b_full[kMaxStages], b_empty[kMaxStages];
  uint64_t acc_full[2], acc_empty[2];
  uint64_t w_full;          // resident weights landed
  uint32_t tmem_base;
  uint32_t pad;
};

// Tile coordinates of a persistent CTA, advanced by gridDim.x tiles per step with carries instead of
// the four integer divisions per tile (measured: ~10 % of all instructions of the narrow layers).
struct TileIter {
  int n_tile, tx, ty, img;
  int dn, dx, dy, dimg;
  __device__ __forceinline__ void init(const TcConvParams &p, int tile, int step) {
    n_tile = tile % p.n_tiles_n; int t = tile / p.n_tiles_n;
    tx = t % p.tiles_x; t /= p.tiles_x;
    ty = t % p.tiles_y; img = t / p.tiles_y;
    dn = step % p.n_tiles_n; t = step / p.n_tiles_n;
    dx = t % p.tiles_x; t /= p.tiles_x;
    dy = t % p.tiles_y; dimg = t / p.tiles_y;
  }
  __device__ __forceinline__ void advance(const TcConvParams &p) {
    n_tile += dn;
    int c = 0;
    if (n_tile >= p.n_tiles_n) { n_tile -= p.n_tiles_n; c = 1; }
    tx += dx + c; c = 0;
    if (tx >= p.tiles_x) { tx -= p.tiles_x; c = 1; }
    ty += dy + c; c = 0;
    if (ty >= p.tiles_y) { ty -= p.tiles_y; c = 1; }
    img += dimg + c;
  }
};

constexpr int kTcIssuers = 3;    // MMA-issuing threads (warps 1..3)
constexpr int kTcEpiWarps = 12;   // epilogue warps 4..15 (three warpgroups)
constexpr int kTcThreads = (4 + kTcEpiWarps) * 32;

// HK: compile-time class count of the fused 1x1-conv + softmax head (0 = no head fusion)
// EPI: epilogue specialisation chosen on the host (tc_epi_kind) so that every layer type gets its own
// register allocation: 0 = generic chunk loop (wide layers, unpaired pixel shuffle, wide fused head),
// 1 = 8-column layers (batched TMEM loads; plain / pool / fused head), 2 = 16-column plain / pool layers,
// 3 = stem pixel groups, 4 = paired pixel-shuffle up-conv, 9 = generic chunk loop that stores only the 2x2 sums of its outputs
// (data gradient of the wide up-conv).  Dead paths are dropped at compile time.
// ST: 1 = the epilogue also accumulates per-channel sum / sum of squares of its fp32 output (training forward)
template <int HK, int EPI, int ST = 0>
__global__ void __launch_bounds__(kTcThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ TcConvParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // provably warp-uniform
  const int lane = threadIdx.x & 31;
  const uint32_t a_stride = (p.a_stage_bytes + 127u) & ~127u;
  const uint32_t b_stride = (p.b_stage_bytes + 127u) & ~127u;
  uint8_t *a_smem = smem;
  uint8_t *b_smem = smem + (size_t)a_stride * p.a_stages;
  TcBarriers *bars = reinterpret_cast<TcBarriers *>(b_smem + (size_t)b_stride * p.b_stages);
  float *s_scale = reinterpret_cast<float *>(bars + 1);     // [cout] folded BN scale
  float *s_shift = s_scale + p.cout;                        // [cout] folded BN shift
  float *s_head = s_shift + p.cout;                         // [cout][HK] + [HK] fused-head weights
  if (!p.pdl_late) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int mt = p.mt_x * p.mt_y;                 // M-tiles (128 rows each) per super-tile
  const uint32_t acc_cols = (uint32_t)(mt * p.n_cols);

  uint32_t tmem_cols = 32;
  while (tmem_cols < 2u * acc_cols) tmem_cols <<= 1;

  if (threadIdx.x == 0) {
    // every one of the kTcIssuers MMA-issuing threads commits to the "consumed" barriers
    for (int i = 0; i < p.a_stages; ++i) { mbar_init(smem_u32(&bars->a_full[i]), 1); mbar_init(smem_u32(&bars->a_empty[i]), kTcIssuers); }
    for (int i = 0; i < p.b_stages; ++i) { mbar_init(smem_u32(&bars->b_full[i]), 1); mbar_init(smem_u32(&bars->b_empty[i]), kTcIssuers); }
    for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&bars->acc_full[i]), kTcIssuers); mbar_init(smem_u32(&bars->acc_empty[i]), kTcEpiWarps); }
    mbar_init(smem_u32(&bars->w_full), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    if (p.b_resident && p.static_weights && lane == 0) {
      // inference: the packed weights do not depend on the preceding kernels, so the resident weight
      // image is fetched before the dependency wait (lane 0 initialised the barrier above)
      const uint32_t wfull = smem_u32(&bars->w_full);
      mbar_expect_tx(wfull, p.b_stage_bytes);
      uint32_t done = 0;
      while (done < p.b_stage_bytes) {
        const uint32_t part = min(p.b_stage_bytes - done, 32768u);
        bulk_load(smem_u32(b_smem) + done, reinterpret_cast<const uint8_t *>(p.wpack) + done, part, wfull);
        done += part;
      }
    }
  }
  // Everything above touches no global memory written by earlier kernels.  From here on the kernel reads
  // what earlier kernels of the stream wrote (activations, and in training the BN tables and repacked
  // weights): wait for them.  In inference the epilogue tables are static too and are fetched first.
  auto load_tables = [&]() {
    for (int i = threadIdx.x; i < p.cout; i += blockDim.x) {
      const int c = i % p.scale_mod;           // row-pair mode: columns [parity][channel] share the channel tables
      s_scale[i] = EPI == 7 ? p.scale[c] * p.scale_mul : p.scale[c];   // split mode: undo the weight pre-scale (a power of two)
      s_shift[i] = p.shift[c];
    }
    if constexpr (EPI == 8 || ST == 1) {
      for (int i = threadIdx.x; i < 2 * p.scale_mod; i += blockDim.x) s_head[i] = 0.f;   // per-CTA sum / sum of squares
    }
    if constexpr (HK > 0) {
      for (int i = threadIdx.x; i < p.cout * HK; i += blockDim.x) s_head[i] = p.head_w[((i / HK) % p.scale_mod) * HK + (i % HK)];
      for (int i = threadIdx.x; i < HK; i += blockDim.x) s_head[p.cout * HK + i] = p.head_b[i];
    } else {
      (void)s_head;
    }
  };
  if (p.static_weights) load_tables();
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (!p.static_weights) load_tables();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == 0) {
    {
      // ===================== TMA producer (warp-uniform; elected lane issues) =====================
      const bool leader = elect_one_sync() != 0;
      if (leader) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
      if (p.b_resident && !p.static_weights && leader) {
        // whole packed weight image of the layer stays in smem for the kernel's lifetime
        const uint32_t wfull = smem_u32(&bars->w_full);
        mbar_expect_tx(wfull, p.b_stage_bytes);
        uint32_t done = 0;
        while (done < p.b_stage_bytes) {           // bulk copies are capped well below 1 MB each
          const uint32_t part = min(p.b_stage_bytes - done, 32768u);
          bulk_load(smem_u32(b_smem) + done, reinterpret_cast<const uint8_t *>(p.wpack) + done, part, wfull);
          done += part;
        }
      }
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      bool ok = true;
      long long w_prod = 0;
      const long long t_start = clock64();
      TileIter ti;
      ti.init(p, blockIdx.x, gridDim.x);
      for (int tile = blockIdx.x; ok && tile < p.num_tiles; tile += gridDim.x, ti.advance(p)) {
        const int n_tile = ti.n_tile, tx = ti.tx, ty = ti.ty, img = ti.img;
        const uint8_t *wsrc = reinterpret_cast<const uint8_t *>(p.wpack) +
                              (size_t)n_tile * p.cin_chunks * p.ksteps * 32u * p.n_cols;
        for (int ch = 0; ok && ch < p.cin_chunks; ++ch) {
          ok = mbar_wait(smem_u32(&bars->a_empty[as]), aph ^ 1u, p.status, 1, &w_prod);
          if (!ok) break;
          const uint32_t afull = smem_u32(&bars->a_full[as]);
          if (leader) {
            mbar_expect_tx(afull, p.a_tx_bytes);
            if (p.s2d) {
              // four pixel parities of the high-res input, each a [plane][row][px] sub-tile of low-res coordinates
#pragma unroll
              for (int par = 0; par < 4; ++par)
                tma_load_5d(smem_u32(a_smem + (size_t)as * a_stride) + (uint32_t)par * p.s2d_part_bytes, p.s2d_maps + par, afull,
                            0, tx * p.mt_x * kTcTileW - p.pad_x, ty * p.mt_y * kTcTileH - p.pad_y, 0, img);
            } else {
              // tensor map is declared in 8-byte elements: x coordinate = px * 2
              tma_load_4d(smem_u32(a_smem + (size_t)as * a_stride), &tmap_a, afull,
                          (tx * p.mt_x * kTcTileW - p.pad_x) * 2, ty * p.mt_y * kTcTileH * p.row_mul - p.pad_y,
                          ch * p.planes_per_chunk, img);
            }
          }
          if (++as == p.a_stages) { as = 0; aph ^= 1u; }
          if (p.b_resident) continue;
          for (int g = 0; g < p.ksteps; g += p.bgroup) {
            ok = mbar_wait(smem_u32(&bars->b_empty[bs]), bph ^ 1u, p.status, 2);
            if (!ok) break;
            const uint32_t bfull = smem_u32(&bars->b_full[bs]);
            if (leader) {
              mbar_expect_tx(bfull, p.b_stage_bytes);
              bulk_load(smem_u32(b_smem + (size_t)bs * b_stride),
                        wsrc + ((size_t)ch * p.ksteps + g) * 32u * p.n_cols, p.b_stage_bytes, bfull);
            }
            if (++bs == p.b_stages) { bs = 0; bph ^= 1u; }
          }
        }
      }
      // Programmatic dependent launch: this CTA has requested its last tile, so the next kernel of the
      // stream may be scheduled; its barrier init / TMEM allocation / weight fetch then overlap the MMA
      // and epilogue tail of this grid.  (Signalling at kernel entry instead lets the dependent CTAs
      // co-reside for the whole kernel -- two ~113 KB CTAs fit one SM -- and slows this grid down.)
      if (p.pdl_late) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
      if (p.dbg && blockIdx.x == 0 && leader) { p.dbg[0] = w_prod; p.dbg[1] = clock64() - t_start; }
    }
  } else if (warp <= kTcIssuers) {
    {
      // ===================== MMA issuers (warps 1..kTcIssuers; warp-uniform code, elected lane issues) ====
      // One thread sustains only ~1 tcgen05.mma per 108 clk (measured, tools/mma_bench.cu) while
      // the tensor pipe accepts one M128xN16xK16 MMA per ~39 clk, so the M-tiles of a super-tile
      // are dealt round-robin to kTcIssuers threads, each with its own commits.
      const int issuer = warp - 1;
      // instruction descriptor: D = f32, A/B = bf16 (format 1) or fp16 (format 0), K-major, N, M = 128
      const uint32_t ab_fmt = p.fp16 ? 0u : ((1u << 7) | (1u << 10));
      const uint32_t idesc = (1u << 4) | ab_fmt | ((uint32_t)(p.n_cols >> 3) << 17) | ((128u >> 4) << 24);
      const uint32_t pitch = (uint32_t)p.box_w * 16u;
      const uint32_t lbo_b = (uint32_t)p.n_cols * 16u;
      const uint32_t kstep_b = 32u * (uint32_t)p.n_cols;
      const uint32_t a_hi = ((pitch * (uint32_t)p.row_mul) >> 4) | (1u << 14);   // SBO | descriptor version bit 46
      const uint32_t b_hi = (128u >> 4) | (1u << 14);
      // this issuer's M-tiles: descriptor start-address deltas (16 B units) and TMEM column offsets
      constexpr int kMaxMine = (16 + kTcIssuers - 1) / kTcIssuers;
      uint32_t t_desc[kMaxMine], t_col[kMaxMine];
      int n_mine = 0;
#pragma unroll
      for (int j = 0; j < kMaxMine; ++j) {
        const int t = issuer + j * kTcIssuers;
        t_desc[j] = 0; t_col[j] = 0;
        if (t < mt) {
          const int iy = t / p.mt_x, ix = t - iy * p.mt_x;
          t_desc[j] = ((uint32_t)iy * kTcTileH * (uint32_t)p.row_mul * pitch + (uint32_t)ix * 128u) >> 4;
          t_col[j] = (uint32_t)(t * p.n_cols);
          n_mine = j + 1;
        }
      }
      int as = 0, bs = 0, acc = 0;
      uint32_t aph = 0, bph = 0, accph = 0;
      bool ok = true;
      long long w_acc = 0, w_a = 0, w_b = 0;
      const long long t_start = clock64();
      if (p.b_resident) {
        ok = mbar_wait(smem_u32(&bars->w_full), 0, p.status, 7);
        tc_fence_after();
      }
      for (int tile = blockIdx.x; ok && tile < p.num_tiles; tile += gridDim.x) {
        ok = mbar_wait(smem_u32(&bars->acc_empty[acc]), accph ^ 1u, p.status, 3, &w_acc);
        if (!ok) break;
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)acc * acc_cols;
        for (int ch = 0; ok && ch < p.cin_chunks; ++ch) {
          ok = mbar_wait(smem_u32(&bars->a_full[as]), aph, p.status, 4, &w_a);
          if (!ok) break;
          tc_fence_after();
          const uint32_t a_base = smem_u32(a_smem + (size_t)as * a_stride);
          for (int g = 0; g < p.ksteps; g += p.bgroup) {
            uint32_t b_base;
            if (p.b_resident) {
              b_base = smem_u32(b_smem) + (uint32_t)(ch * p.ksteps + g) * kstep_b;
            } else {
              ok = mbar_wait(smem_u32(&bars->b_full[bs]), bph, p.status, 5, &w_b);
              if (!ok) break;
              tc_fence_after();
              b_base = smem_u32(b_smem + (size_t)bs * b_stride);
            }
            // Descriptors are assembled from precomputed 32-bit halves: the low word is a table entry
            // plus the stage base (all operands are 16-byte aligned and below 256 KB, so the 14-bit
            // address field cannot carry into the stride field), the high word is constant.  A single
            // issuing thread is latency-bound on this loop, so it is kept to a handful of instructions.
            uint32_t b_lo = ((b_base >> 4) & 0x3FFFu) | ((lbo_b >> 4) << 16);
            const uint32_t a_base16 = (a_base >> 4) & 0x3FFFu;
#pragma unroll 2
            for (int s = 0; s < p.bgroup; ++s) {
              const int ks = g + s;
              const uint32_t a_lo = p.a_desc_lo[ks] + a_base16;
              const uint64_t db = ((uint64_t)b_hi << 32) | b_lo;
              const uint32_t accum = (ch | ks) != 0 ? 1u : 0u;
#pragma unroll
              for (int j = 0; j < kMaxMine; ++j)
                if (j < n_mine) umma_bf16(d_tmem + t_col[j], ((uint64_t)a_hi << 32) | (a_lo + t_desc[j]), db, idesc, accum);
              b_lo += kstep_b >> 4;
            }
            if (!p.b_resident) {
              umma_commit(smem_u32(&bars->b_empty[bs]));
              if (++bs == p.b_stages) { bs = 0; bph ^= 1u; }
            }
          }
          umma_commit(smem_u32(&bars->a_empty[as]));
          if (++as == p.a_stages) { as = 0; aph ^= 1u; }
        }
        umma_commit(smem_u32(&bars->acc_full[acc]));
        if (++acc == 2) { acc = 0; accph ^= 1u; }
      }
      if (p.dbg && blockIdx.x == 0 && issuer == 0 && lane == 0) {
        p.dbg[2] = w_acc; p.dbg[3] = w_a; p.dbg[4] = w_b; p.dbg[5] = clock64() - t_start;
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (2 warpgroups, M-tiles interleaved) =====================
    // Each thread owns one GEMM row (= pixel) of an M-tile: TMEM lane m, columns = channels.
    // Work is done on 8-column chunks (one 16 B output vector), two chunks per TMEM-load batch.
    const int q = warp & 3;                 // TMEM lane quarter
    const int wg = (warp - 4) >> 2;         // epilogue warpgroup 0..kTcEpiWarps/4-1
    const int m = q * 32 + lane;            // GEMM row == TMEM lane
    const int r = m >> 3, px = m & 7;       // row / px inside the 8x16 M-tile
    const float relu_floor = p.relu ? 0.f : -3.0e38f;
    const uint32_t floor2 = p.relu ? 0u : (p.fp16 ? 0xFBFFFBFFu : 0xFF7FFF7Fu);   // packed 16-bit ReLU floor
    const int nch = (min(p.cols_valid, p.n_cols)) >> 3;      // chunks per n-tile holding valid columns
    const long long plane_elems = (long long)p.out_h * p.out_w * 8;
    int acc = 0;
    uint32_t accph = 0;
    bool ok = true;
    long long w_epi = 0;
    const long long t_start = clock64();
    int wgr = wg;                           // warpgroup -> M-tile assignment, rotated every super-tile
    // EPI 8 (training forward): per-thread partial sums of z and z^2 for up to 16 channels, kept across all tiles
    float st_s[16], st_q[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { st_s[i] = 0.f; st_q[i] = 0.f; }
    TileIter ti;
    ti.init(p, blockIdx.x, gridDim.x);
    for (int tile = blockIdx.x; ok && tile < p.num_tiles; tile += gridDim.x, ti.advance(p)) {
      const int n_tile = ti.n_tile, tx = ti.tx, ty = ti.ty, img = ti.img;
      // the M-tile count of a super-tile is a power of two and there are 3 warpgroups: rotating the
      // assignment spreads the odd tile over the warpgroups across consecutive super-tiles (the two
      // accumulator stages absorb the skew) instead of always loading warpgroup 0
      const int wg_cur = wgr;
      if (++wgr == kTcEpiWarps / 4) wgr = 0;
      ok = mbar_wait(smem_u32(&bars->acc_full[acc]), accph, p.status, 6, &w_epi);
      if (!ok) break;
      tc_fence_after();
      const int col_base = n_tile * p.n_cols;
      const int my_nch = min(nch, (p.cols_valid - col_base) >> 3);
      if (EPI == 1) {
        // ---- 8-channel layers (the full-resolution layers that carry most of the bytes): one
        //      chunk per M-tile, so batch the TMEM loads of up to kEpiBatch M-tiles behind ONE
        //      tcgen05.wait::ld and process them back to back (hides the TMEM/LDS/shuffle latency
        //      that otherwise serialises per M-tile)
        constexpr int kEpiBatch = 3;
        constexpr int kWG = kTcEpiWarps / 4;
        const float4 sa = *reinterpret_cast<const float4 *>(s_scale + col_base), sb = *reinterpret_cast<const float4 *>(s_scale + col_base + 4);
        const float4 ha = *reinterpret_cast<const float4 *>(s_shift + col_base), hb = *reinterpret_cast<const float4 *>(s_shift + col_base + 4);
        const float sc[8] = {sa.x, sa.y, sa.z, sa.w, sb.x, sb.y, sb.z, sb.w};
        const float sh[8] = {ha.x, ha.y, ha.z, ha.w, hb.x, hb.y, hb.z, hb.w};
        for (int t0 = wg_cur; t0 < mt; t0 += kEpiBatch * kWG) {
          uint32_t v[kEpiBatch][8];
#pragma unroll
          for (int bb = 0; bb < kEpiBatch; ++bb) {
            const int t = t0 + bb * kWG;
            if (t < mt)
              tmem_ld8(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)acc * acc_cols + (uint32_t)(t * p.n_cols), v[bb]);
          }
          tmem_ld_wait();
#pragma unroll
          for (int bb = 0; bb < kEpiBatch; ++bb) {
            const int t = t0 + bb * kWG;
            if (t >= mt) break;
            const int iy = t >> p.mt_x_log2, ix = t & (p.mt_x - 1);
            const int y = (ty * p.mt_y + iy) * kTcTileH + r, x = (tx * p.mt_x + ix) * kTcTileW + px;
            const bool inside = (y < p.h) && (x < p.w);
            float o[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) o[k] = fmaxf(fmaf(__uint_as_float(v[bb][k]), sc[k], sh[k]), relu_floor);
            if constexpr (ST == 1) {
              if (inside) {
#pragma unroll
                for (int k = 0; k < 8; ++k) { st_s[k] += o[k]; st_q[k] = fmaf(o[k], o[k], st_q[k]); }
              }
            }
            if constexpr (HK > 0) {
              float z[HK];
              if constexpr ((HK & 1) == 0) {
                // even class count: the 1x1 conv runs on the packed f32x2 pipe (class pairs)
                float2 z2[HK / 2];
#pragma unroll
                for (int k = 0; k < HK / 2; ++k) z2[k] = *reinterpret_cast<const float2 *>(s_head + p.cout * HK + 2 * k);
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                  const float2 *wr = reinterpret_cast<const float2 *>(s_head + (col_base + c) * HK);
                  const float2 oo = make_float2(o[c], o[c]);
#pragma unroll
                  for (int k = 0; k < HK / 2; ++k) z2[k] = __ffma2_rn(oo, wr[k], z2[k]);
                }
#pragma unroll
                for (int k = 0; k < HK / 2; ++k) { z[2 * k] = z2[k].x; z[2 * k + 1] = z2[k].y; }
              } else {
#pragma unroll
                for (int k = 0; k < HK; ++k) z[k] = s_head[p.cout * HK + k];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                  const float *wr = s_head + (col_base + c) * HK;
#pragma unroll
                  for (int k = 0; k < HK; ++k) z[k] = fmaf(o[c], wr[k], z[k]);
                }
              }
              if (inside) {
                float mx = z[0];
#pragma unroll
                for (int k = 1; k < HK; ++k) mx = fmaxf(mx, z[k]);
                mx *= 1.4426950408889634f;          // softmax in base 2: exp(z - max) = 2^((z - max) * log2 e)
                float ssum = 0.f;
#pragma unroll
                for (int k = 0; k < HK; ++k) { z[k] = fast_exp2(fmaf(z[k], 1.4426950408889634f, -mx)); ssum += z[k]; }
                const float inv = fast_rcp(ssum);
                const long long pix = ((long long)img * p.h + y) * p.w + x;
                float pm = -1.f;
                int pa = 0;
#pragma unroll
                for (int k = 0; k < HK; ++k) {
                  z[k] *= inv;
                  if (z[k] > pm) { pm = z[k]; pa = k; }
                }
                if (p.probs) {
                  float *dst = p.probs + pix * HK;
                  if constexpr (HK == 4) *reinterpret_cast<float4 *>(dst) = make_float4(z[0], z[1], z[2], z[3]);
                  else {
#pragma unroll
                    for (int k = 0; k < HK; ++k) dst[k] = z[k];
                  }
                }
                if (p.labels) p.labels[pix] = (uint8_t)pa;
              }
            } else {
              uint4 pk;
              uint32_t *h2 = reinterpret_cast<uint32_t *>(&pk);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                h2[k] = bn_relu_pack2(v[bb][2 * k], v[bb][2 * k + 1], sc[2 * k], sc[2 * k + 1], sh[2 * k], sh[2 * k + 1], floor2, p.fp16);
              if (inside)
                *reinterpret_cast<uint4 *>(p.out + (long long)img * p.out_img_stride + (long long)(col_base >> 3) * plane_elems +
                                           ((long long)y * p.out_w + x) * 8) = pk;
              if (p.pool_out) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  uint32_t w0 = reinterpret_cast<uint32_t *>(&pk)[k];
                  w0 = max2(w0, __shfl_xor_sync(0xffffffffu, w0, 1), p.fp16);
                  w0 = max2(w0, __shfl_xor_sync(0xffffffffu, w0, 8), p.fp16);
                  h2[k] = w0;
                }
                if (inside && !(px & 1) && !(r & 1))
                  *reinterpret_cast<uint4 *>(p.pool_out + (long long)img * p.pool_img_stride +
                                             (long long)(col_base >> 3) * (plane_elems >> 2) +
                                             ((long long)(y >> 1) * (p.out_w >> 1) + (x >> 1)) * 8) = pk;
              }
            }
          }
        }
      } else if (EPI == 5 || EPI == 6) {
        // ---- row pairs: this thread's GEMM row is pixel (2r, x); columns [parity][channel] are the
        //      outputs (2r + parity, x).  EPI 5: 8 channels (2 chunks, two M-tiles per TMEM wait; plain /
        //      pool / fused head), EPI 6: 16 channels (4 chunks, one M-tile per wait; plain / pool).
        constexpr int PL = (EPI == 5) ? 1 : 2;         // output planes
        constexpr int kB = (EPI == 5) ? 2 : 1;
        constexpr int kWG = kTcEpiWarps / 4;
        for (int t0 = wg_cur; t0 < mt; t0 += kB * kWG) {
          uint32_t v[kB][2 * PL][8];
#pragma unroll
          for (int bb = 0; bb < kB; ++bb) {
            const int t = t0 + bb * kWG;
            if (t < mt) {
              const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)acc * acc_cols + (uint32_t)(t * p.n_cols);
#pragma unroll
              for (int c = 0; c < 2 * PL; ++c) tmem_ld8(ta + (uint32_t)(c * 8), v[bb][c]);
            }
          }
          tmem_ld_wait();
#pragma unroll
          for (int bb = 0; bb < kB; ++bb) {
            const int t = t0 + bb * kWG;
            if (t >= mt) break;
            const int iy = t >> p.mt_x_log2, ix = t & (p.mt_x - 1);
            const int yb = ((ty * p.mt_y + iy) * kTcTileH + r) * 2, x = (tx * p.mt_x + ix) * kTcTileW + px;
            uint4 pk[2][PL];
#pragma unroll
            for (int par = 0; par < 2; ++par) {
              const int y = yb + par;
              const bool inside = (y < p.h) && (x < p.w);
#pragma unroll
              for (int c8 = 0; c8 < PL; ++c8) {
                const int co0 = c8 * 8;
                const float4 sa = *reinterpret_cast<const float4 *>(s_scale + co0), sb = *reinterpret_cast<const float4 *>(s_scale + co0 + 4);
                const float4 ha = *reinterpret_cast<const float4 *>(s_shift + co0), hb = *reinterpret_cast<const float4 *>(s_shift + co0 + 4);
                const float sc[8] = {sa.x, sa.y, sa.z, sa.w, sb.x, sb.y, sb.z, sb.w};
                const float sh[8] = {ha.x, ha.y, ha.z, ha.w, hb.x, hb.y, hb.z, hb.w};
                float o[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) o[k] = fmaxf(fmaf(__uint_as_float(v[bb][par * PL + c8][k]), sc[k], sh[k]), relu_floor);
                if constexpr (ST == 1) {
                  if (inside) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) { st_s[c8 * 8 + k] += o[k]; st_q[c8 * 8 + k] = fmaf(o[k], o[k], st_q[c8 * 8 + k]); }
                  }
                }
                if constexpr (HK > 0) {
                  float z[HK];
#pragma unroll
                  for (int k = 0; k < HK; ++k) z[k] = s_head[p.cout * HK + k];
#pragma unroll
                  for (int c = 0; c < 8; ++c) {
                    const float *wr = s_head + (co0 + c) * HK;
#pragma unroll
                    for (int k = 0; k < HK; ++k) z[k] = fmaf(o[c], wr[k], z[k]);
                  }
                  if (inside) {
                    float mx = z[0];
#pragma unroll
                    for (int k = 1; k < HK; ++k) mx = fmaxf(mx, z[k]);
                    mx *= 1.4426950408889634f;
                    float ssum = 0.f;
#pragma unroll
                    for (int k = 0; k < HK; ++k) { z[k] = fast_exp2(fmaf(z[k], 1.4426950408889634f, -mx)); ssum += z[k]; }
                    const float inv = fast_rcp(ssum);
                    const long long pix = ((long long)img * p.h + y) * p.w + x;
                    float pm = -1.f;
                    int pa = 0;
#pragma unroll
                    for (int k = 0; k < HK; ++k) {
                      z[k] *= inv;
                      if (z[k] > pm) { pm = z[k]; pa = k; }
                    }
                    if (p.probs) {
                      float *dst = p.probs + pix * HK;
                      if constexpr (HK == 4) *reinterpret_cast<float4 *>(dst) = make_float4(z[0], z[1], z[2], z[3]);
                      else {
#pragma unroll
                        for (int k = 0; k < HK; ++k) dst[k] = z[k];
                      }
                    }
                    if (p.labels) p.labels[pix] = (uint8_t)pa;
                  }
                } else {
                  uint32_t *h2 = reinterpret_cast<uint32_t *>(&pk[par][c8]);
#pragma unroll
                  for (int k = 0; k < 4; ++k)
                    h2[k] = bn_relu_pack2(v[bb][par * PL + c8][2 * k], v[bb][par * PL + c8][2 * k + 1], sc[2 * k], sc[2 * k + 1],
                                          sh[2 * k], sh[2 * k + 1], floor2, p.fp16);
                  if (inside)
                    *reinterpret_cast<uint4 *>(p.out + (long long)img * p.out_img_stride + (long long)c8 * plane_elems +
                                               ((long long)y * p.out_w + x) * 8) = pk[par][c8];
                }
              }
            }
            if constexpr (HK == 0) {
              if (p.pool_out) {
                // 2x2 max: the vertical partner is the other parity of this thread, the horizontal one is lane ^ 1
#pragma unroll
                for (int c8 = 0; c8 < PL; ++c8) {
                  uint4 pm;
                  uint32_t *m2 = reinterpret_cast<uint32_t *>(&pm);
#pragma unroll
                  for (int k = 0; k < 4; ++k) {
                    uint32_t w0 = max2(reinterpret_cast<uint32_t *>(&pk[0][c8])[k], reinterpret_cast<uint32_t *>(&pk[1][c8])[k], p.fp16);
                    m2[k] = max2(w0, __shfl_xor_sync(0xffffffffu, w0, 1), p.fp16);
                  }
                  if (yb + 1 < p.h && x < p.w && !(px & 1))
                    *reinterpret_cast<uint4 *>(p.pool_out + (long long)img * p.pool_img_stride + (long long)c8 * (plane_elems >> 2) +
                                               ((long long)(yb >> 1) * (p.out_w >> 1) + (x >> 1)) * 8) = pm;
                }
              }
            }
          }
        }
      } else if (EPI == 7) {
        // ---- fp32-accurate split mode: every logical 8-channel output plane is the 16 accumulator columns
        //      [main 8 | corr 8] (main = hi x W_hi, corr = hi x W_lo' + lo' x W_hi, both scaled by the weight
        //      pre-scale): value = main + corr * 2^-11.  BN + ReLU in fp32, then either the fused fp32 head or
        //      the (hi, lo') fp16 plane pair (+ pooled pair) of the next layer.
        const int groups = min(p.n_cols, p.cols_valid - col_base) >> 4;
        constexpr int kWG = kTcEpiWarps / 4;
        for (int t = wg_cur; t < mt; t += kWG) {
          const int iy = t >> p.mt_x_log2, ix = t & (p.mt_x - 1);
          const int y = (ty * p.mt_y + iy) * kTcTileH + r, x = (tx * p.mt_x + ix) * kTcTileW + px;
          const bool inside = (y < p.h) && (x < p.w);
          const uint32_t t_base = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)acc * acc_cols + (uint32_t)(t * p.n_cols);
          float z[HK > 0 ? HK : 1];
          if constexpr (HK > 0) {
#pragma unroll
            for (int k = 0; k < HK; ++k) z[k] = s_head[p.cout * HK + k];
          }
          float amax = 0.f;
          if (HK == 0 && p.row_mul == 2) {
            // row pairs: this GEMM row is pixel (2r, x); logical columns [row parity][channel] -> physical groups
            // par * (C/8) + c8.  Both parities of a channel chunk are handled together (the 2x2 max needs them).
            const int C = p.scale_mod, cg8 = C >> 3;
            const int yb = 2 * y;
            const bool in0 = (yb < p.h) && (x < p.w), in1 = (yb + 1 < p.h) && (x < p.w);
            for (int c8 = 0; c8 < cg8; ++c8) {
              uint32_t v[2][16];
              tmem_ld16(t_base + (uint32_t)(c8 * 16), v[0]);
              tmem_ld16(t_base + (uint32_t)((cg8 + c8) * 16), v[1]);
              tmem_ld_wait();
              const int co0 = c8 * 8;
              const float4 sa = *reinterpret_cast<const float4 *>(s_scale + co0), sb = *reinterpret_cast<const float4 *>(s_scale + co0 + 4);
              const float4 ha = *reinterpret_cast<const float4 *>(s_shift + co0), hb = *reinterpret_cast<const float4 *>(s_shift + co0 + 4);
              const float sc[8] = {sa.x, sa.y, sa.z, sa.w, sb.x, sb.y, sb.z, sb.w};
              const float sh[8] = {ha.x, ha.y, ha.z, ha.w, hb.x, hb.y, hb.z, hb.w};
              float o[2][8];
#pragma unroll
              for (int par = 0; par < 2; ++par)
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                  const float val = fmaf(__uint_as_float(v[par][8 + k]), 4.8828125e-4f, __uint_as_float(v[par][k]));
                  o[par][k] = fmaxf(fmaf(val, sc[k], sh[k]), relu_floor);
                  amax = fmaxf(amax, fabsf(o[par][k]));
                }
              uint4 hi, lo;
              __nv_bfloat16 *dst = p.out + (long long)img * p.out_img_stride + (long long)(co0 >> 2) * plane_elems +
                                   ((long long)yb * p.out_w + x) * 8;
              split_pack8(o[0], hi, lo);
              if (in0) { *reinterpret_cast<uint4 *>(dst) = hi; *reinterpret_cast<uint4 *>(dst + plane_elems) = lo; }
              split_pack8(o[1], hi, lo);
              if (in1) {
                *reinterpret_cast<uint4 *>(dst + (long long)p.out_w * 8) = hi;
                *reinterpret_cast<uint4 *>(dst + (long long)p.out_w * 8 + plane_elems) = lo;
              }
              if (p.pool_out) {
                float m[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                  m[k] = fmaxf(o[0][k], o[1][k]);
                  m[k] = fmaxf(m[k], __shfl_xor_sync(0xffffffffu, m[k], 1));
                }
                split_pack8(m, hi, lo);
                if (in1 && !(px & 1)) {
                  __nv_bfloat16 *pd = p.pool_out + (long long)img * p.pool_img_stride + (long long)(co0 >> 2) * (plane_elems >> 2) +
                                      ((long long)y * (p.out_w >> 1) + (x >> 1)) * 8;
                  *reinterpret_cast<uint4 *>(pd) = hi;
                  *reinterpret_cast<uint4 *>(pd + (plane_elems >> 2)) = lo;
                }
              }
            }
            if ((in0 || in1) && !(amax <= 65000.f) && p.overflow) atomicOr(p.overflow, 1);
            continue;
          }
          for (int j = 0; j < groups; j += 2) {
            uint32_t v[2][16];
            const bool two = (j + 1 < groups);
            tmem_ld16(t_base + (uint32_t)(j * 16), v[0]);
            if (two) tmem_ld16(t_base + (uint32_t)(j * 16 + 16), v[1]);
            tmem_ld_wait();
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              if (u == 1 && !two) break;
              const int colp = col_base + (j + u) * 16;          // physical column of the group
              int co0 = (colp >> 4) << 3;                        // logical channel of its first column
              long long off = ((long long)y * p.out_w + x) * 8;
              if (p.mode == 1) {
                const int par = colp / p.cout;                   // p.cout = physical columns per parity
                co0 = ((colp - par * p.cout) >> 4) << 3;
                off = ((long long)(2 * y + (par >> 1)) * p.out_w + 2 * x + (par & 1)) * 8;
              }
              const float4 sa = *reinterpret_cast<const float4 *>(s_scale + co0), sb = *reinterpret_cast<const float4 *>(s_scale + co0 + 4);
              const float4 ha = *reinterpret_cast<const float4 *>(s_shift + co0), hb = *reinterpret_cast<const float4 *>(s_shift + co0 + 4);
              const float sc[8] = {sa.x, sa.y, sa.z, sa.w, sb.x, sb.y, sb.z, sb.w};
              const float sh[8] = {ha.x, ha.y, ha.z, ha.w, hb.x, hb.y, hb.z, hb.w};
              float o[8];
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                const float val = fmaf(__uint_as_float(v[u][8 + k]), 4.8828125e-4f, __uint_as_float(v[u][k]));
                o[k] = fmaxf(fmaf(val, sc[k], sh[k]), relu_floor);
              }
              if constexpr (HK > 0) {
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                  const float *wr = s_head + (co0 + c) * HK;
#pragma unroll
                  for (int k = 0; k < HK; ++k) z[k] = fmaf(o[c], wr[k], z[k]);
                }
              } else {
#pragma unroll
                for (int k = 0; k < 8; ++k) amax = fmaxf(amax, fabsf(o[k]));
                uint4 hi, lo;
                split_pack8(o, hi, lo);
                __nv_bfloat16 *dst = p.out + (long long)img * p.out_img_stride + (long long)(co0 >> 2) * plane_elems + off;
                if (inside) {
                  *reinterpret_cast<uint4 *>(dst) = hi;
                  *reinterpret_cast<uint4 *>(dst + plane_elems) = lo;
                }
                if (p.pool_out) {
                  // 2x2 max on the fp32 values (partners: lanes ^1 in x, ^8 in y), then split again
#pragma unroll
                  for (int k = 0; k < 8; ++k) {
                    o[k] = fmaxf(o[k], __shfl_xor_sync(0xffffffffu, o[k], 1));
                    o[k] = fmaxf(o[k], __shfl_xor_sync(0xffffffffu, o[k], 8));
                  }
                  split_pack8(o, hi, lo);
                  if (inside && !(px & 1) && !(r & 1)) {
                    __nv_bfloat16 *pd = p.pool_out + (long long)img * p.pool_img_stride + (long long)(co0 >> 2) * (plane_elems >> 2) +
                                        ((long long)(y >> 1) * (p.out_w >> 1) + (x >> 1)) * 8;
                    *reinterpret_cast<uint4 *>(pd) = hi;
                    *reinterpret_cast<uint4 *>(pd + (plane_elems >> 2)) = lo;
                  }
                }
              }
            }
          }
          if constexpr (HK > 0) {
            if (inside) {
              float mx = z[0];
#pragma unroll
              for (int k = 1; k < HK; ++k) mx = fmaxf(mx, z[k]);
              float ssum = 0.f;
#pragma unroll
              for (int k = 0; k < HK; ++k) { z[k] = expf(z[k] - mx); ssum += z[k]; }
              const float inv = 1.f / ssum;
              const long long pix = ((long long)img * p.h + y) * p.w + x;
              float pm = -1.f;
              int pa = 0;
#pragma unroll
              for (int k = 0; k < HK; ++k) {
                z[k] *= inv;
                if (z[k] > pm) { pm = z[k]; pa = k; }   // first max, computed on the float32 probabilities
              }
              if (p.probs) {
                float *dst = p.probs + pix * HK;
                if constexpr (HK == 4) *reinterpret_cast<float4 *>(dst) = make_float4(z[0], z[1], z[2], z[3]);
                else {
#pragma unroll
                  for (int k = 0; k < HK; ++k) dst[k] = z[k];
                }
              }
              if (p.labels) p.labels[pix] = (uint8_t)pa;
            }
          } else {
            if (inside && !(amax <= 65000.f) && p.overflow) atomicOr(p.overflow, 1);   // also catches NaN
          }
        }
      } else if (EPI == 8) {
        // ---- training forward: z = acc * scale + shift (scale = 1, shift = conv bias), stored as it is (no ReLU:
        //      BatchNorm comes first), and the per-channel batch statistics sum(z), sum(z^2) of the fp32 values --
        //      the separate pass that re-read z (bn_stats_kernel) is gone.  Column -> channel: col % scale_mod
        //      (row pairs: [row parity][channel]; up-conv: [pixel parity][channel]).  Layers with <= 16 channels keep
        //      their partial sums in registers for the whole kernel; wider layers reduce every 8-column chunk
        //      over the warp (7 shuffles) and add it to the CTA's shared-memory accumulators.
        const int C = p.scale_mod;
        const bool few = (C <= 16);
        constexpr int kWG = kTcEpiWarps / 4;
        for (int t = wg_cur; t < mt; t += kWG) {
          const int iy = t >> p.mt_x_log2, ix = t & (p.mt_x - 1);
          const int yb = ((ty * p.mt_y + iy) * kTcTileH + r) * p.row_mul, x = (tx * p.mt_x + ix) * kTcTileW + px;
          const uint32_t t_base = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)acc * acc_cols + (uint32_t)(t * p.n_cols);
          for (int j = 0; j < my_nch; j += 2) {
            uint32_t v[2][8];
            const bool two = (j + 1 < my_nch);
            tmem_ld8(t_base + (uint32_t)(j * 8), v[0]);
            if (two) tmem_ld8(t_base + (uint32_t)(j * 8 + 8), v[1]);
            tmem_ld_wait();
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              if (u == 1 && !two) break;
              const int col = col_base + (j + u) * 8;
              int co0, y = yb, xo = x;
              if (p.mode == 1) { const int par = col / p.cout; co0 = col - par * p.cout; y = 2 * yb + (par >> 1); xo = 2 * x + (par & 1); }
              else if (p.row_mul == 2) { const int par = col / C; co0 = col - par * C; y = yb + par; }
              else co0 = col;
              const bool inside = (y < p.out_h) && (xo < p.out_w);
              const float4 sa = *reinterpret_cast<const float4 *>(s_scale + co0), sb = *reinterpret_cast<const float4 *>(s_scale + co0 + 4);
              const float4 ha = *reinterpret_cast<const float4 *>(s_shift + co0), hb = *reinterpret_cast<const float4 *>(s_shift + co0 + 4);
              const float sc[8] = {sa.x, sa.y, sa.z, sa.w, sb.x, sb.y, sb.z, sb.w};
              const float sh[8] = {ha.x, ha.y, ha.z, ha.w, hb.x, hb.y, hb.z, hb.w};
              float o[8];
#pragma unroll
              for (int k = 0; k < 8; ++k) o[k] = fmaxf(fmaf(__uint_as_float(v[u][k]), sc[k], sh[k]), relu_floor);
              uint4 pk;
              uint32_t *h2 = reinterpret_cast<uint32_t *>(&pk);
#pragma unroll
              for (int k = 0; k < 4; ++k) h2[k] = pack2(o[2 * k], o[2 * k + 1], p.fp16);
              if (inside)
                *reinterpret_cast<uint4 *>(p.out + (long long)img * p.out_img_stride + (long long)(co0 >> 3) * plane_elems +
                                           ((long long)y * p.out_w + xo) * 8) = pk;
              else {
#pragma unroll
                for (int k = 0; k < 8; ++k) o[k] = 0.f;
              }
              if (few) {
                if (co0 == 0) {
#pragma unroll
                  for (int k = 0; k < 8; ++k) { st_s[k] += o[k]; st_q[k] = fmaf(o[k], o[k], st_q[k]); }
                } else {
#pragma unroll
                  for (int k = 0; k < 8; ++k) { st_s[8 + k] += o[k]; st_q[8 + k] = fmaf(o[k], o[k], st_q[8 + k]); }
                }
              } else {
                float o2[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) o2[k] = o[k] * o[k];
                const float ws = warp_sum8(o, lane), wq = warp_sum8(o2, lane);
                if ((lane & 3) == 0) {
                  atomicAdd(s_head + co0 + warp_sum8_index(lane), ws);
                  atomicAdd(s_head + C + co0 + warp_sum8_index(lane), wq);
                }
              }
            }
          }
        }
      } else if (EPI == 2 && HK == 0) {
        // ---- 16-channel layers: same idea, two chunks (planes) per M-tile and two M-tiles per wait
        constexpr int kB2 = 2;
        constexpr int kWG = kTcEpiWarps / 4;
        for (int t0 = wg_cur; t0 < mt; t0 += kB2 * kWG) {
          uint32_t v[kB2][2][8];
#pragma unroll
          for (int bb = 0; bb < kB2; ++bb) {
            const int t = t0 + bb * kWG;
            if (t < mt) {
              const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)acc * acc_cols + (uint32_t)(t * p.n_cols);
              tmem_ld8(ta, v[bb][0]);
              tmem_ld8(ta + 8u, v[bb][1]);
            }
          }
          tmem_ld_wait();
#pragma unroll
          for (int bb = 0; bb < kB2; ++bb) {
            const int t = t0 + bb * kWG;
            if (t >= mt) break;
            const int iy = t >> p.mt_x_log2, ix = t & (p.mt_x - 1);
            const int y = (ty * p.mt_y + iy) * kTcTileH + r, x = (tx * p.mt_x + ix) * kTcTileW + px;
            const bool inside = (y < p.h) && (x < p.w);
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const int co0 = col_base + u * 8;
              const float4 sa = *reinterpret_cast<const float4 *>(s_scale + co0), sb = *reinterpret_cast<const float4 *>(s_scale + co0 + 4);
              const float4 ha = *reinterpret_cast<const float4 *>(s_shift + co0), hb = *reinterpret_cast<const float4 *>(s_shift + co0 + 4);
              const float sc[8] = {sa.x, sa.y, sa.z, sa.w, sb.x, sb.y, sb.z, sb.w};
              const float sh[8] = {ha.x, ha.y, ha.z, ha.w, hb.x, hb.y, hb.z, hb.w};
              uint4 pk;
              uint32_t *h2 = reinterpret_cast<uint32_t *>(&pk);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                h2[k] = bn_relu_pack2(v[bb][u][2 * k], v[bb][u][2 * k + 1], sc[2 * k], sc[2 * k + 1], sh[2 * k], sh[2 * k + 1], floor2, p.fp16);
              if constexpr (ST == 1) {
                if (inside) {
#pragma unroll
                  for (int k = 0; k < 8; ++k) {
                    const float o = fmaxf(fmaf(__uint_as_float(v[bb][u][k]), sc[k], sh[k]), relu_floor);
                    st_s[u * 8 + k] += o; st_q[u * 8 + k] = fmaf(o, o, st_q[u * 8 + k]);
                  }
                }
              }
              if (inside)
                *reinterpret_cast<uint4 *>(p.out + (long long)img * p.out_img_stride + (long long)(co0 >> 3) * plane_elems +
                                           ((long long)y * p.out_w + x) * 8) = pk;
              if (p.pool_out) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  uint32_t w0 = h2[k];
                  w0 = max2(w0, __shfl_xor_sync(0xffffffffu, w0, 1), p.fp16);
                  w0 = max2(w0, __shfl_xor_sync(0xffffffffu, w0, 8), p.fp16);
                  h2[k] = w0;
                }
                if (inside && !(px & 1) && !(r & 1))
                  *reinterpret_cast<uint4 *>(p.pool_out + (long long)img * p.pool_img_stride +
                                             (long long)(co0 >> 3) * (plane_elems >> 2) +
                                             ((long long)(y >> 1) * (p.out_w >> 1) + (x >> 1)) * 8) = pk;
              }
            }
          }
        }
      } else
      for (int t = wg_cur; t < mt; t += kTcEpiWarps / 4) {
        const int iy = t >> p.mt_x_log2, ix = t & (p.mt_x - 1);
        const int y = (ty * p.mt_y + iy) * kTcTileH + r, x = (tx * p.mt_x + ix) * kTcTileW + px;
        const bool inside = (y < p.h) && (x < p.w);
        const uint32_t t_base = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)acc * acc_cols +
                                (uint32_t)(t * p.n_cols);
        if constexpr (HK > 0) {
          // ---- fused 1x1 conv + softmax head (fp32 activations never leave registers)
          float z[HK];
#pragma unroll
          for (int k = 0; k < HK; ++k) z[k] = s_head[p.cout * HK + k];
          for (int j = 0; j < my_nch; j += 2) {
            uint32_t v[2][8];
            const bool two = (j + 1 < my_nch);
            tmem_ld8(t_base + (uint32_t)(j * 8), v[0]);
            if (two) tmem_ld8(t_base + (uint32_t)(j * 8 + 8), v[1]);
            tmem_ld_wait();
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              if (u == 1 && !two) break;
              const int co0 = col_base + (j + u) * 8;
              const float4 sa = *reinterpret_cast<const float4 *>(s_scale + co0), sb = *reinterpret_cast<const float4 *>(s_scale + co0 + 4);
              const float4 ha = *reinterpret_cast<const float4 *>(s_shift + co0), hb = *reinterpret_cast<const float4 *>(s_shift + co0 + 4);
              const float sc[8] = {sa.x, sa.y, sa.z, sa.w, sb.x, sb.y, sb.z, sb.w};
              const float sh[8] = {ha.x, ha.y, ha.z, ha.w, hb.x, hb.y, hb.z, hb.w};
#pragma unroll
              for (int c = 0; c < 8; ++c) {
                const float o = fmaxf(fmaf(__uint_as_float(v[u][c]), sc[c], sh[c]), relu_floor);
                const float *wr = s_head + (co0 + c) * HK;
#pragma unroll
                for (int k = 0; k < HK; ++k) z[k] = fmaf(o, wr[k], z[k]);
              }
            }
          }
          if (inside) {
            float mx = z[0];
#pragma unroll
            for (int k = 1; k < HK; ++k) mx = fmaxf(mx, z[k]);
            float ssum = 0.f;
#pragma unroll
            for (int k = 0; k < HK; ++k) { z[k] = expf(z[k] - mx); ssum += z[k]; }
            const float inv = 1.f / ssum;
            const long long pix = ((long long)img * p.h + y) * p.w + x;
            float pm = -1.f;
            int pa = 0;
#pragma unroll
            for (int k = 0; k < HK; ++k) {
              z[k] *= inv;
              if (z[k] > pm) { pm = z[k]; pa = k; }   // first max, computed on the float32 probabilities
            }
            if (p.probs) {
              float *dst = p.probs + pix * HK;
              if constexpr (HK == 4) *reinterpret_cast<float4 *>(dst) = make_float4(z[0], z[1], z[2], z[3]);
              else {
#pragma unroll
                for (int k = 0; k < HK; ++k) dst[k] = z[k];
              }
            }
            if (p.labels) p.labels[pix] = (uint8_t)pa;
          }
        } else if (EPI == 4) {
          // ---- up-conv pixel shuffle: the two x-parities of one (y-parity, plane) are loaded
          //      together so each thread stores 32 contiguous bytes (2 output pixels)
          const int planes_per_par = p.cout >> 3;
          const int first = col_base / p.cout;                   // first parity in this n-tile (even)
          const int npar = (my_nch << 3) / p.cout;               // parities in this n-tile (2 or 4)
          for (int pp = 0; pp < npar; pp += 2) {
            const int py = (first + pp) >> 1;
            for (int c8 = 0; c8 < planes_per_par; ++c8) {
              uint32_t v[2][8];
              tmem_ld8(t_base + (uint32_t)(pp * p.cout + c8 * 8), v[0]);
              tmem_ld8(t_base + (uint32_t)((pp + 1) * p.cout + c8 * 8), v[1]);
              tmem_ld_wait();
              const int co0 = c8 * 8;
              const float4 sa = *reinterpret_cast<const float4 *>(s_scale + co0), sb = *reinterpret_cast<const float4 *>(s_scale + co0 + 4);
              const float4 ha = *reinterpret_cast<const float4 *>(s_shift + co0), hb = *reinterpret_cast<const float4 *>(s_shift + co0 + 4);
              const float sc[8] = {sa.x, sa.y, sa.z, sa.w, sb.x, sb.y, sb.z, sb.w};
              const float sh[8] = {ha.x, ha.y, ha.z, ha.w, hb.x, hb.y, hb.z, hb.w};
              uint4 pk[2];
#pragma unroll
              for (int u = 0; u < 2; ++u) {
                uint32_t *h2 = reinterpret_cast<uint32_t *>(&pk[u]);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  h2[k] = bn_relu_pack2(v[u][2 * k], v[u][2 * k + 1], sc[2 * k], sc[2 * k + 1], sh[2 * k], sh[2 * k + 1], floor2, p.fp16);
              }
              if constexpr (ST == 1) {
                // training forward of the narrow up-convs (<= 16 channels per parity): batch statistics in registers
                if (inside) {
#pragma unroll
                  for (int u = 0; u < 2; ++u)
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                      const float o = fmaxf(fmaf(__uint_as_float(v[u][k]), sc[k], sh[k]), relu_floor);
                      if (c8 == 0) { st_s[k] += o; st_q[k] = fmaf(o, o, st_q[k]); }
                      else { st_s[8 + k] += o; st_q[8 + k] = fmaf(o, o, st_q[8 + k]); }
                    }
                }
              }
              if (inside) {
                __nv_bfloat16 *dst = p.out + (long long)img * p.out_img_stride + (long long)c8 * plane_elems +
                                     ((long long)(2 * y + py) * p.out_w + 2 * x) * 8;
                *reinterpret_cast<uint4 *>(dst) = pk[0];
                *reinterpret_cast<uint4 *>(dst + 8) = pk[1];
              }
            }
          }
        } else if (EPI == 3) {
          // ---- stem groups: a GEMM row is 8 adjacent pixels of one image row and the 64 columns of a
          //      plane are the 64 contiguous elements [pixel][channel] of the blocked layout, i.e. every
          //      thread owns one whole 128-byte line.  Storing straight from registers would touch 32
          //      lines per instruction, so the warp transposes through its own 4 KB of shared memory
          //      (16-byte slots XOR-swizzled by row: conflict-free both ways) and writes 512 contiguous
          //      bytes per instruction.
          uint8_t *stg = reinterpret_cast<uint8_t *>(s_scale) + p.stage_off + (size_t)(warp - 4) * 4096;
          const uint32_t stg_u32 = smem_u32(stg);
          for (int j0 = 0; j0 < my_nch; j0 += 8) {
            // software-pipelined TMEM reads: the loads of chunk pair k+1 are in flight while pair k is
            // scaled, packed and staged
            uint32_t v[2][2][8];
            tmem_ld8(t_base + (uint32_t)(j0 * 8), v[0][0]);
            tmem_ld8(t_base + (uint32_t)(j0 * 8 + 8), v[0][1]);
#pragma unroll
            for (int jj = 0; jj < 8; jj += 2) {
              const int cur = (jj >> 1) & 1;
              tmem_ld_wait();
              if (jj + 2 < 8) {
                tmem_ld8(t_base + (uint32_t)((j0 + jj + 2) * 8), v[cur ^ 1][0]);
                tmem_ld8(t_base + (uint32_t)((j0 + jj + 2) * 8 + 8), v[cur ^ 1][1]);
              }
#pragma unroll
              for (int u = 0; u < 2; ++u) {
                const int col = col_base + (j0 + jj + u) * 8;
                const float4 sa = *reinterpret_cast<const float4 *>(s_scale + col), sb = *reinterpret_cast<const float4 *>(s_scale + col + 4);
                const float4 ha = *reinterpret_cast<const float4 *>(s_shift + col), hb = *reinterpret_cast<const float4 *>(s_shift + col + 4);
                const float sc[8] = {sa.x, sa.y, sa.z, sa.w, sb.x, sb.y, sb.z, sb.w};
                const float sh[8] = {ha.x, ha.y, ha.z, ha.w, hb.x, hb.y, hb.z, hb.w};
                uint32_t h2[4];
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  h2[k] = bn_relu_pack2(v[cur][u][2 * k], v[cur][u][2 * k + 1], sc[2 * k], sc[2 * k + 1], sh[2 * k], sh[2 * k + 1], floor2, p.fp16);
                const uint32_t slot = (uint32_t)((jj + u) ^ (lane & 7));
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stg_u32 + (uint32_t)lane * 128u + slot * 16u),
                             "r"(h2[0]), "r"(h2[1]), "r"(h2[2]), "r"(h2[3]) : "memory");
              }
            }
            __syncwarp();
            const long long plane_off = (long long)((col_base >> 6) + (j0 >> 3)) * (plane_elems << 3);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int rr = i * 4 + (lane >> 3), c = lane & 7;          // staged row, 16-byte slot
              uint32_t a, b, cc, d;
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(cc), "=r"(d)
                           : "r"(stg_u32 + (uint32_t)rr * 128u + (uint32_t)((c ^ (rr & 7)) * 16)) : "memory");
              const int y2 = (ty * p.mt_y + iy) * kTcTileH + q * 4 + (rr >> 3);
              const int x2 = (tx * p.mt_x + ix) * kTcTileW + (rr & 7);
              if (y2 < p.h && x2 < p.w)
                *reinterpret_cast<uint4 *>(p.out + (long long)img * p.out_img_stride + plane_off +
                                           (((long long)y2 * p.out_w + x2) << 6) + c * 8) = make_uint4(a, b, cc, d);
            }
            __syncwarp();
          }
        } else {
          // ---- plain store (+ optional fused 2x2 max-pool); generic pixel shuffle falls back here
          const long long px_off = ((long long)y * p.out_w + x) * 8;
          for (int j = 0; j < my_nch; j += 2) {
            uint32_t v[2][8];
            const bool two = (j + 1 < my_nch);
            tmem_ld8(t_base + (uint32_t)(j * 8), v[0]);
            if (two) tmem_ld8(t_base + (uint32_t)(j * 8 + 8), v[1]);
            tmem_ld_wait();
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              if (u == 1 && !two) break;
              const int col = col_base + (j + u) * 8;
              int co0 = col;
              long long off = px_off;
              if (p.mode == 1) {
                const int par = col / p.cout;
                co0 = col - par * p.cout;
                off = ((long long)(2 * y + (par >> 1)) * p.out_w + 2 * x + (par & 1)) * 8;
              }
              const float4 sa = *reinterpret_cast<const float4 *>(s_scale + co0), sb = *reinterpret_cast<const float4 *>(s_scale + co0 + 4);
              const float4 ha = *reinterpret_cast<const float4 *>(s_shift + co0), hb = *reinterpret_cast<const float4 *>(s_shift + co0 + 4);
              const float sc[8] = {sa.x, sa.y, sa.z, sa.w, sb.x, sb.y, sb.z, sb.w};
              const float sh[8] = {ha.x, ha.y, ha.z, ha.w, hb.x, hb.y, hb.z, hb.w};
              uint4 pk;
              uint32_t *h2 = reinterpret_cast<uint32_t *>(&pk);
              if constexpr (EPI == 9) {
                // data gradient of the up-conv: only the 2x2 sums of the fp32 outputs are stored (partners: lanes ^1, ^8)
                float o[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                  o[k] = fmaxf(fmaf(__uint_as_float(v[u][k]), sc[k], sh[k]), relu_floor);
                  o[k] += __shfl_xor_sync(0xffffffffu, o[k], 1);
                  o[k] += __shfl_xor_sync(0xffffffffu, o[k], 8);
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) h2[k] = pack2(o[2 * k], o[2 * k + 1], p.fp16);
                if (inside && !(px & 1) && !(r & 1))
                  *reinterpret_cast<uint4 *>(p.pool_out + (long long)img * p.pool_img_stride +
                                             (long long)(co0 >> 3) * (plane_elems >> 2) +
                                             ((long long)(y >> 1) * (p.out_w >> 1) + (x >> 1)) * 8) = pk;
                continue;
              }
#pragma unroll
              for (int k = 0; k < 4; ++k)
                h2[k] = bn_relu_pack2(v[u][2 * k], v[u][2 * k + 1], sc[2 * k], sc[2 * k + 1], sh[2 * k], sh[2 * k + 1], floor2, p.fp16);
              if (inside)
                *reinterpret_cast<uint4 *>(p.out + (long long)img * p.out_img_stride + (long long)(co0 >> 3) * plane_elems + off) = pk;
              if (EPI != 9 && p.pool_out) {
                // 2x2 max on the packed bf16 pairs (max commutes with the monotonic rounding):
                // partners are lanes ^1 (x) and ^8 (row) of this warp
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  uint32_t w0 = reinterpret_cast<uint32_t *>(&pk)[k];
                  w0 = max2(w0, __shfl_xor_sync(0xffffffffu, w0, 1), p.fp16);
                  w0 = max2(w0, __shfl_xor_sync(0xffffffffu, w0, 8), p.fp16);
                  h2[k] = w0;
                }
                if (inside && !(px & 1) && !(r & 1))
                  *reinterpret_cast<uint4 *>(p.pool_out + (long long)img * p.pool_img_stride +
                                             (long long)(co0 >> 3) * (plane_elems >> 2) +
                                             ((long long)(y >> 1) * (p.out_w >> 1) + (x >> 1)) * 8) = pk;
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bars->acc_empty[acc]));
      if (++acc == 2) { acc = 0; accph ^= 1u; }
    }
    if constexpr (EPI == 8 || ST == 1) {
      if (p.scale_mod <= 16) {
        const int C = p.scale_mod;
        float g[8];
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          if (half * 8 >= C) break;
#pragma unroll
          for (int k = 0; k < 8; ++k) g[k] = st_s[half * 8 + k];
          const float ws = warp_sum8(g, lane);
#pragma unroll
          for (int k = 0; k < 8; ++k) g[k] = st_q[half * 8 + k];
          const float wq = warp_sum8(g, lane);
          if ((lane & 3) == 0) {
            atomicAdd(s_head + half * 8 + warp_sum8_index(lane), ws);
            atomicAdd(s_head + C + half * 8 + warp_sum8_index(lane), wq);
          }
        }
      }
    }
    if (p.dbg && blockIdx.x == 0 && warp == 4 && lane == 0) { p.dbg[6] = w_epi; p.dbg[7] = clock64() - t_start; }
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (EPI == 8 || ST == 1) {
    if (p.stats)
      for (int i = threadIdx.x; i < 2 * p.scale_mod; i += blockDim.x) atomicAdd(p.stats + i, (double)s_head[i]);
  }
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

// ----------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------
bool tc_head_fusable(int num_classes);
static inline int floordiv2(int v) { return v >= 0 ? v / 2 : -((-v + 1) / 2); }

bool tc_supported(int kh, int kw, int cin, int cout, int ups, int h, int w) {
  if (cin % 8 || cout % 8) return false;
  const int cg = cin / 8;
  if (cg != 1 && (cg & 1)) return false;
  if (cg > 8 && cg % 8) return false;
  if (kh < 1 || kw < 1 || kh > 3 || kw > 3) return false;
  if (cg == 1 && kh * kw == 1) return false;
  // (h, w) is the GEMM-row grid (the low-res grid for the up-conv)
  if (h < kTcTileH || w < kTcTileW) return false;
  (void)ups;
  return true;
}

int tc_make_geometry(int kh, int kw, int cin, int cout, int ups, TcGeometry *g, int pad_top, int pad_left) {
  std::memset(g, 0, sizeof(*g));
  g->kh = kh; g->kw = kw; g->cin = cin; g->cout = cout; g->ups = ups;
  g->cin_l = cin; g->cout_l = cout; g->wscale = 1.f;
  const int pt = pad_top >= 0 ? pad_top : (kh - 1) / 2, pl = pad_left >= 0 ? pad_left : (kw - 1) / 2;
  g->pt = pt; g->pl = pl;
  if (!ups) {
    g->dy_min = -pt; g->dy_max = kh - 1 - pt; g->dx_min = -pl; g->dx_max = kw - 1 - pl;
  } else {
    g->dy_min = floordiv2(0 - pt); g->dy_max = floordiv2(1 + kh - 1 - pt);
    g->dx_min = floordiv2(0 - pl); g->dx_max = floordiv2(1 + kw - 1 - pl);
  }
  const int nty = g->dy_max - g->dy_min + 1, ntx = g->dx_max - g->dx_min + 1;
  g->box_h = nty - 1;   // halo rows / px added to the super-tile
  g->box_w = ntx - 1;
  const int cg = cin / 8;
  g->planes_per_chunk = std::min(cg, 8);
  g->cin_chunks = cg / g->planes_per_chunk;
  const int cols = ups ? 4 * cout : cout;
  g->cols_valid = cols;
  const int cols_pad = (cols + 15) / 16 * 16;
  int nc = 256;
  while (cols_pad % nc) nc >>= 1;
  g->n_cols = nc;
  g->n_tiles_n = cols_pad / nc;
  int ks = 0;
  if (g->planes_per_chunk == 1) {
    // pair taps in raster order (byte offsets increase, so LBO stays positive)
    std::vector<std::pair<int, int>> taps;
    for (int ty = 0; ty < nty; ++ty)
      for (int tx = 0; tx < ntx; ++tx) taps.push_back({ty, tx});
    if ((int)(taps.size() + 1) / 2 > kTcMaxKSteps) return 1;
    size_t i = 0;
    if (taps.size() & 1) {
      // odd count: first step = (zero-weight dummy, tap0) is impossible (nothing before
      // tap 0), so put the dummy in front of the LAST tap: (last-1 position, last)
      for (; i + 2 < taps.size(); i += 2) {
        g->half_ty[ks][0] = taps[i].first; g->half_tx[ks][0] = taps[i].second; g->half_pl[ks][0] = 0;
        g->half_ty[ks][1] = taps[i + 1].first; g->half_tx[ks][1] = taps[i + 1].second; g->half_pl[ks][1] = 0;
        ++ks;
      }
      const auto last = taps.back();
      int dty = last.first, dtx = last.second - 1;
      if (dtx < 0) { dtx = last.second; dty = last.first - 1; }
      if (dty < 0) return 1;
      g->half_ty[ks][0] = -1 - dty; g->half_tx[ks][0] = dtx; g->half_pl[ks][0] = 0;   // dummy (encoded ty<0)
      g->half_ty[ks][1] = last.first; g->half_tx[ks][1] = last.second; g->half_pl[ks][1] = 0;
      ++ks;
    } else {
      for (; i < taps.size(); i += 2) {
        g->half_ty[ks][0] = taps[i].first; g->half_tx[ks][0] = taps[i].second; g->half_pl[ks][0] = 0;
        g->half_ty[ks][1] = taps[i + 1].first; g->half_tx[ks][1] = taps[i + 1].second; g->half_pl[ks][1] = 0;
        ++ks;
      }
    }
  } else {
    if (nty * ntx * (g->planes_per_chunk / 2) > kTcMaxKSteps) return 1;     // would overrun the per-k-step tables
    for (int ty = 0; ty < nty; ++ty)
      for (int tx = 0; tx < ntx; ++tx)
        for (int jj = 0; jj < g->planes_per_chunk / 2; ++jj) {
          for (int hf = 0; hf < 2; ++hf) {
            g->half_ty[ks][hf] = ty; g->half_tx[ks][hf] = tx; g->half_pl[ks][hf] = 2 * jj + hf;
          }
          ++ks;
        }
  }
  if (ks > kTcMaxKSteps) return 1;
  g->ksteps = ks;
  g->bgroup = (ks * 32 * nc <= 32768) ? ks : std::max(1, g->planes_per_chunk / 2);   // (>= 1: one-plane chunks pair taps)
  if (ks % g->bgroup) return 1;
  return 0;
}

int tc_make_geometry_s2d(int cz, int cout, TcGeometry *g) {
  std::memset(g, 0, sizeof(*g));
  if (cz % 8 || cout % 8 || cz > 32 || (cz != 8 && (cz / 8) % 2)) return 1;
  g->kh = 3; g->kw = 3; g->cin = cz; g->cout = cout; g->ups = 0; g->s2d = 1;
  g->cin_l = cz; g->cout_l = cout; g->wscale = 1.f;
  g->pt = 1; g->pl = 1;
  g->dy_min = -1; g->dy_max = 0; g->dx_min = -1; g->dx_max = 0;     // one halo row / px BEFORE the tile (taps a, b = -1)
  g->box_h = 1; g->box_w = 1;
  const int P = cz / 8;
  g->planes_per_chunk = 4 * P;       // "planes" of the stage: [parity][plane]
  g->cin_chunks = 1;
  g->cols_valid = cout;
  const int cols_pad = (cout + 15) / 16 * 16;
  int nc = 256;
  while (cols_pad % nc) nc >>= 1;
  g->n_cols = nc;
  g->n_tiles_n = cols_pad / nc;
  struct Half { int v, ty, tx; };
  std::vector<Half> halves;
  for (int par = 0; par < 4; ++par) {              // ascending smem offset: (parity, plane, row, px)
    const int py = par >> 1, px = par & 1;
    for (int pl8 = 0; pl8 < P; ++pl8)
      for (int ty = (py ? 0 : 1); ty < 2; ++ty)     // parity 0 rows / columns carry only the tap a = 0 (box offset 1)
        for (int tx = (px ? 0 : 1); tx < 2; ++tx) halves.push_back({par * P + pl8, ty, tx});
  }
  int ks = 0;
  if (P == 1) {
    // 9 single-plane halves: consecutive pairs (LBO = offset difference > 0); the odd one gets a zero-weight dummy in front
    size_t i = 0;
    for (; i + 2 < halves.size() || (halves.size() % 2 == 0 && i < halves.size()); i += 2) {
      for (int hf = 0; hf < 2; ++hf) {
        g->half_ty[ks][hf] = halves[i + hf].ty; g->half_tx[ks][hf] = halves[i + hf].tx; g->half_pl[ks][hf] = halves[i + hf].v;
      }
      ++ks;
    }
    if (halves.size() % 2) {
      const Half &d = halves[halves.size() - 2], &l = halves.back();
      g->half_ty[ks][0] = -1 - d.ty; g->half_tx[ks][0] = d.tx; g->half_pl[ks][0] = d.v;     // dummy (encoded ty < 0)
      g->half_ty[ks][1] = l.ty; g->half_tx[ks][1] = l.tx; g->half_pl[ks][1] = l.v;
      ++ks;
    }
  } else {
    // plane pairs (2j, 2j+1) of one parity at one tap: LBO = plane pitch
    for (int par = 0; par < 4; ++par) {
      const int py = par >> 1, px = par & 1;
      for (int ty = (py ? 0 : 1); ty < 2; ++ty)
        for (int tx = (px ? 0 : 1); tx < 2; ++tx)
          for (int j = 0; j < P / 2; ++j) {
            for (int hf = 0; hf < 2; ++hf) { g->half_ty[ks][hf] = ty; g->half_tx[ks][hf] = tx; g->half_pl[ks][hf] = par * P + 2 * j + hf; }
            ++ks;
          }
    }
  }
  if (ks > kTcMaxKSteps) return 1;
  g->ksteps = ks;
  g->bgroup = ks;
  if ((size_t)ks * 32 * nc > 80 * 1024) return 1;      // the plan keeps these weights resident
  return 0;
}

static inline uint16_t f2h(float f) {
  const __half_raw r = static_cast<__half_raw>(__float2half_rn(f));   // host-callable, round to nearest even
  return r.x;
}
static inline uint16_t f2bf(float f) {
  uint32_t u;
  std::memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return 0x7fc0;
  u += 0x7fffu + ((u >> 16) & 1u);   // round to nearest even
  return (uint16_t)(u >> 16);
}

// Banded weights of the tensor-core stem.  GEMM row (y, g) holds pixels 8g..8g+7 of image row y as its
// 8 input channels; GEMM column plane*64 + jo*8 + c8 is output pixel 8g+jo, channel plane*8+c8.  Tap
// (dy, dgx) of the 3x3 conv over the GROUP grid connects input pixel 8(g+dgx-1)+ji to output pixel 8g+jo
// with the original filter tap dx = 8(dgx-1) + ji - jo + 1 when that lies in 0..2, else zero.
void tc_stem_group_weights(const float *w, int cout, std::vector<float> *out) {
  const int cols = 8 * cout;
  out->assign((size_t)9 * 8 * cols, 0.f);
  for (int dy = 0; dy < 3; ++dy)
    for (int dgx = 0; dgx < 3; ++dgx)
      for (int ji = 0; ji < 8; ++ji)
        for (int col = 0; col < cols; ++col) {
          const int jo = (col >> 3) & 7, c = (col >> 6) * 8 + (col & 7);
          const int dx = 8 * (dgx - 1) + ji - jo + 1;
          if (dx < 0 || dx > 2) continue;
          (*out)[(((size_t)dy * 3 + dgx) * 8 + ji) * cols + col] = w[((size_t)dy * 3 + dx) * cout + c];
        }
}

void tc_rowpair_weights(const float *w, int cin, int cout, std::vector<float> *out) {
  const int cols = 2 * cout;
  out->assign((size_t)4 * 3 * cin * cols, 0.f);
  for (int a = 0; a < 4; ++a)
    for (int par = 0; par < 2; ++par) {
      const int dy = a - par;
      if (dy < 0 || dy > 2) continue;
      for (int b = 0; b < 3; ++b)
        for (int ci = 0; ci < cin; ++ci)
          for (int co = 0; co < cout; ++co)
            (*out)[(((size_t)a * 3 + b) * cin + ci) * cols + par * cout + co] = w[(((size_t)dy * 3 + b) * cin + ci) * cout + co];
    }
}

int tc_make_geometry_split(int kh, int kw, int cin, int cout, int ups, TcGeometry *g, int pad_top, int pad_left) {
  if (tc_make_geometry(kh, kw, 2 * cin, 2 * cout, ups, g, pad_top, pad_left)) return 1;
  g->split = 1; g->cin_l = cin; g->cout_l = cout; g->wscale = 1.f;
  return 0;
}

float tc_split_weight_scale(const float *w, size_t count) {
  float mx = 0.f;
  for (size_t i = 0; i < count; ++i) mx = std::max(mx, std::fabs(w[i]));
  if (!(mx > 0.f) || !std::isfinite(mx)) return 1.f;
  int e;
  std::frexp(mx, &e);                       // mx = f * 2^e, f in [0.5, 1)
  return std::ldexp(1.f, 4 - e);            // mx * scale in [2^3, 2^4)
}

// Value of the (tap-folded) filter for half `hf` of k-step `s`, LOGICAL input channel ci and LOGICAL GEMM column col
static double tc_folded_weight(const TcGeometry &g, const float *w, int s, int hf, int ci, int col) {
  const int cin = g.split ? g.cin_l : g.cin, cout = g.split ? g.cout_l : g.cout;
  const int pt = g.pt, pl = g.pl;
  const int dy = g.half_ty[s][hf] + g.dy_min, dx = g.half_tx[s][hf] + g.dx_min;
  if (!g.ups) {
    const int a = dy + pt, b = dx + pl;
    return w[(((size_t)a * g.kw + b) * cin + ci) * cout + col];
  }
  const int par = col / cout, co = col % cout;
  const int py = par >> 1, px = par & 1;
  double val = 0.0;
  float valf = 0.f;     // the 16-bit modes sum in float, as the device-side packer does
  for (int a = 0; a < g.kh; ++a)
    for (int b = 0; b < g.kw; ++b)
      if (floordiv2(py + a - pt) == dy && floordiv2(px + b - pl) == dx) {
        val += w[(((size_t)a * g.kw + b) * cin + ci) * cout + co];
        valf += w[(((size_t)a * g.kw + b) * cin + ci) * cout + co];
      }
  return g.split ? val : (double)valf;
}

// s2d geometry (tc_make_geometry_s2d): weight of K half (s, hf), input channel kk of its plane, output column col, summed
// from the up-conv's FORWARD kernel w[2][2][g.cout][g.cin]
__host__ __device__ __forceinline__ float s2d_weight(const TcGeometry &g, const float *__restrict__ w, int s, int hf, int kk, int col) {
  const int P = g.cin / 8, v = g.half_pl[s][hf], par = v / P, cz = (v - par * P) * 8 + kk;
  const int py = par >> 1, px = par & 1;
  const int a = py ? (g.half_ty[s][hf] ? 1 : -1) : 0, b = px ? (g.half_tx[s][hf] ? 1 : -1) : 0;
  float val = 0.f;
  for (int ky = 0; ky < 2; ++ky) {
    if ((a == -1 && ky != 1) || (a == 1 && ky != 0)) continue;
    for (int kx = 0; kx < 2; ++kx) {
      if ((b == -1 && kx != 1) || (b == 1 && kx != 0)) continue;
      val += w[(((long long)ky * 2 + kx) * g.cout + col) * g.cin + cz];
    }
  }
  return val;
}

void tc_pack_weights(const TcGeometry &g, const float *w, std::vector<uint16_t> *out, int fp16) {
  const size_t per_step = (size_t)2 * g.n_cols * 8;
  out->assign((size_t)g.n_tiles_n * g.cin_chunks * g.ksteps * per_step, 0);
  for (int nt = 0; nt < g.n_tiles_n; ++nt)
    for (int ch = 0; ch < g.cin_chunks; ++ch)
      for (int s = 0; s < g.ksteps; ++s)
        for (int hf = 0; hf < 2; ++hf) {
          if (g.half_ty[s][hf] < 0) continue;   // dummy half: zero weights
          const int plane = ch * g.planes_per_chunk + g.half_pl[s][hf];
          for (int n = 0; n < g.n_cols; ++n) {
            const int col = nt * g.n_cols + n;
            if (col >= g.cols_valid) continue;
            for (int kk = 0; kk < 8; ++kk) {
              uint16_t bits;
              if (g.s2d) {
                const float val = s2d_weight(g, w, s, hf, kk, col);
                bits = fp16 ? f2h(val) : f2bf(val);
              } else if (!g.split) {
                const float val = (float)tc_folded_weight(g, w, s, hf, plane * 8 + kk, col);
                bits = fp16 ? f2h(val) : f2bf(val);
              } else {
                // physical plane 2p = hi, 2p+1 = lo' of logical plane p; physical columns 16j.. = [main 8 | corr 8]
                const int part_k = plane & 1, part_n = (col >> 3) & 1;
                const float v = (float)(tc_folded_weight(g, w, s, hf, (plane >> 1) * 8 + kk, (col >> 4) * 8 + (col & 7)) * (double)g.wscale);
                const uint16_t hi = f2h(v);
                __half_raw hr; hr.x = hi;
                const float lo = (v - __half2float(__half(hr))) * 2048.f;
                if (part_k == 0) bits = part_n == 0 ? hi : f2h(lo);          // hi x (W_hi | W_lo')
                else bits = part_n == 0 ? (uint16_t)0 : hi;                  // lo' x (0 | W_hi)
              }
              (*out)[(((size_t)(nt * g.cin_chunks + ch) * g.ksteps + s) * 2 + hf) * g.n_cols * 8 + (size_t)n * 8 + kk] = bits;
            }
          }
        }
}

// Device-side twin of tc_pack_weights (training: the weights change every step).
// transposed=1 packs the data-gradient operator: W'[a][b][ci'][co'] = w[kh-1-a][kw-1-b][co'][ci']
// where w is the forward kernel [kh][kw][g.cout][g.cin] (g describes the dgrad conv).
__global__ void tc_pack_kernel(TcGeometry g, const float *__restrict__ w, int transposed,
                               __nv_bfloat16 *__restrict__ out, long long total) {
  const int pt = g.pt, pl = g.pl;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int kk = (int)(i & 7);
    long long t = i >> 3;
    const int n = (int)(t % g.n_cols); t /= g.n_cols;
    const int hf = (int)(t & 1); t >>= 1;
    const int s = (int)(t % g.ksteps); t /= g.ksteps;
    const int ch = (int)(t % g.cin_chunks);
    const int nt = (int)(t / g.cin_chunks);
    float val = 0.f;
    const int col = nt * g.n_cols + n;
    if (g.half_ty[s][hf] >= 0 && col < g.cols_valid) {
      const int dy = g.half_ty[s][hf] + g.dy_min, dx = g.half_tx[s][hf] + g.dx_min;
      const int ci = (ch * g.planes_per_chunk + g.half_pl[s][hf]) * 8 + kk;
      if (g.s2d) {
        val = s2d_weight(g, w, s, hf, kk, col);
      } else if (g.rows2) {
        // banded 4x3 row-pair filter built on the fly from the 3x3 kernel (tc_rowpair_weights)
        const int cr = g.cout >> 1, par = col / cr, co = col - par * cr;
        const int a = dy + pt - par, b = dx + pl;
        if (a >= 0 && a <= 2) {
          if (!transposed) val = w[(((long long)a * 3 + b) * g.cin + ci) * cr + co];
          else val = w[(((long long)(2 - a) * 3 + (2 - b)) * cr + co) * g.cin + ci];
        }
      } else if (!g.ups) {
        const int a = dy + pt, b = dx + pl;
        if (!transposed) val = w[(((long long)a * g.kw + b) * g.cin + ci) * g.cout + col];
        else val = w[(((long long)(g.kh - 1 - a) * g.kw + (g.kw - 1 - b)) * g.cout + col) * g.cin + ci];
      } else {
        const int par = col / g.cout, co = col % g.cout;
        const int py = par >> 1, px = par & 1;
        for (int a = 0; a < g.kh; ++a)
          for (int b = 0; b < g.kw; ++b) {
            const int ya = py + a - pt, xb = px + b - pl;
            const int fy = ya >= 0 ? ya / 2 : -((-ya + 1) / 2), fx = xb >= 0 ? xb / 2 : -((-xb + 1) / 2);
            if (fy == dy && fx == dx) val += w[(((long long)a * g.kw + b) * g.cin + ci) * g.cout + co];
          }
      }
    }
    out[i] = __float2bfloat16(val);
  }
}

// One launch packs every tensor-core weight image of the net (training: after each optimizer step)
__global__ void tc_pack_all_kernel(const TcPackJob *__restrict__ jobs) {
  const TcPackJob &j = jobs[blockIdx.y];
  const TcGeometry &g = j.g;
  const int pt = g.pt, pl = g.pl;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < j.total;
       i += (long long)gridDim.x * blockDim.x) {
    const int kk = (int)(i & 7);
    long long t = i >> 3;
    const int n = (int)(t % g.n_cols); t /= g.n_cols;
    const int hf = (int)(t & 1); t >>= 1;
    const int s = (int)(t % g.ksteps); t /= g.ksteps;
    const int ch = (int)(t % g.cin_chunks);
    const int nt = (int)(t / g.cin_chunks);
    float val = 0.f;
    const int col = nt * g.n_cols + n;
    if (g.half_ty[s][hf] >= 0 && col < g.cols_valid) {
      const int dy = g.half_ty[s][hf] + g.dy_min, dx = g.half_tx[s][hf] + g.dx_min;
      const int ci = (ch * g.planes_per_chunk + g.half_pl[s][hf]) * 8 + kk;
      if (g.s2d) {
        val = s2d_weight(g, j.w, s, hf, kk, col);
      } else if (g.rows2) {
        const int cr = g.cout >> 1, par = col / cr, co = col - par * cr;
        const int a = dy + pt - par, b = dx + pl;
        if (a >= 0 && a <= 2) {
          if (!j.transposed) val = j.w[(((long long)a * 3 + b) * g.cin + ci) * cr + co];
          else val = j.w[(((long long)(2 - a) * 3 + (2 - b)) * cr + co) * g.cin + ci];
        }
      } else if (!g.ups) {
        const int a = dy + pt, b = dx + pl;
        if (!j.transposed) val = j.w[(((long long)a * g.kw + b) * g.cin + ci) * g.cout + col];
        else val = j.w[(((long long)(g.kh - 1 - a) * g.kw + (g.kw - 1 - b)) * g.cout + col) * g.cin + ci];
      } else {
        const int par = col / g.cout, co = col % g.cout;
        const int py = par >> 1, px = par & 1;
        for (int a = 0; a < g.kh; ++a)
          for (int b = 0; b < g.kw; ++b) {
            const int ya = py + a - pt, xb = px + b - pl;
            const int fy = ya >= 0 ? ya / 2 : -((-ya + 1) / 2), fx = xb >= 0 ? xb / 2 : -((-xb + 1) / 2);
            if (fy == dy && fx == dx) val += j.w[(((long long)a * g.kw + b) * g.cin + ci) * g.cout + co];
          }
      }
    }
    j.out[i] = __float2bfloat16(val);
  }
}

int tc_pack_all_device(const TcPackJob *jobs_dev, int n_jobs, cudaStream_t st) {
  if (n_jobs <= 0) return 0;
  dim3 grid(32, n_jobs);
  tc_pack_all_kernel<<<grid, 256, 0, st>>>(jobs_dev);
  OCTSEG_CUDA(cudaGetLastError());
  return 0;
}

int tc_pack_weights_device(const TcGeometry &g, const float *w_dev, int transposed, __nv_bfloat16 *out,
                           cudaStream_t st) {
  const long long total = (long long)g.n_tiles_n * g.cin_chunks * g.ksteps * 2 * g.n_cols * 8;
  unsigned grid = (unsigned)std::min<long long>((total + 255) / 256, 148 * 4);
  tc_pack_kernel<<<grid, 256, 0, st>>>(g, w_dev, transposed, out, total);
  OCTSEG_CUDA(cudaGetLastError());
  return 0;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void *ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(ptr);
  }
  return fn;
}

int tc_fill_params(const TcGeometry &g, int n, int h, int w, TcConvParams *pp, size_t *smem_bytes) {
  TcConvParams &p = *pp;
  std::memset(&p, 0, sizeof(p));
  p.n = n; p.h = h; p.w = w;
  p.n_tiles_n = g.n_tiles_n;
  p.cin_chunks = g.cin_chunks;
  p.planes_per_chunk = g.planes_per_chunk;
  p.ksteps = g.ksteps;
  p.n_cols = g.n_cols;
  p.cols_valid = g.cols_valid;
  p.pad_y = -g.dy_min; p.pad_x = -g.dx_min;
  // ---- weights: resident in smem when the whole packed image of the layer is small
  const size_t w_total = (size_t)g.cin_chunks * g.ksteps * 32u * g.n_cols;
  p.b_resident = (g.n_tiles_n == 1 && w_total <= 80 * 1024) ? 1 : 0;
  size_t b_bytes_total;
  if (p.b_resident) {
    p.bgroup = g.ksteps; p.b_stages = 1; p.b_stage_bytes = (uint32_t)w_total;
    b_bytes_total = (w_total + 127) & ~(size_t)127;
  } else {
    p.bgroup = g.bgroup;
    p.b_stage_bytes = (uint32_t)g.bgroup * 32u * (uint32_t)g.n_cols;
    p.b_stages = 4;
    while (p.b_stages > 2 && (size_t)p.b_stages * p.b_stage_bytes > 96 * 1024) --p.b_stages;
    b_bytes_total = (size_t)p.b_stages * ((p.b_stage_bytes + 127u) & ~127u);
  }
  // ---- super-tile: mt_x x mt_y M-tiles of 8 px x 16 rows share one TMA halo box
  size_t epi_bytes = ((size_t)g.cout * (2 + kTcMaxClasses) + kTcMaxClasses) * sizeof(float);
  if (g.stem_groups) {     // + per-epilogue-warp store staging (32 rows x 128 B)
    epi_bytes = (epi_bytes + 127) & ~(size_t)127;
    p.stage_off = (uint32_t)epi_bytes;
    epi_bytes += (size_t)kTcEpiWarps * 4096;
  }
  const size_t budget = 208 * 1024 - b_bytes_total - sizeof(TcBarriers) - epi_bytes - 1024;
  const int max_mt = std::max(1, 256 / g.n_cols);          // 2 accumulator stages in 512 TMEM columns
  const int tile_h = kTcTileH * (g.rows2 ? 2 : 1);          // image rows covered by one M-tile
  p.row_mul = g.rows2 ? 2 : 1;
  p.scale_mod = g.split ? (g.rows2 ? g.cout_l / 2 : g.cout_l) : (g.rows2 ? g.cout / 2 : g.cout);
  p.split = g.split;
  p.scale_mul = g.split ? 1.f / g.wscale : 1.f;
  if (g.split && g.stem_groups) { set_error("tc plan: split mode has no stem-group variant"); return 1; }
  static const int cand[][2] = {{8, 2}, {4, 2}, {8, 1}, {4, 1}, {2, 2}, {2, 1}, {1, 2}, {1, 1}};
  int best_x = 1, best_y = 1;
  for (auto &c : cand) {
    const int mx = c[0], my = c[1];
    if (mx * my > max_mt) continue;
    if (mx * kTcTileW > std::max(w, kTcTileW) || my * tile_h > std::max(h, tile_h)) continue;
    const size_t stage = (size_t)g.planes_per_chunk * (mx * kTcTileW + g.box_w) * (my * tile_h + g.box_h) * 16;
    // three stages of <= 48 KB, or two stages of <= 80 KB (tall row-pair tiles, 8-plane chunks)
    if (!((stage <= 48 * 1024 && 3 * stage <= budget) || (stage <= 80 * 1024 && 2 * stage <= budget))) continue;
    const long long tiles = (long long)n * ((w + mx * kTcTileW - 1) / (mx * kTcTileW)) *
                            ((h + my * tile_h - 1) / (my * tile_h)) * g.n_tiles_n;
    // keep every SM busy, but prefer >= 2 M-tiles per super-tile while one full round remains: a lone
    // M-tile is issued by a single thread (~2x slower than the tensor pipe, see the issuer comment)
    if (tiles < 148 && mx * my > 1) continue;
    best_x = mx; best_y = my;
    break;
  }
  p.mt_x = best_x; p.mt_y = best_y;
  p.mt_x_log2 = best_x == 8 ? 3 : best_x == 4 ? 2 : best_x == 2 ? 1 : 0;
  p.shuffle_pairs = (g.ups && (g.n_cols % (2 * g.cout)) == 0 && (g.cols_valid % (2 * g.cout)) == 0) ? 1 : 0;
  p.box_w = best_x * kTcTileW + g.box_w;
  p.box_h = best_y * tile_h + g.box_h;
  p.tiles_x = (w + best_x * kTcTileW - 1) / (best_x * kTcTileW);
  p.tiles_y = (h + best_y * tile_h - 1) / (best_y * tile_h);
  p.num_tiles = n * p.tiles_x * p.tiles_y * p.n_tiles_n;
  const uint32_t pitch = (uint32_t)p.box_w * 16u;
  const uint32_t plane = pitch * (uint32_t)p.box_h;
  p.a_stage_bytes = plane * (uint32_t)g.planes_per_chunk;
  p.a_tx_bytes = p.a_stage_bytes;
  p.s2d = g.s2d;
  if (g.s2d) {
    p.s2d_part_bytes = (plane * (uint32_t)(g.cin / 8) + 127u) & ~127u;     // TMA destinations are 128-byte aligned
    p.a_stage_bytes = 4u * p.s2d_part_bytes;
  }
  for (int s = 0; s < g.ksteps; ++s) {
    uint32_t off[2];
    for (int hf = 0; hf < 2; ++hf) {
      int ty = g.half_ty[s][hf];
      if (ty < 0) ty = -1 - ty;   // dummy half: valid in-tile address, zero weights
      if (g.s2d) {
        const uint32_t P = (uint32_t)g.cin / 8u, v = (uint32_t)g.half_pl[s][hf];
        off[hf] = (v / P) * p.s2d_part_bytes + (v % P) * plane + (uint32_t)ty * pitch + (uint32_t)g.half_tx[s][hf] * 16u;
      } else
      off[hf] = (uint32_t)g.half_pl[s][hf] * plane + (uint32_t)ty * pitch + (uint32_t)g.half_tx[s][hf] * 16u;
    }
    if (off[1] <= off[0]) { set_error("tc plan: non-positive LBO"); return 1; }
    p.a_off[s] = off[0];
    p.a_lbo[s] = off[1] - off[0];
    p.a_desc_lo[s] = (off[0] >> 4) | (((off[1] - off[0]) >> 4) << 16);
    if (p.a_lbo[s] >= (1u << 18) || pitch >= (1u << 18)) { set_error("tc plan: descriptor stride overflow"); return 1; }
  }
  const uint32_t a_stride = (p.a_stage_bytes + 127u) & ~127u;
  p.a_stages = (int)std::min<size_t>(kMaxStages, budget / a_stride);
  if (p.a_stages < 1) { set_error("tc plan: smem budget exceeded"); return 1; }
  // beyond ~96 KB in flight per SM more stages buy nothing
  while (p.a_stages > 3 && (size_t)(p.a_stages - 1) * a_stride >= 96 * 1024) --p.a_stages;
  *smem_bytes = (size_t)p.a_stages * a_stride + b_bytes_total + sizeof(TcBarriers) + epi_bytes + 1024;
  if (*smem_bytes > 227 * 1024) { set_error("tc plan: smem budget exceeded"); return 1; }
  p.mode = g.ups ? 1 : 0;
  p.cout = g.cout;
  return 0;
}

// 4-D tensor map over a dense blocked bf16 tensor [n][planes][h][w][8], declared in 8-byte elements
// (2 per pixel-plane vector); box = box_w px x box_h rows x box_planes planes of one image.
int tc_encode_map_4d(const void *base, int w, int h, int planes, int n, int box_w, int box_h, int box_planes,
                     CUtensorMap *out) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)"); return 1; }
  cuuint64_t dims[4] = {(cuuint64_t)w * 2, (cuuint64_t)h, (cuuint64_t)planes, (cuuint64_t)n};
  cuuint64_t strides[3] = {(cuuint64_t)w * 16, (cuuint64_t)h * w * 16, (cuuint64_t)planes * h * w * 16};
  cuuint32_t box[4] = {(cuuint32_t)box_w * 2, (cuuint32_t)box_h, (cuuint32_t)box_planes, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_UINT64, 4, const_cast<void *>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed: " + std::to_string((int)r)); return 1; }
  return 0;
}

int tc_make_plan(const TcGeometry &g, const __nv_bfloat16 *in, int n, int h, int w,
                 const __nv_bfloat16 *wpack_dev, const TcEpilogue &epi, int *status_dev, TcPlan *plan) {
  TcConvParams &p = plan->p;
  if (tc_fill_params(g, n, h, w, &p, &plan->smem_bytes)) return 1;
  p.relu = epi.relu;
  p.fp16 = epi.fp16;
  p.static_weights = epi.static_weights;
  p.overflow = epi.overflow;
  p.stats = epi.stats;
  if (epi.stats && (epi.head_w || epi.pool_out || g.split || g.stem_groups)) { set_error("tc plan: batch statistics belong to the plain training epilogue"); return 1; }
  if (g.split && !epi.fp16) { set_error("tc plan: split mode stores fp16 pairs"); return 1; }
  p.scale = epi.scale; p.shift = epi.shift;
  p.out = epi.out.ptr; p.out_img_stride = epi.out.img_stride; p.out_h = epi.out.h; p.out_w = epi.out.w;
  p.pool_out = epi.pool_out; p.pool_img_stride = epi.pool_img_stride; p.pool_sum = epi.pool_sum;
  if (epi.pool_sum && (!epi.pool_out || epi.head_w || epi.stats || g.split || g.rows2 || g.stem_groups)) {
    set_error("tc plan: sum-pool epilogue not applicable");
    return 1;
  }
  if (epi.head_w) {
    if (g.ups || g.n_tiles_n != 1 || !tc_head_fusable(epi.head_k)) { set_error("tc plan: head fusion not applicable"); return 1; }
    p.mode = 2;
    p.head_w = epi.head_w; p.head_b = epi.head_b; p.head_k = epi.head_k;
    p.probs = epi.probs; p.labels = epi.labels;
    p.out_h = h; p.out_w = w;
  }
  if (g.rows2) {
    const int cr = (g.split ? g.cout_l : g.cout) / 2;
    if (g.ups || g.n_tiles_n != 1 || (cr != 8 && cr != 16) || (epi.head_w && cr != 8)) { set_error("tc plan: row-pair mode not applicable"); return 1; }
  }
  if (g.stem_groups) {
    if (g.ups || epi.head_w || epi.pool_out || g.n_cols % 64 || g.cols_valid % 64) { set_error("tc plan: stem-group epilogue not applicable"); return 1; }
    p.mode = 3;
  }
  if (epi.pool_out && (g.ups || (h & 1) || (w & 1))) { set_error("tc plan: pool fusion not applicable"); return 1; }
  p.wpack = wpack_dev;
  p.status = status_dev;

  PFN_encodeTiled enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)"); return 1; }
  if (g.s2d) {
    // `in` = dz on the HIGH-res grid [n][cz/8][2h][2w][8]; (h, w) = the low-res GEMM-row grid.  One 5-D map per pixel
    // parity: (8-byte half, X stride 32 B, Y stride 2 rows, plane, image), base shifted by the parity.
    if (epi.head_w || epi.pool_out || epi.stats) { set_error("tc plan: s2d data gradient takes the plain epilogue"); return 1; }
    const int P = g.cin / 8, H = 2 * h, W = 2 * w;
    CUtensorMap maps[4];
    for (int par = 0; par < 4; ++par) {
      const int py = par >> 1, px = par & 1;
      const uint8_t *base = reinterpret_cast<const uint8_t *>(in) + ((size_t)py * W + px) * 16;
      cuuint64_t dims5[5] = {2, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)P, (cuuint64_t)n};
      cuuint64_t str5[4] = {32, (cuuint64_t)2 * W * 16, (cuuint64_t)H * W * 16, (cuuint64_t)P * H * W * 16};
      cuuint32_t box5[5] = {2, (cuuint32_t)p.box_w, (cuuint32_t)p.box_h, (cuuint32_t)P, 1};
      cuuint32_t es5[5] = {1, 1, 1, 1, 1};
      CUresult r5 = enc(&maps[par], CU_TENSOR_MAP_DATA_TYPE_UINT64, 5, const_cast<uint8_t *>(base), dims5, str5, box5, es5,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r5 != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (s2d) failed: " + std::to_string((int)r5)); return 1; }
    }
    // Re-planning (new batch / image size): kernels of the previous step may still be queued with the old maps, and a tensor
    // map that is rewritten in place in global memory needs a tensormap-proxy fence in every CTA that uses it afterwards.
    // A fresh buffer per plan avoids both: the old one stays untouched until the plan is released.
    if (plan->dev_maps) { plan->retired.push_back(plan->dev_maps); plan->dev_maps = nullptr; }
    OCTSEG_CUDA(cudaMalloc(&plan->dev_maps, sizeof(maps)));
    OCTSEG_CUDA(cudaMemcpy(plan->dev_maps, maps, sizeof(maps), cudaMemcpyHostToDevice));
    p.s2d_maps = reinterpret_cast<const CUtensorMap *>(plan->dev_maps);
  }
  const int cg = g.s2d ? 4 * (g.cin / 8) : g.cin / 8;
  // declared as 8-byte elements (2 per pixel-plane vector) so that one box row may span up to 128 px
  cuuint64_t dims[4] = {(cuuint64_t)w * 2, (cuuint64_t)h, (cuuint64_t)cg, (cuuint64_t)n};
  cuuint64_t strides[3] = {(cuuint64_t)w * 16, (cuuint64_t)h * w * 16, (cuuint64_t)cg * h * w * 16};
  cuuint32_t box[4] = {(cuuint32_t)p.box_w * 2, (cuuint32_t)p.box_h, (cuuint32_t)g.planes_per_chunk, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(&plan->tmap, CU_TENSOR_MAP_DATA_TYPE_UINT64, 4, const_cast<__nv_bfloat16 *>(in), dims,
                   strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed: " + std::to_string((int)r)); return 1; }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  plan->grid = std::min(p.num_tiles, sms);
  {
    // Programmatic dependent launch works best (measured, B200, default net: 1.53 -> 1.47 ms per 64 B-scans)
    // when dependents are signalled after this CTA's last tile request and when two CTAs can never share an
    // SM: the dynamic smem request is padded to 116 KB so a dependent CTA only starts on an SM whose CTA
    // of the previous kernel has exited.  OCTSEG_PDL_LATE=0 / OCTSEG_TC_MIN_SMEM_KB=n override (experiments).
    const char *e = std::getenv("OCTSEG_PDL_LATE");
    p.pdl_late = (e && e[0] == '0') ? 0 : 1;
    const char *m = std::getenv("OCTSEG_TC_MIN_SMEM_KB");
    plan->smem_bytes = std::max(plan->smem_bytes, (size_t)(m ? std::atoi(m) : 116) * 1024);
    plan->smem_bytes = std::min(plan->smem_bytes, (size_t)227 * 1024);
  }
  plan->valid = true;
  return 0;
}

void tc_release_plan(TcPlan *plan) {
  if (plan->dev_maps) cudaFree(plan->dev_maps);
  for (void *q : plan->retired) cudaFree(q);
  plan->retired.clear();
  plan->dev_maps = nullptr;
  plan->valid = false;
}

template <int HK, int EPI, int ST = 0>
static int tc_launch_k(const TcPlan &plan, cudaStream_t st) {
  // the attribute is per device (one process may drive several GPUs, one host thread each)
  static PerDeviceOnce attr_set;
  if (const int dev = attr_set.pending(); dev >= 0) {
    OCTSEG_CUDA(cudaFuncSetAttribute(conv_tc_kernel<HK, EPI, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set.mark(dev);
  }
  static const bool pdl = []() { const char *e = std::getenv("OCTSEG_NO_PDL"); return !(e && e[0] == '1'); }();
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(plan.grid); cfg.blockDim = dim3(kTcThreads);
  cfg.dynamicSmemBytes = plan.smem_bytes; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
  OCTSEG_CUDA(cudaLaunchKernelEx(&cfg, conv_tc_kernel<HK, EPI, ST>, plan.tmap, plan.p));
  return 0;
}

bool tc_head_fusable(int num_classes) { return num_classes >= 2 && num_classes <= 8; }

// epilogue specialisation of a plan (see conv_tc_kernel)
static int tc_epi_kind(const TcConvParams &p) {
  const bool one_ntile = p.n_tiles_n == 1;
  if (p.pool_sum) return 9;                                         // generic chunk loop storing 2x2 sums only
  if (p.split) return 7;
  if (p.row_mul == 2) return p.scale_mod == 8 ? 5 : 6;
  if (p.mode == 3) return 3;
  if (p.mode == 1) return p.shuffle_pairs ? 4 : 0;
  if (one_ntile && p.cols_valid == 8) return 1;                     // mode 0 or 2 (fused head)
  if (one_ntile && p.cols_valid == 16 && p.mode == 0) return 2;
  return 0;
}

template <int HK>
static int tc_launch_head(const TcPlan &plan, cudaStream_t st) {
  const int kind = tc_epi_kind(plan.p);
  if (kind == 7) return tc_launch_k<HK, 7>(plan, st);
  if (kind == 5) return tc_launch_k<HK, 5>(plan, st);
  return kind == 1 ? tc_launch_k<HK, 1>(plan, st) : tc_launch_k<HK, 0>(plan, st);
}

int tc_launch(const TcPlan &plan, cudaStream_t st) {
  if (plan.p.stats) {
    // training forward: the narrow-layer epilogues (<= 16 channels: register accumulators) exist with the statistics
    // compiled in; everything else takes the generic statistics epilogue
    if (plan.p.mode == 2) { set_error("tc launch: batch statistics with a fused head"); return 1; }
    switch (tc_epi_kind(plan.p)) {
      case 1: return tc_launch_k<0, 1, 1>(plan, st);
      case 2: return tc_launch_k<0, 2, 1>(plan, st);
      case 4: if (plan.p.scale_mod <= 16) return tc_launch_k<0, 4, 1>(plan, st); break;
      case 5: return tc_launch_k<0, 5, 1>(plan, st);
      case 6: return tc_launch_k<0, 6, 1>(plan, st);
    }
    return tc_launch_k<0, 8>(plan, st);
  }
  if (plan.p.pool_sum) {
    if (plan.p.mode != 0 || plan.p.stats || plan.p.split || plan.p.row_mul != 1) { set_error("tc launch: sum-pool epilogue not applicable"); return 1; }
    return tc_launch_k<0, 9>(plan, st);
  }
  if (plan.p.mode != 2) {
    switch (tc_epi_kind(plan.p)) {
      case 1: return tc_launch_k<0, 1>(plan, st);
      case 2: return tc_launch_k<0, 2>(plan, st);
      case 3: return tc_launch_k<0, 3>(plan, st);
      case 4: return tc_launch_k<0, 4>(plan, st);
      case 5: return tc_launch_k<0, 5>(plan, st);
      case 6: return tc_launch_k<0, 6>(plan, st);
      case 7: return tc_launch_k<0, 7>(plan, st);
      case 8: return tc_launch_k<0, 8>(plan, st);
    }
    return tc_launch_k<0, 0>(plan, st);
  }
  switch (plan.p.head_k) {
    case 2: return tc_launch_head<2>(plan, st);
    case 3: return tc_launch_head<3>(plan, st);
    case 4: return tc_launch_head<4>(plan, st);
    case 5: return tc_launch_head<5>(plan, st);
    case 6: return tc_launch_head<6>(plan, st);
    case 7: return tc_launch_head<7>(plan, st);
    case 8: return tc_launch_head<8>(plan, st);
  }
  set_error("tc launch: fused head supports 2..8 classes");
  return 1;
}

}  // namespace octseg
