// Weight gradient on tensor cores (bf16 mode):
//   dW[tap][ci][co] = sum over pixels of a[p + tap][ci] * dz[p][co]
// GEMM view: M = (tap, ci) , N = co, K = pixels.  The narrow default net has Cin*taps = 72..1152
// and Cout = 8..128, so the natural tile is the warp-level m16n8k16 MMA (an M=64/128 tcgen05
// tile would be >80 % padding for Cin = 8/16); both operands come out of the blocked
// [plane][row][px][8ch] smem tile with ldmatrix.trans (a 16 B smem row = 8 channels of one
// pixel, so the 8x8 transpose turns "pixel-major" storage into the K=pixel fragments), and
// every filter tap is just a different row address -- the same no-im2col idea as conv_tc.
// CTAs are persistent over pixel tiles and keep their dW block in registers; one smem
// reduction + one global atomic per element per CTA at the end.
//
// Kernels in this file, in the order launch_wgrad_mma tries them:
//   wgrad_rows_kernel    Cin = 8 / 16 / 32 (3x3 and 2x2) and the 64-channel up-conv: warps walk the input rows of a band,
//                        one ldmatrix per (dx, plane) pair feeds all kh taps; reads the up-conv's LOW-res input directly
//   wgrad_deep_kernel    Cin >= 32 otherwise (64 -> 32 3x3 in the default net): MP x NP warp tiling of the output block
//   wgrad_mma_tma_kernel / wgrad_mma_kernel   the first TMA-pipelined / staged kernels, kept for shapes the others refuse
// (layers with Cin, Cout >= 64 go to the tcgen05 kernel in wgrad_tc.cu before this file is reached)
#include <cuda_bf16.h>

#include <algorithm>
#include <cstdlib>

#include <cuda.h>

#include "conv_tc.cuh"
#include "train_kernels.cuh"

namespace octseg {

constexpr int kWmTW = 32, kWmTH = 8;      // pixel tile: 32 px (2 k-steps) x 8 rows (1 row per warp)
constexpr int kWmMaxGroups = 24;          // (tap, plane) groups per CTA block -> 12 m16 tiles
constexpr int kWmMaxMT = 13;              // + 1 for the "ones" group (bias gradient)
constexpr int kWmNB = 2;                  // n8 tiles per CTA block

struct WgParams {
  const __nv_bfloat16 *a; long long a_img_stride; int a_h, a_w;
  const __nv_bfloat16 *dz; long long dz_img_stride; int H, W;
  int n, kh, kw, pt, pl, ups;
  int cin, cout, cin_planes, cout_planes;
  int pc, n_pchunks, n_gblocks, n_nchunks;
  int tiles_x, tiles_y, num_tiles;
  int d_planes;            // planes in the dz TMA box (min(kWmNB, cout_planes))
  float *dW, *db;
};

__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x2_trans(uint32_t addr, uint32_t (&r)[2]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

__global__ void __launch_bounds__(256) wgrad_mma_kernel(const WgParams p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int aw = kWmTW + p.kw - 1, ah = kWmTH + p.kh - 1;
  // ---- which block of the output does this CTA own?
  int by = blockIdx.y;
  const int nchunk = by % p.n_nchunks; by /= p.n_nchunks;
  const int gblock = by % p.n_gblocks;
  const int pchunk = by / p.n_gblocks;
  const int plane0 = pchunk * p.pc;
  const int pc_cur = min(p.pc, p.cin_planes - plane0);
  const int ntaps = p.kh * p.kw;
  const int g_total = ntaps * pc_cur;
  const int g0 = gblock * kWmMaxGroups;
  if (g0 >= g_total) return;
  int ng = min(kWmMaxGroups, g_total - g0);
  const bool want_db = (p.db != nullptr) && pchunk == 0 && gblock == 0;
  const int ones_group = want_db ? ng : -1;       // extra group of all-ones rows: its D row is sum(dz)
  const int ng_all = ng + (want_db ? 1 : 0);
  const int n_mt = (ng_all + 1) / 2;
  const int nb0 = nchunk * kWmNB;
  const int nb_cur = min(kWmNB, p.cout_planes - nb0);

  __nv_bfloat16 *s_a = reinterpret_cast<__nv_bfloat16 *>(smem_raw);                 // [pc][ah][aw][8]
  __nv_bfloat16 *s_d = s_a + (size_t)p.pc * ah * aw * 8;                            // [NB][TH][TW][8]
  __nv_bfloat16 *s_ones = s_d + (size_t)kWmNB * kWmTH * kWmTW * 8;                  // [8][8]
  float *s_acc = reinterpret_cast<float *>(s_ones + 64);                            // [(MaxGroups+2)][8][NB*8]
  for (int i = threadIdx.x; i < 64; i += blockDim.x) s_ones[i] = __float2bfloat16(1.0f);
  for (int i = threadIdx.x; i < (kWmMaxGroups + 2) * 8 * kWmNB * 8; i += blockDim.x) s_acc[i] = 0.f;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // per-lane ldmatrix row offsets (bytes) of every m-tile: matrix i = lane>>3, row r = lane&7
  uint32_t a_off[kWmMaxMT];
  {
    const int i = lane >> 3, r = lane & 7;
#pragma unroll
    for (int mt = 0; mt < kWmMaxMT; ++mt) {
      int g = 2 * mt + (i & 1);
      uint32_t off = 0;
      if (mt < n_mt) {
        if (g >= ng_all) g = 0;                      // dummy second half of the last m-tile
        if (g == ones_group) off = 0x80000000u | (uint32_t)(r * 16);
        else {
          const int gg = g0 + g;
          const int t = gg / pc_cur, cgl = gg - t * pc_cur;
          const int dy = t / p.kw, dx = t - dy * p.kw;
          off = (uint32_t)((((cgl * ah + dy) * aw + dx) + r + 8 * (i >> 1)) * 16);
        }
      }
      a_off[mt] = off;
    }
  }
  const uint32_t sa_base = (uint32_t)__cvta_generic_to_shared(s_a);
  const uint32_t sd_base = (uint32_t)__cvta_generic_to_shared(s_d);
  const uint32_t so_base = (uint32_t)__cvta_generic_to_shared(s_ones);
  float acc[kWmMaxMT][kWmNB][4];
#pragma unroll
  for (int mt = 0; mt < kWmMaxMT; ++mt)
#pragma unroll
    for (int nb = 0; nb < kWmNB; ++nb)
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[mt][nb][k] = 0.f;

  for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
    const int txi = tile % p.tiles_x;
    const int tyi = (tile / p.tiles_x) % p.tiles_y;
    const int img = tile / (p.tiles_x * p.tiles_y);
    const int x0 = txi * kWmTW, y0 = tyi * kWmTH;
    __syncthreads();     // previous tile fully consumed
    // ---- stage the input halo tile (virtual x2-upsampled coordinates when ups) and the dz tile
    for (int i = threadIdx.x; i < pc_cur * ah * aw; i += blockDim.x) {
      const int tx = i % aw, ty = (i / aw) % ah, cgl = i / (aw * ah);
      int vy = y0 + ty - p.pt, vx = x0 + tx - p.pl;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (vy >= 0 && vy < p.H && vx >= 0 && vx < p.W) {
        if (p.ups) { vy >>= 1; vx >>= 1; }
        v = *reinterpret_cast<const uint4 *>(p.a + (long long)img * p.a_img_stride +
                                             (((long long)(plane0 + cgl) * p.a_h + vy) * p.a_w + vx) * 8);
      }
      *reinterpret_cast<uint4 *>(s_a + (size_t)i * 8) = v;
    }
    for (int i = threadIdx.x; i < nb_cur * kWmTH * kWmTW; i += blockDim.x) {
      const int tx = i % kWmTW, ty = (i / kWmTW) % kWmTH, nbi = i / (kWmTW * kWmTH);
      const int y = y0 + ty, x = x0 + tx;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (y < p.H && x < p.W)
        v = *reinterpret_cast<const uint4 *>(p.dz + (long long)img * p.dz_img_stride +
                                             (((long long)(nb0 + nbi) * p.H + y) * p.W + x) * 8);
      *reinterpret_cast<uint4 *>(s_d + (size_t)i * 8) = v;
    }
    __syncthreads();
    // ---- this warp's row of the tile: 2 k-steps of 16 pixels
    const int y = warp;
#pragma unroll
    for (int ks = 0; ks < kWmTW / 16; ++ks) {
      uint32_t bfrag[kWmNB][2];
#pragma unroll
      for (int nb = 0; nb < kWmNB; ++nb) {
        if (nb < nb_cur) {
          const int i = (lane >> 3) & 1, r = lane & 7;
          ldmatrix_x2_trans(sd_base + (uint32_t)((((nb * kWmTH + y) * kWmTW) + ks * 16 + i * 8 + r) * 16), bfrag[nb]);
        } else { bfrag[nb][0] = 0u; bfrag[nb][1] = 0u; }
      }
      const uint32_t row_off = (uint32_t)((y * aw + ks * 16) * 16);
#pragma unroll
      for (int mt = 0; mt < kWmMaxMT; ++mt) {
        if (mt < n_mt) {
          uint32_t afrag[4];
          const uint32_t o = a_off[mt];
          const uint32_t addr = (o & 0x80000000u) ? so_base + (o & 0x7FFFFFFFu) : sa_base + o + row_off;
          ldmatrix_x4_trans(addr, afrag);
#pragma unroll
          for (int nb = 0; nb < kWmNB; ++nb) mma_bf16_16816(acc[mt][nb], afrag, bfrag[nb]);
        }
      }
    }
  }
  // ---- reduce the 8 warps' partial blocks in smem, then one global atomic per element
  __syncthreads();
#pragma unroll
  for (int mt = 0; mt < kWmMaxMT; ++mt) {
    if (mt < n_mt) {
#pragma unroll
      for (int nb = 0; nb < kWmNB; ++nb) {
        const int ci = lane >> 2, co = nb * 8 + (lane & 3) * 2;
        atomicAdd(&s_acc[((2 * mt) * 8 + ci) * (kWmNB * 8) + co], acc[mt][nb][0]);
        atomicAdd(&s_acc[((2 * mt) * 8 + ci) * (kWmNB * 8) + co + 1], acc[mt][nb][1]);
        atomicAdd(&s_acc[((2 * mt + 1) * 8 + ci) * (kWmNB * 8) + co], acc[mt][nb][2]);
        atomicAdd(&s_acc[((2 * mt + 1) * 8 + ci) * (kWmNB * 8) + co + 1], acc[mt][nb][3]);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < ng_all * 8 * nb_cur * 8; i += blockDim.x) {
    const int co = i % (nb_cur * 8);
    const int ci = (i / (nb_cur * 8)) % 8;
    const int g = i / (nb_cur * 8 * 8);
    const float v = s_acc[(g * 8 + ci) * (kWmNB * 8) + co];
    if (g == ones_group) {
      if (ci == 0) atomicAdd(&p.db[nb0 * 8 + co], v);
    } else {
      const int gg = g0 + g;
      const int t = gg / pc_cur, cgl = gg - t * pc_cur;
      atomicAdd(&p.dW[((long long)t * p.cin + (plane0 + cgl) * 8 + ci) * p.cout + nb0 * 8 + co], v);
    }
  }
}

// ----------------------------------------------------------------------------------
// TMA-pipelined variant (non-upsampled inputs): warp 0 streams (input halo tile, dz tile)
// pairs through an mbarrier ring, warps 1..8 run the MMAs; no staging instructions at all.
// ----------------------------------------------------------------------------------
constexpr int kWmSplit = 3;                        // warp groups sharing the m16 tiles of a CTA
constexpr int kWmCW = 8 * kWmSplit;                // compute warps of the TMA kernel: 8 tile rows x kWmSplit parts of the m16 tiles
constexpr int kWmLocalMT = (kWmMaxMT + kWmSplit - 1) / kWmSplit;   // m16 tiles per compute warp
constexpr int kWmTmaThreads = 32 * (1 + kWmCW);
constexpr int kWmMaxStages = 8;   // ring depth is chosen per launch: small tiles need many loads in flight

__device__ __forceinline__ uint32_t wm_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool wm_wait(uint32_t bar, uint32_t parity) {
  const long long t0 = clock64();
  for (uint32_t it = 1;; ++it) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.u32 %0, 1, 0, q;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) return true;
    if ((it & 0x3FFu) == 0 && clock64() - t0 > 4000000000ll) return false;   // never hang the GPU
  }
}

template <int NB>
__global__ void __launch_bounds__(kWmTmaThreads) wgrad_mma_tma_kernel(const __grid_constant__ CUtensorMap map_a,
                                                            const __grid_constant__ CUtensorMap map_d,
                                                            const WgParams p, int th, int n_stages, int *status) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const int aw = kWmTW + p.kw - 1, ah = th + p.kh - 1;
  int by = blockIdx.y;
  const int nchunk = by % p.n_nchunks; by /= p.n_nchunks;
  const int gblock = by % p.n_gblocks;
  const int pchunk = by / p.n_gblocks;
  const int plane0 = pchunk * p.pc;
  const int pc_cur = min(p.pc, p.cin_planes - plane0);
  const int g_total = p.kh * p.kw * pc_cur;
  const int g0 = gblock * kWmMaxGroups;
  if (g0 >= g_total) return;
  const int ng = min(kWmMaxGroups, g_total - g0);
  const bool want_db = (p.db != nullptr) && pchunk == 0 && gblock == 0;
  const int ones_group = want_db ? ng : -1;
  const int ng_all = ng + (want_db ? 1 : 0);
  const int n_mt = (ng_all + 1) / 2;
  const int nb0 = nchunk * NB;
  const int nb_cur = min(NB, p.cout_planes - nb0);

  const uint32_t a_bytes = (uint32_t)p.pc * ah * aw * 16, d_bytes = (uint32_t)p.d_planes * th * kWmTW * 16;
  const uint32_t stage_bytes = ((a_bytes + 127u) & ~127u) + ((d_bytes + 127u) & ~127u);
  uint8_t *s_stage = smem_raw;
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + (size_t)n_stages * stage_bytes);   // full[S], empty[S]
  __nv_bfloat16 *s_ones = reinterpret_cast<__nv_bfloat16 *>(bars + 2 * n_stages);
  float *s_acc = reinterpret_cast<float *>(s_ones + 64);
  for (int i = threadIdx.x; i < 64; i += blockDim.x) s_ones[i] = __float2bfloat16(1.0f);
  for (int i = threadIdx.x; i < (kWmMaxGroups + 2) * 8 * NB * 8; i += blockDim.x) s_acc[i] = 0.f;
  if (threadIdx.x == 0) {
    for (int i = 0; i < n_stages; ++i) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(wm_smem_u32(&bars[i])), "r"(1) : "memory");
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(wm_smem_u32(&bars[n_stages + i])), "r"(kWmCW) : "memory");
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;

  if (warp == 0) {
    // ---------------- producer ----------------
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, %1;\n\t@px mov.s32 %0, 1;\n\t}" : "+r"(pred) : "r"(0xFFFFFFFFu));
    int st = 0;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int txi = tile % p.tiles_x;
      const int tyi = (tile / p.tiles_x) % p.tiles_y;
      const int img = tile / (p.tiles_x * p.tiles_y);
      if (!wm_wait(wm_smem_u32(&bars[n_stages + st]), ph ^ 1u)) { if (pred) atomicCAS(status, 0, 21); break; }
      if (pred) {
        const uint32_t full = wm_smem_u32(&bars[st]);
        const uint32_t dst_a = wm_smem_u32(s_stage + (size_t)st * stage_bytes);
        const uint32_t dst_d = dst_a + ((a_bytes + 127u) & ~127u);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full), "r"(a_bytes + d_bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                     ::"r"(dst_a), "l"(&map_a), "r"(full), "r"((txi * kWmTW - p.pl) * 2), "r"(tyi * th - p.pt), "r"(plane0), "r"(img) : "memory");
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                     ::"r"(dst_d), "l"(&map_d), "r"(full), "r"(txi * kWmTW * 2), "r"(tyi * th), "r"(nb0), "r"(img) : "memory");
      }
      if (++st == n_stages) { st = 0; ph ^= 1u; }
    }
  } else {
    // ---------------- consumers: 16 warps = 8 tile rows x 2 halves of the m16 tiles ----------------
    // Warp (row cw, half) owns rows cw, cw+8, ... of every tile and the m16 tiles [mt_base, mt_base+n_loc):
    // splitting M instead of pixels halves the accumulator registers per thread, so twice as many warps
    // fit and hide the ldmatrix -> mma latency (the 8-warp version kept the tensor pipe 22-28 % busy).
    const int cwi = warp - 1;
    const int cw = cwi & 7, part = cwi >> 3;
    const int per = (n_mt + kWmSplit - 1) / kWmSplit;
    const int mt_base = part * per;
    const int n_loc = max(0, min(per, n_mt - mt_base));
    uint32_t a_off[kWmLocalMT];
    {
      const int i = lane >> 3, r = lane & 7;
#pragma unroll
      for (int j = 0; j < kWmLocalMT; ++j) {
        const int mt = mt_base + j;
        int g = 2 * mt + (i & 1);
        uint32_t off = 0;
        if (j < n_loc) {
          if (g >= ng_all) g = 0;
          if (g == ones_group) off = 0x80000000u | (uint32_t)(r * 16);
          else {
            const int gg = g0 + g;
            const int t = gg / pc_cur, cgl = gg - t * pc_cur;
            const int dy = t / p.kw, dx = t - dy * p.kw;
            off = (uint32_t)((((cgl * ah + dy) * aw + dx) + r + 8 * (i >> 1)) * 16);
          }
        }
        a_off[j] = off;
      }
    }
    const uint32_t so_base = wm_smem_u32(s_ones);
    float acc[kWmLocalMT][NB][4];
#pragma unroll
    for (int mt = 0; mt < kWmLocalMT; ++mt)
#pragma unroll
      for (int nb = 0; nb < NB; ++nb)
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[mt][nb][k] = 0.f;
    int st = 0;
    uint32_t ph = 0;
    bool ok = true;
    for (int tile = blockIdx.x; ok && tile < p.num_tiles; tile += gridDim.x) {
      ok = wm_wait(wm_smem_u32(&bars[st]), ph);
      if (!ok) { if (lane == 0) atomicCAS(status, 0, 22); break; }
      const uint32_t sa_base = wm_smem_u32(s_stage + (size_t)st * stage_bytes);
      const uint32_t sd_base = sa_base + ((a_bytes + 127u) & ~127u);
      for (int y = cw; y < th; y += 8) {
#pragma unroll
        for (int ks = 0; ks < kWmTW / 16; ++ks) {
          uint32_t bfrag[NB][2];
#pragma unroll
          for (int nb = 0; nb < NB; ++nb) {
            if (nb < nb_cur) {
              const int i = (lane >> 3) & 1, r = lane & 7;
              ldmatrix_x2_trans(sd_base + (uint32_t)((((nb * th + y) * kWmTW) + ks * 16 + i * 8 + r) * 16), bfrag[nb]);
            } else { bfrag[nb][0] = 0u; bfrag[nb][1] = 0u; }
          }
          const uint32_t row_off = (uint32_t)((y * aw + ks * 16) * 16);
#pragma unroll
          for (int mt = 0; mt < kWmLocalMT; ++mt) {
            if (mt < n_loc) {
              uint32_t afrag[4];
              const uint32_t o = a_off[mt];
              const uint32_t addr = (o & 0x80000000u) ? so_base + (o & 0x7FFFFFFFu) : sa_base + o + row_off;
              ldmatrix_x4_trans(addr, afrag);
#pragma unroll
              for (int nb = 0; nb < NB; ++nb) mma_bf16_16816(acc[mt][nb], afrag, bfrag[nb]);
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(wm_smem_u32(&bars[n_stages + st])) : "memory");
      if (++st == n_stages) { st = 0; ph ^= 1u; }
    }
#pragma unroll
    for (int mt = 0; mt < kWmLocalMT; ++mt) {
      if (mt < n_loc) {
        const int gm = mt_base + mt;            // global m16 tile
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) {
          const int ci = lane >> 2, co = nb * 8 + (lane & 3) * 2;
          atomicAdd(&s_acc[((2 * gm) * 8 + ci) * (NB * 8) + co], acc[mt][nb][0]);
          atomicAdd(&s_acc[((2 * gm) * 8 + ci) * (NB * 8) + co + 1], acc[mt][nb][1]);
          atomicAdd(&s_acc[((2 * gm + 1) * 8 + ci) * (NB * 8) + co], acc[mt][nb][2]);
          atomicAdd(&s_acc[((2 * gm + 1) * 8 + ci) * (NB * 8) + co + 1], acc[mt][nb][3]);
        }
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < ng_all * 8 * nb_cur * 8; i += blockDim.x) {
    const int co = i % (nb_cur * 8);
    const int ci = (i / (nb_cur * 8)) % 8;
    const int g = i / (nb_cur * 8 * 8);
    const float v = s_acc[(g * 8 + ci) * (NB * 8) + co];
    if (g == ones_group) {
      if (ci == 0) atomicAdd(&p.db[nb0 * 8 + co], v);
    } else {
      const int gg = g0 + g;
      const int t = gg / pc_cur, cgl = gg - t * pc_cur;
      atomicAdd(&p.dW[((long long)t * p.cin + (plane0 + cgl) * 8 + ci) * p.cout + nb0 * 8 + co], v);
    }
  }
}

// ---------------------------------------------------------------------------------
// Deep layers (Cin >= 32: few pixels, many channels).  The kernel above replicates every activation tile
// over (group block, Cout chunk) CTAs -- 24 CTAs fetch the same tile for a 128->128 layer -- and each CTA
// does little math per fetched tile, so those layers were L2->smem and latency bound (0.14 ms for 9.7 GFLOP).
// Here one CTA owns ALL (tap, plane) groups of an 8-plane input chunk and up to 64 output channels: the 24
// compute warps tile the output block as MP x NP (3 m16 tiles x NBW n8 tiles per warp, accumulators in
// registers), every warp walks all rows of the pixel tile, and each A fragment feeds NBW MMAs.  No smem
// reduction: a warp owns its accumulators exclusively and adds them to dW with 8-byte vector atomics.
// The bias gradient (the ones-row of the kernel above) is a separate tiny reduction (bias_grad_kernel).
// ---------------------------------------------------------------------------------
constexpr int kWdLocalMT = 3;
constexpr int kWdCW = 24;
constexpr int kWdThreads = 32 * (1 + kWdCW);

template <int NBW>
__global__ void __launch_bounds__(kWdThreads) wgrad_deep_kernel(const __grid_constant__ CUtensorMap map_a,
                                                                const __grid_constant__ CUtensorMap map_d,
                                                                const WgParams p, int th, int n_stages, int MP, int NP,
                                                                int *status) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const int aw = kWmTW + p.kw - 1, ah = th + p.kh - 1;
  const int nchunk = blockIdx.y % p.n_nchunks;
  const int pchunk = blockIdx.y / p.n_nchunks;
  const int plane0 = pchunk * p.pc;
  const int pc_cur = min(p.pc, p.cin_planes - plane0);
  const int g_total = p.kh * p.kw * pc_cur;
  const int n_mt = (g_total + 1) / 2;
  const int nb0 = nchunk * NP * NBW;                       // first output plane (n8 tile) of this CTA
  const int nb_cur = min(NP * NBW, p.cout_planes - nb0);

  const uint32_t a_bytes = (uint32_t)p.pc * ah * aw * 16, d_bytes = (uint32_t)p.d_planes * th * kWmTW * 16;
  const uint32_t stage_bytes = ((a_bytes + 127u) & ~127u) + ((d_bytes + 127u) & ~127u);
  uint8_t *s_stage = smem_raw;
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + (size_t)n_stages * stage_bytes);   // full[S], empty[S]
  if (threadIdx.x == 0) {
    for (int i = 0; i < n_stages; ++i) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(wm_smem_u32(&bars[i])), "r"(1) : "memory");
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(wm_smem_u32(&bars[n_stages + i])), "r"(kWdCW) : "memory");
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;

  if (warp == 0) {
    // ---------------- producer ----------------
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, %1;\n\t@px mov.s32 %0, 1;\n\t}" : "+r"(pred) : "r"(0xFFFFFFFFu));
    int st = 0;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int txi = tile % p.tiles_x;
      const int tyi = (tile / p.tiles_x) % p.tiles_y;
      const int img = tile / (p.tiles_x * p.tiles_y);
      if (!wm_wait(wm_smem_u32(&bars[n_stages + st]), ph ^ 1u)) { if (pred) atomicCAS(status, 0, 23); break; }
      if (pred) {
        const uint32_t full = wm_smem_u32(&bars[st]);
        const uint32_t dst_a = wm_smem_u32(s_stage + (size_t)st * stage_bytes);
        const uint32_t dst_d = dst_a + ((a_bytes + 127u) & ~127u);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full), "r"(a_bytes + d_bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                     ::"r"(dst_a), "l"(&map_a), "r"(full), "r"((txi * kWmTW - p.pl) * 2), "r"(tyi * th - p.pt), "r"(plane0), "r"(img) : "memory");
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                     ::"r"(dst_d), "l"(&map_d), "r"(full), "r"(txi * kWmTW * 2), "r"(tyi * th), "r"(nb0), "r"(img) : "memory");
      }
      if (++st == n_stages) { st = 0; ph ^= 1u; }
    }
  } else {
    // ---------------- consumers: warp (mpart, npart) owns 3 m16 tiles x NBW n8 tiles of the CTA's block ----------------
    const int cwi = warp - 1;
    const bool active = cwi < MP * NP;
    const int mpart = cwi / NP, npart = cwi - mpart * NP;
    const int mt_base = mpart * kWdLocalMT;
    const int n_loc = active ? max(0, min(kWdLocalMT, n_mt - mt_base)) : 0;
    const int nbw0 = npart * NBW;                           // first n8 tile of this warp inside the CTA's dz box
    const int nbw_cur = active ? max(0, min(NBW, nb_cur - nbw0)) : 0;
    uint32_t a_off[kWdLocalMT];
    {
      const int i = lane >> 3, r = lane & 7;
#pragma unroll
      for (int j = 0; j < kWdLocalMT; ++j) {
        int g = 2 * (mt_base + j) + (i & 1);
        uint32_t off = 0;
        if (j < n_loc) {
          if (g >= g_total) g = 0;                          // odd group count: the unused half reads valid memory, result dropped
          const int t = g / pc_cur, cgl = g - t * pc_cur;
          const int dy = t / p.kw, dx = t - dy * p.kw;
          off = (uint32_t)((((cgl * ah + dy) * aw + dx) + r + 8 * (i >> 1)) * 16);
        }
        a_off[j] = off;
      }
    }
    float acc[kWdLocalMT][NBW][4];
#pragma unroll
    for (int j = 0; j < kWdLocalMT; ++j)
#pragma unroll
      for (int nb = 0; nb < NBW; ++nb)
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[j][nb][k] = 0.f;
    int st = 0;
    uint32_t ph = 0;
    bool ok = true;
    for (int tile = blockIdx.x; ok && tile < p.num_tiles; tile += gridDim.x) {
      ok = wm_wait(wm_smem_u32(&bars[st]), ph);
      if (!ok) { if (lane == 0) atomicCAS(status, 0, 24); break; }
      const uint32_t sa_base = wm_smem_u32(s_stage + (size_t)st * stage_bytes);
      const uint32_t sd_base = sa_base + ((a_bytes + 127u) & ~127u);
      if (n_loc > 0 && nbw_cur > 0) {
        for (int y = 0; y < th; ++y) {
#pragma unroll
          for (int ks = 0; ks < kWmTW / 16; ++ks) {
            uint32_t bfrag[NBW][2];
#pragma unroll
            for (int nb = 0; nb < NBW; ++nb) {
              if (nb < nbw_cur) {
                const int i = (lane >> 3) & 1, r = lane & 7;
                ldmatrix_x2_trans(sd_base + (uint32_t)(((((nbw0 + nb) * th + y) * kWmTW) + ks * 16 + i * 8 + r) * 16), bfrag[nb]);
              } else { bfrag[nb][0] = 0u; bfrag[nb][1] = 0u; }
            }
            const uint32_t row_off = (uint32_t)((y * aw + ks * 16) * 16);
#pragma unroll
            for (int j = 0; j < kWdLocalMT; ++j) {
              if (j < n_loc) {
                uint32_t afrag[4];
                ldmatrix_x4_trans(sa_base + a_off[j] + row_off, afrag);
#pragma unroll
                for (int nb = 0; nb < NBW; ++nb) mma_bf16_16816(acc[j][nb], afrag, bfrag[nb]);
              }
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(wm_smem_u32(&bars[n_stages + st])) : "memory");
      if (++st == n_stages) { st = 0; ph ^= 1u; }
    }
    // ---- this warp's block goes straight to global memory: 8-byte vector atomics (two adjacent output channels)
#pragma unroll
    for (int j = 0; j < kWdLocalMT; ++j) {
      if (j < n_loc) {
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          const int g = 2 * (mt_base + j) + hf;
          if (g < g_total) {
            const int t = g / pc_cur, cgl = g - t * pc_cur;
            const int ci = (plane0 + cgl) * 8 + (lane >> 2);
#pragma unroll
            for (int nb = 0; nb < NBW; ++nb) {
              if (nb < nbw_cur) {
                const int co = (nb0 + nbw0 + nb) * 8 + (lane & 3) * 2;
                float2 *dst = reinterpret_cast<float2 *>(&p.dW[((long long)t * p.cin + ci) * p.cout + co]);
                atomicAdd(dst, make_float2(acc[j][nb][2 * hf], acc[j][nb][2 * hf + 1]));
              }
            }
          }
        }
      }
    }
  }
}

// bias gradient of the deep layers: db[co] += sum over pixels of dz[.., co]  (dz is dense [n][planes][h][w][8])
__global__ void __launch_bounds__(256) bias_grad_kernel(const __nv_bfloat16 *__restrict__ dz, long long img_stride, int hw,
                                                        float *__restrict__ db) {
  const int pl = blockIdx.y, img = blockIdx.z;
  const __nv_bfloat16 *base = dz + (long long)img * img_stride + (long long)pl * hw * 8;
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < hw; v += gridDim.x * blockDim.x) {
    const uint4 u = *reinterpret_cast<const uint4 *>(base + (long long)v * 8);
    const __nv_bfloat162 *h2 = reinterpret_cast<const __nv_bfloat162 *>(&u);
#pragma unroll
    for (int k = 0; k < 4; ++k) { const float2 f = __bfloat1622float2(h2[k]); acc[2 * k] += f.x; acc[2 * k + 1] += f.y; }
  }
  __shared__ float red[8][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    float v = acc[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < 8) {
    float v = 0.f;
    for (int wv = 0; wv < 8; ++wv) v += red[wv][threadIdx.x];
    atomicAdd(&db[pl * 8 + threadIdx.x], v);
  }
}

// ---------------------------------------------------------------------------------
// Row-walking kernel for the narrow layers (Cin = 8 / 16: the full-resolution layers that carry most of the bytes).
// The TMA kernel above gives every warp single tile rows, so each of the kh*kw taps of a row is its own ldmatrix.x4
// plus address arithmetic: ~26 instructions per MMA, issue-bound at IPC 2.5 with DRAM at 33-45 % (profiles/r1_wgrad_*).
// Here a warp owns a band of R consecutive output rows of one 16-pixel column and ONE output plane, and walks the
// INPUT rows of its band: the fragments of input row j (one ldmatrix.x4 per pair of (dx, plane) column groups) feed
// the taps dy = 0..kh-1 of output rows j, j-1, j-2, whose dz fragments sit in a rolling register window.  Everything
// is compile-time (PC planes, NB output planes per CTA, KS = kernel size, band height), the loop over the band is
// fully unrolled and every shared-memory address is `base + immediate`: ~1.4 instructions per MMA.  The bias gradient
// is one more MMA per dz fragment against an all-ones A fragment held in registers.
// UPS = 1: the decoder's 2x2 convolution reads the x2-upsampled tensor; ldmatrix takes one row address per lane, so
// the virtual pixel (y, x) is simply the low-res smem pixel (y >> 1, x >> 1) -- the up-sampled tensor is never built.
// ---------------------------------------------------------------------------------
constexpr int kWrCW = 16;                          // compute warps: bands x 2 k-steps x NB output planes x MP parts of the (dx, plane) groups
constexpr int kWrThreads = 32 * (1 + kWrCW);
constexpr int kWrTH = 32;                          // tile rows (16 for Cin >= 32)

template <int PC, int NB, int KS, int UPS>
struct WrGeo {
  static constexpr int TH = PC > 2 ? kWrTH / 2 : kWrTH, TW = kWmTW;      // >= 32 input channels: half-height tiles (smem)
  static constexpr int AW = UPS ? (TW / 2 + 1) : (TW + KS - 1);
  static constexpr int AH = UPS ? (TH / 2 + 1) : (TH + KS - 1);
  static constexpr int U = KS * PC, LT = (U + 1) / 2;      // (dx, plane) column groups, m16 loads (pairs of groups)
  static constexpr int MP = PC > 4 ? 2 : 1;                // 64 input channels: two warps share the loads of a band (registers)
  static constexpr int L = (LT + MP - 1) / MP;             // loads per warp
  static constexpr int BANDS = 8 / (NB * MP), R = TH / BANDS;
  static constexpr uint32_t A_BYTES = PC * AH * AW * 16, D_BYTES = NB * TH * TW * 16;
  static constexpr uint32_t A_PAD = (A_BYTES + 127u) & ~127u;
  static constexpr uint32_t STAGE = A_PAD + ((D_BYTES + 127u) & ~127u);
  static constexpr int ACC_FLOATS = KS * KS * PC * 8 * NB * 8 + NB * 8;
};

template <int PC, int NB, int KS, int UPS>
__global__ void __launch_bounds__(kWrThreads, 1) wgrad_rows_kernel(const __grid_constant__ CUtensorMap map_a,
                                                                   const __grid_constant__ CUtensorMap map_d,
                                                                   const WgParams p, int n_stages, int *status) {
  using G = WrGeo<PC, NB, KS, UPS>;
  constexpr int TH = G::TH, TW = G::TW, AW = G::AW, AH = G::AH, U = G::U, L = G::L, LT = G::LT, MP = G::MP, R = G::R;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *s_stage = smem_raw;
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + (size_t)n_stages * G::STAGE);   // full[S], empty[S]
  float *s_acc = reinterpret_cast<float *>(smem_raw);      // [tap][ci][NB*8] + bias[NB*8]: reuses the stages once they are drained
  if (threadIdx.x == 0) {
    for (int i = 0; i < n_stages; ++i) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(wm_smem_u32(&bars[i])), "r"(1) : "memory");
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(wm_smem_u32(&bars[n_stages + i])), "r"(kWrCW) : "memory");
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int nb0 = blockIdx.y * NB;                  // first output plane of this CTA

  if (warp == 0) {
    // ---------------- producer ----------------
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, %1;\n\t@px mov.s32 %0, 1;\n\t}" : "+r"(pred) : "r"(0xFFFFFFFFu));
    int st = 0;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int txi = tile % p.tiles_x;
      const int tyi = (tile / p.tiles_x) % p.tiles_y;
      const int img = tile / (p.tiles_x * p.tiles_y);
      if (!wm_wait(wm_smem_u32(&bars[n_stages + st]), ph ^ 1u)) { if (pred) atomicCAS(status, 0, 25); break; }
      if (pred) {
        const uint32_t full = wm_smem_u32(&bars[st]);
        const uint32_t dst_a = wm_smem_u32(s_stage + (size_t)st * G::STAGE);
        const uint32_t dst_d = dst_a + G::A_PAD;
        const int ax = UPS ? txi * (TW / 2) : txi * TW - p.pl, ay = UPS ? tyi * (TH / 2) : tyi * TH - p.pt;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full), "r"(G::A_BYTES + G::D_BYTES) : "memory");
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                     ::"r"(dst_a), "l"(&map_a), "r"(full), "r"(ax * 2), "r"(ay), "r"(0), "r"(img) : "memory");
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                     ::"r"(dst_d), "l"(&map_d), "r"(full), "r"(txi * TW * 2), "r"(tyi * TH), "r"(nb0), "r"(img) : "memory");
      }
      if (++st == n_stages) { st = 0; ph ^= 1u; }
    }
  } else {
    // ---------------- consumers: warp = (band, k-step, output plane) ----------------
    const int cwi = warp - 1;
    const int nbp = cwi % NB, ks = (cwi / NB) & 1, mpart = (cwi / (2 * NB)) % MP, band = cwi / (2 * NB * MP);
    const int y0 = band * R;
    const bool live = (nb0 + nbp) < p.cout_planes;
    // per-lane ldmatrix row addresses (bytes inside a stage): matrix i = lane >> 3 (group half i & 1, pixel octet i >> 1),
    // row r = lane & 7 = pixel inside the octet
    uint32_t a_lane[L];
    {
      const int i = lane >> 3, r = lane & 7;
#pragma unroll
      for (int l = 0; l < L; ++l) {
        int u = 2 * (mpart * L + l) + (i & 1);
        if (u >= U) u = U - 1;                      // odd group count / short last part: re-reads the last group, result dropped
        const int dx = u / PC, c = u - dx * PC;
        const int px = ks * 16 + dx + r + 8 * (i >> 1);
        const int row0 = UPS ? (y0 >> 1) : y0;
        a_lane[l] = (uint32_t)(((c * AH + row0) * AW + (UPS ? (px >> 1) : px)) * 16);
      }
    }
    const uint32_t b_lane = G::A_PAD + (uint32_t)((((nbp * TH + y0) * TW) + ks * 16 + ((lane >> 3) & 1) * 8 + (lane & 7)) * 16);
    const uint32_t ones[4] = {0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u};     // bf16 1.0 pairs
    float acc[L][KS][4], accb[4];
#pragma unroll
    for (int l = 0; l < L; ++l)
#pragma unroll
      for (int d = 0; d < KS; ++d)
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[l][d][k] = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) accb[k] = 0.f;
    int st = 0;
    uint32_t ph = 0;
    bool ok = true;
    for (int tile = blockIdx.x; ok && tile < p.num_tiles; tile += gridDim.x) {
      ok = wm_wait(wm_smem_u32(&bars[st]), ph);
      if (!ok) { if (lane == 0) atomicCAS(status, 0, 26); break; }
      if (live) {
        const uint32_t base = wm_smem_u32(s_stage + (size_t)st * G::STAGE);
        uint32_t bw[KS][2];
        uint32_t af[L][4];
#pragma unroll
        for (int d = 0; d < KS; ++d) { bw[d][0] = 0u; bw[d][1] = 0u; }
#pragma unroll
        for (int jj = 0; jj < R + KS - 1; ++jj) {
          // input row y0 + jj serves output rows y0 + jj - dy; dz fragments of the last KS rows in a rolling window
#pragma unroll
          for (int d = KS - 1; d > 0; --d) { bw[d][0] = bw[d - 1][0]; bw[d][1] = bw[d - 1][1]; }
          if (jj < R) {
            ldmatrix_x2_trans(base + b_lane + (uint32_t)(jj * TW * 16), bw[0]);
            if (MP == 1 || mpart == 0) mma_bf16_16816(accb, ones, bw[0]);
          }
          if (!UPS || (jj & 1) == 0) {
#pragma unroll
            for (int l = 0; l < L; ++l)
              if (MP == 1 || mpart * L + l < LT) ldmatrix_x4_trans(base + a_lane[l] + (uint32_t)((UPS ? (jj >> 1) : jj) * AW * 16), af[l]);
          }
#pragma unroll
          for (int d = 0; d < KS; ++d) {
            if (jj - d >= 0 && jj - d < R) {
#pragma unroll
              for (int l = 0; l < L; ++l)
                if (MP == 1 || mpart * L + l < LT) mma_bf16_16816(acc[l][d], af[l], bw[d]);
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(wm_smem_u32(&bars[n_stages + st])) : "memory");
      if (++st == n_stages) { st = 0; ph ^= 1u; }
    }
    // ---- the CTA's dW block: shared-memory accumulation over the warps of each output plane.  The accumulator block
    //      lives in the (now drained) stage memory: every TMA load issued has been waited for by all consumers.
    asm volatile("bar.sync 1, %0;" ::"r"(kWrCW * 32) : "memory");
    for (int i = threadIdx.x - 32; i < G::ACC_FLOATS; i += kWrCW * 32) s_acc[i] = 0.f;
    asm volatile("bar.sync 1, %0;" ::"r"(kWrCW * 32) : "memory");
    if (live) {
      const int ci = lane >> 2, co = nbp * 8 + (lane & 3) * 2;
#pragma unroll
      for (int l = 0; l < L; ++l) {
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          const int u = 2 * (mpart * L + l) + hf;
          if (u < U) {
            const int dx = u / PC, c = u - dx * PC;
#pragma unroll
            for (int d = 0; d < KS; ++d) {
              float *dst = s_acc + (((d * KS + dx) * PC + c) * 8 + ci) * (NB * 8) + co;
              atomicAdd(dst, acc[l][d][2 * hf]);
              atomicAdd(dst + 1, acc[l][d][2 * hf + 1]);
            }
          }
        }
      }
      if (lane < 4 && (MP == 1 || mpart == 0)) {      // every row of the ones-MMA holds sum(dz): take row 0
        atomicAdd(s_acc + KS * KS * PC * 8 * NB * 8 + co, accb[0]);
        atomicAdd(s_acc + KS * KS * PC * 8 * NB * 8 + co + 1, accb[1]);
      }
    }
  }
  __syncthreads();
  const int nb_cur = min(NB, p.cout_planes - nb0);
  for (int i = threadIdx.x; i < KS * KS * PC * 8 * NB * 8; i += blockDim.x) {
    const int co = i % (NB * 8);
    if (co >= nb_cur * 8) continue;
    const int row = i / (NB * 8);                   // (tap * PC + c) * 8 + ci
    const int t = row / (PC * 8), cc = row - t * (PC * 8);
    atomicAdd(&p.dW[((long long)t * p.cin + cc) * p.cout + nb0 * 8 + co], s_acc[i]);
  }
  if (p.db && threadIdx.x < nb_cur * 8) atomicAdd(&p.db[nb0 * 8 + threadIdx.x], s_acc[KS * KS * PC * 8 * NB * 8 + threadIdx.x]);
}

template <int PC, int NB, int KS, int UPS>
static int launch_wgrad_rows_t(const CUtensorMap &map_a, const CUtensorMap &map_d, const WgParams &p, int *status, cudaStream_t st) {
  using G = WrGeo<PC, NB, KS, UPS>;
  const size_t fixed = 2 * kWmMaxStages * 8 + 1024;
  int n_stages = (int)std::min<size_t>(kWmMaxStages, (224 * 1024 - fixed) / G::STAGE);
  if (n_stages < 2 || (size_t)n_stages * G::STAGE < (size_t)G::ACC_FLOATS * sizeof(float)) { set_error("wgrad_rows: smem budget exceeded"); return 1; }
  const size_t smem = (size_t)n_stages * G::STAGE + fixed;
  static PerDeviceOnce attr;
  if (const int dev_ = attr.pending(); dev_ >= 0) {
    OCTSEG_CUDA(cudaFuncSetAttribute(wgrad_rows_kernel<PC, NB, KS, UPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
    attr.mark(dev_);
  }
  const int blocks_y = p.n_nchunks;
  const int gx = std::max(1, std::min(p.num_tiles, (148 + blocks_y - 1) / blocks_y));
  wgrad_rows_kernel<PC, NB, KS, UPS><<<dim3(gx, blocks_y), kWrThreads, smem, st>>>(map_a, map_d, p, n_stages, status);
  OCTSEG_CUDA(cudaGetLastError());
  return 0;
}

bool wgrad_rows_applicable(int kh, int kw, int cin, int ups) {
  static const bool off = []() { const char *e = std::getenv("OCTSEG_WGRAD_ROWS"); return e && e[0] == '0'; }();
  if (off || kh != kw || (cin != 8 && cin != 16 && cin != 32 && cin != 64)) return false;
  if (cin == 64 && kh != 2) return false;      // 64 x 3x3: measured slower than wgrad_deep_kernel (0.24 vs 0.21 ms at batch 256)
  return ups ? kh == 2 : (kh == 2 || kh == 3);
}

// a_in: the conv's input (low-res for ups = 1), dz on the output grid; both dense
static int launch_wgrad_rows(View<const __nv_bfloat16> a_in, View<const __nv_bfloat16> dz, WgParams p, int *status, cudaStream_t st) {
  const int ks = p.kh, pc = p.cin_planes, ups = p.ups;
  const int nb = pc > 2 ? 2 : std::min(2, p.cout_planes);
  const int th = pc > 2 ? kWrTH / 2 : kWrTH;
  p.n_nchunks = (p.cout_planes + nb - 1) / nb;
  p.tiles_y = (dz.h + th - 1) / th;
  p.num_tiles = dz.n * p.tiles_x * p.tiles_y;
  p.d_planes = nb;
  const int aw = ups ? kWmTW / 2 + 1 : kWmTW + ks - 1, ah = ups ? th / 2 + 1 : th + ks - 1;
  CUtensorMap map_a, map_d;
  if (tc_encode_map_4d(a_in.ptr, a_in.w, a_in.h, p.cin_planes, a_in.n, aw, ah, pc, &map_a)) return 1;
  if (tc_encode_map_4d(dz.ptr, dz.w, dz.h, p.cout_planes, dz.n, kWmTW, th, nb, &map_d)) return 1;
#define OCTSEG_WR(PC_, NB_, KS_, UPS_) \
  if (pc == PC_ && nb == NB_ && ks == KS_ && ups == UPS_) return launch_wgrad_rows_t<PC_, NB_, KS_, UPS_>(map_a, map_d, p, status, st);
  OCTSEG_WR(1, 1, 3, 0) OCTSEG_WR(1, 2, 3, 0) OCTSEG_WR(2, 1, 3, 0) OCTSEG_WR(2, 2, 3, 0)
  OCTSEG_WR(1, 1, 2, 0) OCTSEG_WR(1, 2, 2, 0) OCTSEG_WR(2, 1, 2, 0) OCTSEG_WR(2, 2, 2, 0)
  OCTSEG_WR(1, 1, 2, 1) OCTSEG_WR(1, 2, 2, 1) OCTSEG_WR(2, 1, 2, 1) OCTSEG_WR(2, 2, 2, 1)
  OCTSEG_WR(4, 2, 3, 0) OCTSEG_WR(4, 2, 2, 0) OCTSEG_WR(4, 2, 2, 1)
  OCTSEG_WR(8, 2, 2, 0) OCTSEG_WR(8, 2, 2, 1)
#undef OCTSEG_WR
  set_error("wgrad_rows: no instantiation");
  return 1;
}

int launch_wgrad_mma(View<const __nv_bfloat16> a_in, View<const __nv_bfloat16> dz, int kh, int kw, int pad_top,
                     int pad_left, int ups, int cin, int cout, float *dW, float *db, int *status,
                     cudaStream_t st) {
  if (kh > 3 || kw > 3) { set_error("wgrad: kernel larger than 3x3 not supported"); return 1; }
  WgParams p{};
  p.a = a_in.ptr; p.a_img_stride = a_in.img_stride; p.a_h = a_in.h; p.a_w = a_in.w;
  p.dz = dz.ptr; p.dz_img_stride = dz.img_stride; p.H = dz.h; p.W = dz.w;
  p.n = dz.n; p.kh = kh; p.kw = kw; p.pt = pad_top; p.pl = pad_left; p.ups = ups;
  p.cin = cin; p.cout = cout; p.cin_planes = cin / 8; p.cout_planes = cout / 8;
  p.pc = std::min(p.cin_planes, 8);
  p.n_pchunks = (p.cin_planes + p.pc - 1) / p.pc;
  p.n_gblocks = (kh * kw * p.pc + kWmMaxGroups - 1) / kWmMaxGroups;
  p.n_nchunks = (p.cout_planes + kWmNB - 1) / kWmNB;
  p.tiles_x = (dz.w + kWmTW - 1) / kWmTW;
  p.tiles_y = (dz.h + kWmTH - 1) / kWmTH;
  p.num_tiles = dz.n * p.tiles_x * p.tiles_y;
  p.dW = dW; p.db = db;
  const bool dense = a_in.img_stride == (long long)a_in.planes * a_in.h * a_in.w * 8 &&
                     dz.img_stride == (long long)dz.planes * dz.h * dz.w * 8 && a_in.planes == p.cin_planes;
  if (dense && status && wgrad_rows_applicable(kh, kw, cin, ups) &&
      (ups ? (pad_top == 0 && pad_left == 0 && dz.h == 2 * a_in.h && dz.w == 2 * a_in.w) : (dz.h == a_in.h && dz.w == a_in.w)))
    return launch_wgrad_rows(a_in, dz, p, status, st);
  static const bool no_deep = []() { const char *e = std::getenv("OCTSEG_NO_DEEP_WGRAD"); return e && e[0] == '1'; }();
  if (!ups && dense && status && !no_deep && p.cin_planes >= 4) {
    // ---- deep-layer kernel: pick the warp tiling MP x NP (3 m16 tiles x NBW n8 tiles per warp)
    const int g_total = kh * kw * p.pc, n_mt = (g_total + 1) / 2;
    const int MP = (n_mt + kWdLocalMT - 1) / kWdLocalMT;
    int NBW = 0, NP = 0;
    for (int cand : {4, 2, 1}) {
      const int np = std::min(kWdCW / std::max(1, MP), (p.cout_planes + cand - 1) / cand);
      if (MP <= kWdCW && np >= 1 && MP * np >= 12) { NBW = cand; NP = np; break; }
    }
    if (NBW) {
      const int th = 8;
      p.tiles_y = (dz.h + th - 1) / th;
      p.num_tiles = dz.n * p.tiles_x * p.tiles_y;
      p.d_planes = std::min(NP * NBW, p.cout_planes);
      p.n_nchunks = (p.cout_planes + NP * NBW - 1) / (NP * NBW);
      const int aw2 = kWmTW + kw - 1, ah2 = th + kh - 1;
      CUtensorMap map_a, map_d;
      if (tc_encode_map_4d(a_in.ptr, a_in.w, a_in.h, p.cin_planes, a_in.n, aw2, ah2, p.pc, &map_a)) return 1;
      if (tc_encode_map_4d(dz.ptr, dz.w, dz.h, p.cout_planes, dz.n, kWmTW, th, p.d_planes, &map_d)) return 1;
      const size_t a_bytes = (size_t)p.pc * ah2 * aw2 * 16, d_bytes = (size_t)p.d_planes * th * kWmTW * 16;
      const size_t stage = ((a_bytes + 127) & ~(size_t)127) + ((d_bytes + 127) & ~(size_t)127);
      const size_t fixed = 2 * kWmMaxStages * 8 + 1024;
      int n_stages = (int)std::min<size_t>(kWmMaxStages, (200 * 1024 - fixed) / stage);
      if (n_stages >= 2) {
        const size_t smem3 = n_stages * stage + fixed;
        static PerDeviceOnce attr3;
        if (const int dev_ = attr3.pending(); dev_ >= 0) {
          OCTSEG_CUDA(cudaFuncSetAttribute(wgrad_deep_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
          OCTSEG_CUDA(cudaFuncSetAttribute(wgrad_deep_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
          OCTSEG_CUDA(cudaFuncSetAttribute(wgrad_deep_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
          attr3.mark(dev_);
        }
        const int blocks_y3 = p.n_pchunks * p.n_nchunks;
        const int gx3 = std::max(1, std::min(p.num_tiles, (148 + blocks_y3 - 1) / blocks_y3));
        dim3 grid3(gx3, blocks_y3);
        if (NBW == 4) wgrad_deep_kernel<4><<<grid3, kWdThreads, smem3, st>>>(map_a, map_d, p, th, n_stages, MP, NP, status);
        else if (NBW == 2) wgrad_deep_kernel<2><<<grid3, kWdThreads, smem3, st>>>(map_a, map_d, p, th, n_stages, MP, NP, status);
        else wgrad_deep_kernel<1><<<grid3, kWdThreads, smem3, st>>>(map_a, map_d, p, th, n_stages, MP, NP, status);
        OCTSEG_CUDA(cudaGetLastError());
        if (db) {
          const int hw = dz.h * dz.w;
          dim3 gb(std::max(1, std::min((hw + 2047) / 2048, 32)), p.cout_planes, dz.n);
          bias_grad_kernel<<<gb, 256, 0, st>>>(dz.ptr, dz.img_stride, hw, db);
          OCTSEG_CUDA(cudaGetLastError());
        }
        return 0;
      }
      // fall through to the generic kernels with the original decomposition
      p.n_nchunks = (p.cout_planes + kWmNB - 1) / kWmNB;
    }
  }
  if (!ups && dense && status) {
    const int th = p.pc == 1 ? 32 : (p.pc == 2 ? 16 : 8);
    p.tiles_y = (dz.h + th - 1) / th;
    p.num_tiles = dz.n * p.tiles_x * p.tiles_y;
    const int aw2 = kWmTW + kw - 1, ah2 = th + kh - 1;
    CUtensorMap map_a, map_d;
    if (tc_encode_map_4d(a_in.ptr, a_in.w, a_in.h, p.cin_planes, a_in.n, aw2, ah2, p.pc, &map_a)) return 1;
    p.d_planes = std::min(kWmNB, p.cout_planes);
    if (tc_encode_map_4d(dz.ptr, dz.w, dz.h, p.cout_planes, dz.n, kWmTW, th, p.d_planes, &map_d)) return 1;
    const size_t a_bytes = (size_t)p.pc * ah2 * aw2 * 16, d_bytes = (size_t)p.d_planes * th * kWmTW * 16;
    const size_t stage = ((a_bytes + 127) & ~(size_t)127) + ((d_bytes + 127) & ~(size_t)127);
    const size_t fixed = 2 * kWmMaxStages * 8 + 128 + (size_t)(kWmMaxGroups + 2) * 8 * kWmNB * 8 * sizeof(float) + 1024;
    int n_stages = (int)std::min<size_t>(kWmMaxStages, (200 * 1024 - fixed) / stage);
    if (n_stages < 2) n_stages = 2;
    const size_t smem2 = n_stages * stage + fixed;
    static PerDeviceOnce attr2;
    if (const int dev_ = attr2.pending(); dev_ >= 0) {
      OCTSEG_CUDA(cudaFuncSetAttribute(wgrad_mma_tma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
      OCTSEG_CUDA(cudaFuncSetAttribute(wgrad_mma_tma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
      attr2.mark(dev_);
    }
    if (smem2 <= 220 * 1024) {
      const int blocks_y2 = p.n_pchunks * p.n_gblocks * p.n_nchunks;
      int gx2 = std::max(1, std::min(p.num_tiles, (148 + blocks_y2 - 1) / blocks_y2));
      dim3 grid2(gx2, blocks_y2);
      if (p.d_planes == 1) wgrad_mma_tma_kernel<1><<<grid2, kWmTmaThreads, smem2, st>>>(map_a, map_d, p, th, n_stages, status);
      else wgrad_mma_tma_kernel<2><<<grid2, kWmTmaThreads, smem2, st>>>(map_a, map_d, p, th, n_stages, status);
      OCTSEG_CUDA(cudaGetLastError());
      return 0;
    }
    p.tiles_y = (dz.h + kWmTH - 1) / kWmTH;
    p.num_tiles = dz.n * p.tiles_x * p.tiles_y;
  }
  const int aw = kWmTW + kw - 1, ah = kWmTH + kh - 1;
  const size_t smem = ((size_t)p.pc * ah * aw * 8 + (size_t)kWmNB * kWmTH * kWmTW * 8 + 64) * 2 +
                      (size_t)(kWmMaxGroups + 2) * 8 * kWmNB * 8 * sizeof(float);
  static PerDeviceOnce attr;
  if (const int dev_ = attr.pending(); dev_ >= 0) {
    OCTSEG_CUDA(cudaFuncSetAttribute(wgrad_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    attr.mark(dev_);
  }
  if (smem > 100 * 1024) { set_error("wgrad_mma: smem budget exceeded"); return 1; }
  const int blocks_y = p.n_pchunks * p.n_gblocks * p.n_nchunks;
  // persistent CTAs: ~2 per SM overall, at least one per output block
  int gx = std::max(1, std::min(p.num_tiles, (148 * 2 + blocks_y - 1) / blocks_y));
  dim3 grid(gx, blocks_y);
  wgrad_mma_kernel<<<grid, 256, smem, st>>>(p);
  OCTSEG_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace octseg
