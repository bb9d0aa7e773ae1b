// tcgen05 weight gradient for layers with >= 64 input channels (reference: the dW Keras' autodiff computes for
// Conv2D, models/unet.py:27).
//
//   dW[dy][dx][ci][co] = sum over (image, y, x) of  a_in[ci][y + dy - pt][x + dx - pl] * dz[co][y][x]
//
// GEMM view, one per filter tap: D_tap[M = ci][N = co] += A_tap^T[ci][K = pixels] * dZ[pixels][co].
// Both operands are the blocked activation tiles [plane][row][px][8 channels] exactly as TMA delivers them:
// with the channel index as the MMA's M (resp. N) dimension this IS the canonical no-swizzle "MN-major"
// UMMA operand -- 8 consecutive pixels x 16 B of one plane form a core matrix (K along the pixels, 16 B apart),
// LBO = 128 B to the next pixel octet of the row, SBO = the plane pitch to the next 8 channels.  A filter tap
// is again a descriptor start-address shift (dy rows, dx pixels), so the halo tile is fetched once and no
// transposed copy of either tensor is ever made.  K = 16 pixels of one tile row per MMA.
//
// Decomposition: a CTA owns one M-block (64 or 128 input channels), one N-chunk (<= 256 output channels with
// kw * N <= 512 TMEM columns), ONE filter row dy and every ksplit-th pixel tile; its kw accumulators
// (one per dx) stay in TMEM for the whole kernel.  Partial sums of the `ksplit` CTAs are written to a scratch
// buffer and added up in a fixed order by wgrad_tc_reduce_kernel: bit-reproducible, no atomics.
//
// Warp roles (256 threads): warp 0 TMA producer + TMEM allocator, warps 1..3 MMA issuers (tap dx = warp - 1),
// warps 4..7 epilogue (TMEM lane quarter = warp % 4).
#include "conv_tc.cuh"
#include "tc_ptx.cuh"
#include "train_kernels.cuh"

#include <algorithm>

namespace octseg {

constexpr int kWtTW = 16;          // pixels per tile row == K of one MMA
constexpr int kWtMaxStages = 6;
constexpr int kWtThreads = 256;

struct WgTcParams {
  int kh, kw, pt, pl;
  int cin, cout;
  int mp;                 // 8-channel planes per M-block: 8 (M = 64) or 16 (M = 128)
  int n_mblocks;
  int nc;                 // output channels per N-chunk
  int n_nchunks;
  int th;                 // tile rows
  int tiles_x, tiles_y, num_tiles;
  int ksplit;
  int stages;
  uint32_t a_bytes, b_bytes;     // per stage, padded to 128 B
  uint32_t a_pitch, a_plane, b_pitch, b_plane;
  float *partials;        // [ksplit][kh][kw][cin][cout]
  int *status;
};

struct __align__(8) WtBarriers {
  uint64_t full[kWtMaxStages], empty[kWtMaxStages];
  uint64_t acc_full;
  uint32_t tmem_base;
  uint32_t pad;
};

// MN-major, no-swizzle descriptor: LBO = bytes between the two 8-pixel K halves, SBO = bytes between 8-channel groups
__device__ __forceinline__ uint64_t wt_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) |
         (1ull << 46);
}

__global__ void __launch_bounds__(kWtThreads, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_d,
                const __grid_constant__ WgTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  uint8_t *a_smem = smem;
  uint8_t *b_smem = smem + (size_t)p.a_bytes * p.stages;
  WtBarriers *bars = reinterpret_cast<WtBarriers *>(b_smem + (size_t)p.b_bytes * p.stages);
  // work item of this CTA
  int j = blockIdx.y;
  const int dy = j % p.kh; j /= p.kh;
  const int ncx = j % p.n_nchunks;
  const int mb = j / p.n_nchunks;
  const int ks = blockIdx.x;
  const uint32_t acc_cols = (uint32_t)(p.kw * p.nc);
  uint32_t tmem_cols = 32;
  while (tmem_cols < acc_cols) tmem_cols <<= 1;

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.stages; ++i) { mbar_init(smem_u32(&bars->full[i]), 1); mbar_init(smem_u32(&bars->empty[i]), 3); }
    mbar_init(smem_u32(&bars->acc_full), 3);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  const int tiles_per_img = p.tiles_x * p.tiles_y;

  if (warp == 0) {
    // ===================== TMA producer =====================
    const bool leader = elect_one_sync() != 0;
    if (leader) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_d) : "memory");
    }
    int s = 0;
    uint32_t ph = 0;
    bool ok = true;
    for (int t = ks; ok && t < p.num_tiles; t += p.ksplit) {
      const int img = t / tiles_per_img, r = t - img * tiles_per_img;
      const int ty = r / p.tiles_x, tx = r - ty * p.tiles_x;
      ok = mbar_wait(smem_u32(&bars->empty[s]), ph ^ 1u, p.status, 11);
      if (!ok) break;
      const uint32_t full = smem_u32(&bars->full[s]);
      if (leader) {
        mbar_expect_tx(full, (uint32_t)(p.a_plane * p.mp + p.b_plane * (p.nc >> 3)));
        // tensor maps are declared in 8-byte elements: x coordinate = px * 2
        tma_load_4d(smem_u32(a_smem + (size_t)s * p.a_bytes), &map_a, full, (tx * kWtTW - p.pl) * 2, ty * p.th + dy - p.pt,
                    mb * p.mp, img);
        tma_load_4d(smem_u32(b_smem + (size_t)s * p.b_bytes), &map_d, full, tx * kWtTW * 2, ty * p.th, ncx * (p.nc >> 3), img);
      }
      if (++s == p.stages) { s = 0; ph ^= 1u; }
    }
  } else if (warp <= 3) {
    // ===================== MMA issuers: one filter tap (dx) each =====================
    const int dx = warp - 1;
    // instruction descriptor: D = f32, A = B = bf16, both MN-major (bits 15, 16), N = nc, M = 8 * mp
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(p.nc >> 3) << 17) |
                           ((uint32_t)((p.mp * 8) >> 4) << 24);
    const uint32_t d_tmem = tmem_base + (uint32_t)(dx * p.nc);
    int s = 0;
    uint32_t ph = 0;
    bool ok = true;
    uint32_t first = 1;
    for (int t = ks; ok && t < p.num_tiles; t += p.ksplit) {
      ok = mbar_wait(smem_u32(&bars->full[s]), ph, p.status, 12);
      if (!ok) break;
      tc_fence_after();
      if (dx < p.kw) {
        const uint32_t a_base = smem_u32(a_smem + (size_t)s * p.a_bytes) + (uint32_t)dx * 16u;
        const uint32_t b_base = smem_u32(b_smem + (size_t)s * p.b_bytes);
        for (int r = 0; r < p.th; ++r) {
          const uint64_t da = wt_desc(a_base + (uint32_t)r * p.a_pitch, 128u, p.a_plane);
          const uint64_t db = wt_desc(b_base + (uint32_t)r * p.b_pitch, 128u, p.b_plane);
          umma_bf16(d_tmem, da, db, idesc, (first && r == 0) ? 0u : 1u);
        }
      }
      first = 0;
      umma_commit(smem_u32(&bars->empty[s]));
      if (++s == p.stages) { s = 0; ph ^= 1u; }
    }
    umma_commit(smem_u32(&bars->acc_full));
  } else {
    // ===================== epilogue: TMEM -> partial dW =====================
    const int q = warp & 3;
    bool ok = mbar_wait(smem_u32(&bars->acc_full), 0, p.status, 13);
    tc_fence_after();
    // accumulator row of this thread: M = 128 -> lane quarter q holds rows 32q .. 32q+31;
    // M = 64 -> rows 16q .. 16q+15 live in lanes 0..15 of quarter q (lanes 16..31 carry nothing)
    const bool m128 = (p.mp == 16);
    const int row = m128 ? q * 32 + lane : q * 16 + lane;
    const bool row_ok = m128 || lane < 16;
    const int ci = mb * p.mp * 8 + row;
    if (ok) {
      for (int dx = 0; dx < p.kw; ++dx) {
        float *dst = p.partials + ((((size_t)ks * p.kh + dy) * p.kw + dx) * p.cin + ci) * p.cout + (size_t)ncx * p.nc;
        for (int c = 0; c < p.nc; c += 8) {
          uint32_t v[8];
          tmem_ld8(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(dx * p.nc + c), v);
          tmem_ld_wait();
          if (row_ok) {
            *reinterpret_cast<float4 *>(dst + c) = make_float4(__uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[2]), __uint_as_float(v[3]));
            *reinterpret_cast<float4 *>(dst + c + 4) = make_float4(__uint_as_float(v[4]), __uint_as_float(v[5]), __uint_as_float(v[6]), __uint_as_float(v[7]));
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

// dW[i] += partials[0][i] + partials[1][i] + ... in that order (deterministic)
__global__ void __launch_bounds__(256) wgrad_tc_reduce_kernel(const float *__restrict__ partials, int ksplit, long long count,
                                                               float *__restrict__ dW) {
  for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < count; i += (long long)gridDim.x * blockDim.x * 4) {
    float4 acc = *reinterpret_cast<const float4 *>(partials + i);
    for (int s = 1; s < ksplit; ++s) {
      const float4 v = *reinterpret_cast<const float4 *>(partials + (long long)s * count + i);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    float4 o = *reinterpret_cast<float4 *>(dW + i);
    o.x += acc.x; o.y += acc.y; o.z += acc.z; o.w += acc.w;
    *reinterpret_cast<float4 *>(dW + i) = o;
  }
}

// deterministic bias gradient, db[co] = sum over pixels of dz, in two fixed-order stages: kBiasChunks blocks per
// output plane reduce a contiguous span each (fixed tree), then one thread per channel adds the chunk sums in order
constexpr int kBiasChunks = 64;
__global__ void __launch_bounds__(256) bias_grad_stage1_kernel(const __nv_bfloat16 *__restrict__ dz, long long img_stride, int n, int hw,
                                                               float *__restrict__ part /*[chunks][planes*8]*/, int c) {
  const int pl = blockIdx.y;
  const long long total = (long long)n * hw;
  const long long per = (total + gridDim.x - 1) / gridDim.x;
  const long long i0 = (long long)blockIdx.x * per, i1 = min(total, i0 + per);
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
  for (long long i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
    const long long img = i / hw, off = i - img * hw;
    const Vec8f v = load8(dz + img * img_stride + ((long long)pl * hw + off) * 8);
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] += v.v[k];
  }
  __shared__ float red[256][8];
#pragma unroll
  for (int k = 0; k < 8; ++k) red[threadIdx.x][k] = acc[k];
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) {
#pragma unroll
      for (int k = 0; k < 8; ++k) red[threadIdx.x][k] += red[threadIdx.x + s][k];
    }
    __syncthreads();
  }
  if (threadIdx.x < 8) part[(long long)blockIdx.x * c + pl * 8 + threadIdx.x] = red[0][threadIdx.x];
}
__global__ void bias_grad_stage2_kernel(const float *__restrict__ part, int chunks, int c, float *__restrict__ db) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c) return;
  float s = 0.f;
  for (int k = 0; k < chunks; ++k) s += part[(long long)k * c + i];
  db[i] += s;
}

bool wgrad_tc_applicable(int kh, int kw, int cin, int cout, int ups, int h, int w) {
  if (ups || kh < 1 || kw < 1 || kh > 3 || kw > 3) return false;
  // measured on B200 (tools/train_bench.py, per-layer profile): faster than the mma.sync kernels from 64 output channels
  // on (0.58 vs 0.84 ms over the six such layers of the default net at batch 256), slower at N = 32
  if (cin < 64 || (cin % 64) || (cout % 8) || cout < 64) return false;
  if ((cin / 8) % 16 && (cout % 8)) return false;
  const int mp = ((cin / 8) % 16 == 0) ? 16 : 8;
  if (mp == 16 && (cout % 16)) return false;     // M = 128 needs N % 16 == 0
  return h >= 1 && w >= 1;
}

// tiling / chunking of one layer: fills everything of `p` that does not depend on pointers
static int wgrad_tc_plan(int kh, int kw, int pad_top, int pad_left, int cin, int cout, int n, int H, int W, WgTcParams *pp) {
  WgTcParams &p = *pp;
  p = WgTcParams{};
  p.kh = kh; p.kw = kw; p.pt = pad_top; p.pl = pad_left; p.cin = cin; p.cout = cout;
  p.mp = ((cin / 8) % 16 == 0) ? 16 : 8;
  p.n_mblocks = (cin / 8) / p.mp;
  // N-chunk: kw accumulators of nc columns must fit the 512 TMEM columns; nc <= 256
  int chunks = 1;
  const int gran = p.mp == 16 ? 16 : 8;
  while (true) {
    if (cout % chunks == 0) {
      const int nc = cout / chunks;
      if (nc % gran == 0 && nc <= 256 && kw * nc <= 512) break;
    }
    if (++chunks > cout / 8) { set_error("wgrad_tc: no N-chunking for this layer"); return 1; }
  }
  p.n_nchunks = chunks;
  p.nc = cout / chunks;
  p.tiles_x = (W + kWtTW - 1) / kWtTW;
  const size_t budget = 200 * 1024;
  p.th = 8;
  for (;;) {
    p.a_pitch = (uint32_t)(kWtTW + kw - 1) * 16u;
    p.a_plane = p.a_pitch * (uint32_t)p.th;
    p.b_pitch = (uint32_t)kWtTW * 16u;
    p.b_plane = p.b_pitch * (uint32_t)p.th;
    p.a_bytes = (p.a_plane * (uint32_t)p.mp + 127u) & ~127u;
    p.b_bytes = (p.b_plane * (uint32_t)(p.nc >> 3) + 127u) & ~127u;
    p.stages = (int)std::min<size_t>(kWtMaxStages, budget / (p.a_bytes + p.b_bytes));
    if (p.stages >= 3 || p.th == 1) break;
    p.th >>= 1;
  }
  if (p.stages < 2) { set_error("wgrad_tc: smem budget exceeded"); return 1; }
  p.tiles_y = (H + p.th - 1) / p.th;
  p.num_tiles = n * p.tiles_x * p.tiles_y;
  const int base = p.n_mblocks * p.n_nchunks * kh;
  p.ksplit = std::max(1, std::min(p.num_tiles, 148 / std::max(1, base)));
  return 0;
}

// scratch floats needed for (kh, kw, cin, cout) at n images of h x w (0 = not applicable)
size_t wgrad_tc_scratch_floats(int kh, int kw, int cin, int cout, int n, int h, int w) {
  if (!wgrad_tc_applicable(kh, kw, cin, cout, 0, h, w)) return 0;
  WgTcParams p;
  if (wgrad_tc_plan(kh, kw, (kh - 1) / 2, (kw - 1) / 2, cin, cout, n, h, w, &p)) return 0;
  return (size_t)p.ksplit * kh * kw * cin * cout + (size_t)kBiasChunks * cout;
}

int launch_wgrad_tc(View<const __nv_bfloat16> a_in, View<const __nv_bfloat16> dz, int kh, int kw, int pad_top, int pad_left,
                    int cin, int cout, float *dW, float *db, float *scratch, size_t scratch_floats, int *status,
                    cudaStream_t st) {
  if (a_in.planes * 8 != cin || dz.planes * 8 != cout || a_in.h != dz.h || a_in.w != dz.w) {
    set_error("wgrad_tc: views do not match the layer");
    return 1;
  }
  if (a_in.img_stride != (long long)a_in.planes * a_in.h * a_in.w * 8 || dz.img_stride != (long long)dz.planes * dz.h * dz.w * 8) {
    set_error("wgrad_tc: dense tensors only");
    return 1;
  }
  WgTcParams p;
  if (wgrad_tc_plan(kh, kw, pad_top, pad_left, cin, cout, dz.n, dz.h, dz.w, &p)) return 1;
  const int base = p.n_mblocks * p.n_nchunks * kh;
  const size_t w_count = (size_t)kh * kw * cin * cout;
  if ((size_t)p.ksplit * w_count + (size_t)kBiasChunks * cout > scratch_floats) { set_error("wgrad_tc: scratch buffer too small"); return 1; }
  p.partials = scratch;
  p.status = status;
  CUtensorMap map_a, map_d;
  if (tc_encode_map_4d(a_in.ptr, a_in.w, a_in.h, a_in.planes, a_in.n, kWtTW + kw - 1, p.th, p.mp, &map_a)) return 1;
  if (tc_encode_map_4d(dz.ptr, dz.w, dz.h, dz.planes, dz.n, kWtTW, p.th, p.nc >> 3, &map_d)) return 1;
  const size_t smem = (size_t)p.stages * (p.a_bytes + p.b_bytes) + sizeof(WtBarriers) + 1024;
  static PerDeviceOnce attr;
  if (const int dev = attr.pending(); dev >= 0) {
    OCTSEG_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr.mark(dev);
  }
  dim3 grid(p.ksplit, base);
  wgrad_tc_kernel<<<grid, kWtThreads, smem, st>>>(map_a, map_d, p);
  OCTSEG_CUDA(cudaGetLastError());
  const long long cnt = (long long)w_count;
  wgrad_tc_reduce_kernel<<<(unsigned)std::min<long long>((cnt / 4 + 255) / 256, 148 * 4), 256, 0, st>>>(scratch, p.ksplit, cnt, dW);
  OCTSEG_CUDA(cudaGetLastError());
  if (db) {
    float *part = scratch + (size_t)p.ksplit * w_count;
    bias_grad_stage1_kernel<<<dim3(kBiasChunks, dz.planes), 256, 0, st>>>(dz.ptr, dz.img_stride, dz.n, dz.h * dz.w, part, cout);
    OCTSEG_CUDA(cudaGetLastError());
    bias_grad_stage2_kernel<<<(cout + 127) / 128, 128, 0, st>>>(part, kBiasChunks, cout, db);
    OCTSEG_CUDA(cudaGetLastError());
  }
  return 0;
}

}  // namespace octseg
