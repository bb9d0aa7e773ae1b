// Training kernels: batch-statistics BatchNorm (fwd/bwd), fused head + weighted CE,
// pool backward, weight gradient, weight transforms, Adam.  Memory-bound ops are vectorised
// over the blocked [N][C/8][H][W][8] layout (one 16/32 B vector per thread access),
// reductions are warp-shuffle -> shared -> one double/float atomic per block.
#include "train_kernels.cuh"

namespace octseg {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// address of vector `v` (0 .. n*h*w-1) of plane `pl` in a view
template <typename T>
__device__ __forceinline__ const T *vec_ptr(const View<const T> &t, int pl, long long v) {
  const long long hw = (long long)t.h * t.w;
  const long long img = v / hw, off = v - img * hw;
  return t.ptr + img * t.img_stride + ((long long)pl * hw + off) * 8;
}
template <typename T>
__device__ __forceinline__ T *vec_ptr_w(const View<T> &t, int pl, long long v) {
  const long long hw = (long long)t.h * t.w;
  const long long img = v / hw, off = v - img * hw;
  return t.ptr + img * t.img_stride + ((long long)pl * hw + off) * 8;
}

// ---------------------------------------------------------------------------------
// per-channel reductions: grid = (chunks, planes); each block reduces 16 values
// ---------------------------------------------------------------------------------
constexpr int kRedThreads = 256;
constexpr int kRedPerThread = 32;   // vectors per thread -> bounded fp32 partial sums

__device__ __forceinline__ void block_reduce16_to_double(float (&acc)[16], double *dst_lo, double *dst_hi, int c0) {
  __shared__ float red[kRedThreads / 32][16];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = warp_sum(acc[i]);
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < 16; ++i) red[warp][i] = acc[i];
  }
  __syncthreads();
  if (threadIdx.x < 16) {
    double s = 0;
    for (int w = 0; w < kRedThreads / 32; ++w) s += (double)red[w][threadIdx.x];
    if (threadIdx.x < 8) atomicAdd(dst_lo + c0 + threadIdx.x, s);
    else atomicAdd(dst_hi + c0 + threadIdx.x - 8, s);
  }
}

// grid = (chunks, planes, images): every block walks a contiguous span of one (image, plane), so
// the address is base + linear offset (no 64-bit divisions in the loop)
constexpr int kSpanVecs = 8192;   // vectors per block span

template <typename T>
__global__ void __launch_bounds__(kRedThreads) bn_stats_kernel(View<const T> z, double *sums, int c) {
  const int pl = blockIdx.y, img = blockIdx.z;
  const int hw = z.h * z.w;
  const T *base = z.ptr + (long long)img * z.img_stride + (long long)pl * hw * 8;
  const int v_end = min(hw, (int)(blockIdx.x + 1) * kSpanVecs);
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = 0.f;
  for (int v = blockIdx.x * kSpanVecs + threadIdx.x; v < v_end; v += kRedThreads) {
    const Vec8f x = load8(base + (long long)v * 8);
#pragma unroll
    for (int i = 0; i < 8; ++i) { acc[i] += x.v[i]; acc[8 + i] = fmaf(x.v[i], x.v[i], acc[8 + i]); }
  }
  block_reduce16_to_double(acc, sums, sums + c, pl * 8);
}

template <typename T>
int launch_bn_stats(View<const T> z, double *sums, cudaStream_t st) {
  const int c = z.planes * 8;   // `sums` is zeroed by the caller (one memset per train step)
  const int hw = z.h * z.w;
  dim3 grid((hw + kSpanVecs - 1) / kSpanVecs, z.planes, z.n);
  bn_stats_kernel<T><<<grid, kRedThreads, 0, st>>>(z, sums, c);
  OCTSEG_CUDA(cudaGetLastError());
  return 0;
}

__global__ void bn_finalize_kernel(const double *sums, long long count, int c, float eps, float momentum,
                                   const float *gamma, const float *beta, float *moving_mean,
                                   float *moving_var, float *mean, float *invstd, float *scale, float *shift) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c) return;
  const double m = sums[i] / (double)count;
  double var = sums[c + i] / (double)count - m * m;
  if (var < 0) var = 0;
  const float mf = (float)m, vf = (float)var;
  const float inv = 1.0f / sqrtf(vf + eps);
  mean[i] = mf;
  invstd[i] = inv;
  const float s = inv * gamma[i];
  scale[i] = s;
  shift[i] = beta[i] - mf * s;
  // Keras fused BN: moving <- moving*momentum + batch*(1-momentum); Bessel-corrected variance
  const float unbiased = (float)(var * ((double)count / (double)(count > 1 ? count - 1 : 1)));
  moving_mean[i] = moving_mean[i] * momentum + mf * (1.f - momentum);
  moving_var[i] = moving_var[i] * momentum + unbiased * (1.f - momentum);
}

int launch_bn_finalize(const double *sums, long long count, int c, float eps, float momentum,
                       const float *gamma, const float *beta, float *moving_mean, float *moving_var,
                       float *mean, float *invstd, float *scale, float *shift, cudaStream_t st) {
  bn_finalize_kernel<<<(c + 127) / 128, 128, 0, st>>>(sums, count, c, eps, momentum, gamma, beta, moving_mean,
                                                      moving_var, mean, invstd, scale, shift);
  OCTSEG_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------
// BatchNorm forward, second half, in ONE pass: batch mean / variance from the per-channel sums (accumulated by the
// convolution's epilogue or by bn_stats_kernel), a = relu(z * scale + shift) [* dropout multiplier], and for the
// encoder-final blocks the 2x2 max-pooled copy.  Block (0,0,0) also publishes mean / invstd / scale / shift for
// the backward pass and performs the Keras moving-statistics update (momentum 0.99, Bessel-corrected variance).
// ---------------------------------------------------------------------------------
struct BnChan { float mean, inv, scale, shift; };
__device__ __forceinline__ BnChan bn_channel(const double *sums, int c, int ch, double inv_count, float eps,
                                             const float *gamma, const float *beta, double *var_out = nullptr) {
  const double m = sums[ch] * inv_count;
  double var = sums[c + ch] * inv_count - m * m;
  if (var < 0) var = 0;
  if (var_out) *var_out = var;
  BnChan r;
  r.mean = (float)m;
  r.inv = 1.0f / sqrtf((float)var + eps);
  r.scale = r.inv * gamma[ch];
  r.shift = beta[ch] - r.mean * r.scale;
  return r;
}

template <typename T, int POOL>
__global__ void __launch_bounds__(256) bn_finalize_apply_kernel(
    View<const T> z, const double *__restrict__ sums, long long count, float eps, float momentum,
    const float *__restrict__ gamma, const float *__restrict__ beta, float *moving_mean, float *moving_var,
    float *mean, float *invstd, float *scale, float *shift, const T *__restrict__ mask, View<T> a, View<T> pooled) {
  const int c = z.planes * 8;
  const double inv_count = 1.0 / (double)count;
  if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) {
    for (int i = threadIdx.x; i < c; i += blockDim.x) {
      double var;
      const BnChan b = bn_channel(sums, c, i, inv_count, eps, gamma, beta, &var);
      mean[i] = b.mean; invstd[i] = b.inv; scale[i] = b.scale; shift[i] = b.shift;
      const float unbiased = (float)(var * ((double)count / (double)(count > 1 ? count - 1 : 1)));
      moving_mean[i] = moving_mean[i] * momentum + b.mean * (1.f - momentum);
      moving_var[i] = moving_var[i] * momentum + unbiased * (1.f - momentum);
    }
  }
  const int pl = blockIdx.y, img = blockIdx.z;
  const int hw = z.h * z.w;
  const T *zb = z.ptr + (long long)img * z.img_stride + (long long)pl * hw * 8;
  T *ab = a.ptr + (long long)img * a.img_stride + (long long)pl * hw * 8;
  const T *mb = mask ? mask + ((long long)img * z.planes + pl) * hw * 8 : nullptr;
  // the 8 channels of this plane: computed once per block (double math + rsqrt), broadcast through shared memory
  __shared__ float s_sc[8], s_sh[8];
  if (threadIdx.x < 8) {
    const BnChan b = bn_channel(sums, c, pl * 8 + threadIdx.x, inv_count, eps, gamma, beta);
    s_sc[threadIdx.x] = b.scale; s_sh[threadIdx.x] = b.shift;
  }
  __syncthreads();
  float sc[8], sh[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { sc[k] = s_sc[k]; sh[k] = s_sh[k]; }
  if constexpr (POOL == 0) {
    constexpr int kU = 4;
    const int stride = gridDim.x * blockDim.x;
    for (int v0 = blockIdx.x * blockDim.x + threadIdx.x; v0 < hw; v0 += stride * kU) {
      Raw8<T> rx[kU], rm[kU];
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int v = v0 + u * stride;
        if (v < hw) {
          rx[u] = load_raw8(zb + (long long)v * 8);
          if (mb) rm[u] = load_raw8(mb + (long long)v * 8);
        }
      }
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int v = v0 + u * stride;
        if (v >= hw) break;
        Vec8f x = cvt8(rx[u]);
        Vec8f mk = zero8();
        if (mb) mk = cvt8(rm[u]);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          x.v[k] = fmaxf(fmaf(x.v[k], sc[k], sh[k]), 0.f);
          if (mb) x.v[k] *= mk.v[k];
        }
        store8(ab + (long long)v * 8, x);
      }
    }
  } else {
    // one thread = one 2x2 window: four activations out, their maximum to the pooled tensor
    const int W = z.w, Wo = W >> 1, Ho = z.h >> 1;
    T *pb = pooled.ptr + (long long)img * pooled.img_stride + (long long)pl * Ho * Wo * 8;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < Ho * Wo; i += gridDim.x * blockDim.x) {
      const int x = i % Wo, y = i / Wo;
      const long long o00 = ((long long)(2 * y) * W + 2 * x) * 8, o10 = o00 + (long long)W * 8;
      Vec8f q[4] = {load8(zb + o00), load8(zb + o00 + 8), load8(zb + o10), load8(zb + o10 + 8)};
      Vec8f m;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
#pragma unroll
        for (int j = 0; j < 4; ++j) q[j].v[k] = fmaxf(fmaf(q[j].v[k], sc[k], sh[k]), 0.f);
        m.v[k] = fmaxf(fmaxf(q[0].v[k], q[1].v[k]), fmaxf(q[2].v[k], q[3].v[k]));
      }
      store8(ab + o00, q[0]); store8(ab + o00 + 8, q[1]); store8(ab + o10, q[2]); store8(ab + o10 + 8, q[3]);
      // the pooled tensor is the maximum of the STORED (rounded) activations; rounding is monotonic, so max-then-round == round-then-max
      store8(pb + (long long)i * 8, m);
    }
  }
}

template <typename T>
int launch_bn_finalize_apply(View<const T> z, const double *sums, long long count, float eps, float momentum,
                             const float *gamma, const float *beta, float *moving_mean, float *moving_var, float *mean,
                             float *invstd, float *scale, float *shift, const T *mask, View<T> a, View<T> pooled,
                             cudaStream_t st) {
  const int hw = z.h * z.w;
  if (pooled.ptr) {
    if (mask || (z.h & 1) || (z.w & 1)) { set_error("bn_finalize_apply: pooled variant needs even dims and no dropout"); return 1; }
    dim3 grid(std::max(1, std::min((hw / 4 + 2047) / 2048, 64)), z.planes, z.n);     // >= 8 windows (32 vectors) per thread
    bn_finalize_apply_kernel<T, 1><<<grid, 256, 0, st>>>(z, sums, count, eps, momentum, gamma, beta, moving_mean, moving_var,
                                                         mean, invstd, scale, shift, mask, a, pooled);
  } else {
    dim3 grid(std::max(1, std::min((hw + 4095) / 4096, 64)), z.planes, z.n);         // >= 16 vectors per thread
    bn_finalize_apply_kernel<T, 0><<<grid, 256, 0, st>>>(z, sums, count, eps, momentum, gamma, beta, moving_mean, moving_var,
                                                         mean, invstd, scale, shift, mask, a, pooled);
  }
  OCTSEG_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) bn_apply_relu_kernel(View<const T> z, const float *__restrict__ scale,
                                                            const float *__restrict__ shift,
                                                            const T *__restrict__ mask, View<T> a) {
  const int pl = blockIdx.y, img = blockIdx.z;
  const int hw = z.h * z.w;
  const T *zb = z.ptr + (long long)img * z.img_stride + (long long)pl * hw * 8;
  T *ab = a.ptr + (long long)img * a.img_stride + (long long)pl * hw * 8;
  const T *mb = mask ? mask + ((long long)img * z.planes + pl) * hw * 8 : nullptr;
  float sc[8], sh[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { sc[k] = scale[pl * 8 + k]; sh[k] = shift[pl * 8 + k]; }
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < hw; v += gridDim.x * blockDim.x) {
    Vec8f x = load8(zb + (long long)v * 8);
#pragma unroll
    for (int k = 0; k < 8; ++k) x.v[k] = fmaxf(fmaf(x.v[k], sc[k], sh[k]), 0.f);
    if (mb) {
      const Vec8f mk = load8(mb + (long long)v * 8);
#pragma unroll
      for (int k = 0; k < 8; ++k) x.v[k] *= mk.v[k];
    }
    store8(ab + (long long)v * 8, x);
  }
}

template <typename T>
int launch_bn_apply_relu(View<const T> z, const float *scale, const float *shift, const T *mask,
                         View<T> a, cudaStream_t st) {
  const int hw = z.h * z.w;
  dim3 grid(std::max(1, std::min((hw + 1023) / 1024, 64)), z.planes, z.n);
  bn_apply_relu_kernel<T><<<grid, 256, 0, st>>>(z, scale, shift, mask, a);
  OCTSEG_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t hash_u64(unsigned long long x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
  return (uint32_t)x;
}

template <typename T>
__global__ void dropout_mask_kernel(const uint8_t *__restrict__ mask_nhwc, unsigned long long seed,
                                    const StepState *__restrict__ state, float rate,
                                    int n, int c, int h, int w, T *__restrict__ out) {
  const long long total = (long long)n * c * h * w;
  if (state) seed += state->step * 0x51ED27ULL;     // per-step stream, read on the device (graph-replay safe)
  const float keep_scale = 1.f / (1.f - rate);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    // i indexes the blocked layout [n][c/8][h][w][8]
    const int c8 = (int)(i & 7);
    long long t = i >> 3;
    const int x = (int)(t % w); t /= w;
    const int y = (int)(t % h); t /= h;
    const int cg = (int)(t % (c / 8));
    const int b = (int)(t / (c / 8));
    const int ch = cg * 8 + c8;
    bool keep;
    if (mask_nhwc) keep = mask_nhwc[(((long long)b * h + y) * w + x) * c + ch] != 0;
    else keep = (hash_u64(seed + (unsigned long long)i * 0x9E3779B97F4A7C15ULL) * 2.3283064365386963e-10f) >= rate;
    out[i] = (T)(keep ? keep_scale : 0.f);
  }
}

template <typename T>
int launch_dropout_mask(const uint8_t *mask_nhwc, unsigned long long seed, const StepState *state, float rate,
                        int n, int c, int h, int w, T *out, cudaStream_t st) {
  const long long total = (long long)n * c * h * w;
  unsigned grid = (unsigned)std::min<long long>((total + 255) / 256, 148 * 8);
  dropout_mask_kernel<T><<<grid, 256, 0, st>>>(mask_nhwc, seed, state, rate, n, c, h, w, out);
  OCTSEG_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------
// fused head: logits -> softmax -> weighted CE (reference common/custom_losses.py:27-35)
//   p /= sum(p); p = clip(p, 1e-7, 1-1e-7); loss = -w_t*log(p_t)
//   dlogit_j = w_t*(p_j - [j==t]) * inv_denominator, zero where the clip is active
// ---------------------------------------------------------------------------------
constexpr int kHeadMaxK = 16;

// K and the number of input planes are compile-time so that every per-thread partial
// (d_wgt[cin][K], d_bias[K], loss) lives in registers for the whole grid-stride loop and is
// reduced ONCE per block (warp shuffles -> smem -> one global atomic per value per block).
template <typename T, int K, int PL>
__global__ void __launch_bounds__(256, (PL * 8 * K <= 32) ? 2 : 1) head_loss_kernel(View<const T> a, const float *__restrict__ wgt,
                                                        const float *__restrict__ bias,
                                                        const uint8_t *__restrict__ labels,
                                                        const float *__restrict__ class_w, float inv_den,
                                                        View<T> da, float *__restrict__ d_wgt,
                                                        float *__restrict__ d_bias, double *loss_acc) {
  constexpr int CIN = PL * 8;
  __shared__ float s_w[CIN * K], s_b[K], s_cw[K];
  __shared__ float s_red[8][CIN * K + K + 1];
  for (int i = threadIdx.x; i < CIN * K; i += blockDim.x) s_w[i] = wgt[i];
  for (int i = threadIdx.x; i < K; i += blockDim.x) { s_b[i] = bias[i]; s_cw[i] = class_w[i]; }
  __syncthreads();
  // grid = (blocks per image, images): no 64-bit division per pixel (it cost more than the softmax)
  const int hw = a.h * a.w;
  const long long b = blockIdx.y;
  const uint8_t *lab = labels + b * hw;
  float p_dw[CIN * K], p_db[K], p_loss = 0.f;
#pragma unroll
  for (int i = 0; i < CIN * K; ++i) p_dw[i] = 0.f;
#pragma unroll
  for (int k = 0; k < K; ++k) p_db[k] = 0.f;
  // kU pixels per thread and iteration with every load issued before the first use: at 128 registers the kernel runs 16
  // warps per SM, and one 16-byte load in flight per thread left the memory system idle (ncu: DRAM 16 %)
  constexpr int kU = (PL == 1) ? 4 : 2;
  const int stride = gridDim.x * blockDim.x;
  for (int off0 = blockIdx.x * blockDim.x + threadIdx.x; off0 < hw; off0 += stride * kU) {
    Raw8<T> rv[kU][PL];
    int tt[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int off = off0 + u * stride;
      tt[u] = 0;
      if (off < hw) {
#pragma unroll
        for (int pl = 0; pl < PL; ++pl) rv[u][pl] = load_raw8(a.ptr + b * a.img_stride + ((long long)pl * hw + off) * 8);
        tt[u] = lab[off];
      }
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int off = off0 + u * stride;
      if (off >= hw) break;
      Vec8f v[PL];
      float z[K];
#pragma unroll
      for (int pl = 0; pl < PL; ++pl) v[pl] = cvt8(rv[u][pl]);
      const int t = tt[u];
#pragma unroll
      for (int k = 0; k < K; ++k) z[k] = s_b[k];
#pragma unroll
      for (int c = 0; c < CIN; ++c)
#pragma unroll
        for (int k = 0; k < K; ++k) z[k] = fmaf(v[c >> 3].v[c & 7], s_w[c * K + k], z[k]);
      float mx = z[0];
#pragma unroll
      for (int k = 1; k < K; ++k) mx = fmaxf(mx, z[k]);
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < K; ++k) { z[k] = expf(z[k] - mx); s += z[k]; }
      const float inv = 1.f / s;
      float S = 0.f;
#pragma unroll
      for (int k = 0; k < K; ++k) { z[k] *= inv; S += z[k]; }
      // the loss renormalises its input (custom_losses.py:30): S = 1 +- a few ulp, so one reciprocal serves all classes
      const float rS = 1.f / S;
      float pt = 0.f;
#pragma unroll
      for (int k = 0; k < K; ++k) { z[k] *= rS; if (k == t) pt = z[k]; }
      const float wt = (t < K) ? s_cw[t] : 0.f;
      const bool active = (pt >= 1e-7f) && (pt <= 1.f - 1e-7f);
      p_loss += -wt * logf(fminf(fmaxf(pt, 1e-7f), 1.f - 1e-7f)) * inv_den;
      const float gs = active ? wt * inv_den : 0.f;
      float dl[K];
#pragma unroll
      for (int k = 0; k < K; ++k) {
        dl[k] = gs * (z[k] - (k == t ? 1.f : 0.f));
        p_db[k] += dl[k];
      }
#pragma unroll
      for (int pl = 0; pl < PL; ++pl) {
        Vec8f g;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          float acc = 0.f;
#pragma unroll
          for (int k = 0; k < K; ++k) {
            acc = fmaf(dl[k], s_w[(pl * 8 + c) * K + k], acc);
            p_dw[(pl * 8 + c) * K + k] = fmaf(v[pl].v[c], dl[k], p_dw[(pl * 8 + c) * K + k]);
          }
          g.v[c] = acc;
        }
        store8(da.ptr + b * da.img_stride + ((long long)pl * hw + off) * 8, g);
      }
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < CIN * K; ++i) {
    const float r = warp_sum(p_dw[i]);
    if (lane == 0) s_red[warp][i] = r;
  }
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const float r = warp_sum(p_db[k]);
    if (lane == 0) s_red[warp][CIN * K + k] = r;
  }
  p_loss = warp_sum(p_loss);
  if (lane == 0) s_red[warp][CIN * K + K] = p_loss;
  __syncthreads();
  for (int i = threadIdx.x; i < CIN * K + K + 1; i += blockDim.x) {
    float r = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) r += s_red[w][i];
    if (i < CIN * K) atomicAdd(&d_wgt[i], r);
    else if (i < CIN * K + K) atomicAdd(&d_bias[i - CIN * K], r);
    else atomicAdd(loss_acc, (double)r);
  }
}

template <typename T, int K, int PL>
static int launch_head_loss_kp(View<const T> a, const float *wgt, const float *bias, const uint8_t *labels,
                               const float *class_w, float inv_den, View<T> da, float *d_wgt, float *d_bias,
                               double *loss_acc, cudaStream_t st) {
  // ~148 * 4 blocks in total (every block reduces its partial sums once: few, long-lived blocks)
  const int hw = a.h * a.w;
  const int per_img = std::max(1, std::min((hw + 255) / 256, (148 * 4 + a.n - 1) / a.n));
  dim3 grid(per_img, a.n);
  head_loss_kernel<T, K, PL><<<grid, 256, 0, st>>>(a, wgt, bias, labels, class_w, inv_den, da, d_wgt, d_bias,
                                                   loss_acc);
  OCTSEG_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------
// Generic head (any channel count): the register-resident d(head weights) of head_loss_kernel does not scale past 16
// input channels, so wide heads run two passes: (A) logits / softmax / loss / dlogits -> scratch, d(input), d(bias);
// (B) d(weights)[ci][k] = sum_px a[px][ci] * dlogit[px][k], one block column per input plane.
// ---------------------------------------------------------------------------------
template <typename T, int K>
__global__ void __launch_bounds__(256) head_loss_wide_kernel(View<const T> a, const float *__restrict__ wgt,
                                                             const float *__restrict__ bias, int cin,
                                                             const uint8_t *__restrict__ labels,
                                                             const float *__restrict__ class_w, float inv_den, View<T> da,
                                                             float *__restrict__ dlog, float *__restrict__ d_bias,
                                                             double *loss_acc) {
  extern __shared__ float s_w[];      // [cin][K] + bias[K] + class_w[K]
  for (int i = threadIdx.x; i < cin * K; i += blockDim.x) s_w[i] = wgt[i];
  for (int i = threadIdx.x; i < K; i += blockDim.x) { s_w[cin * K + i] = bias[i]; s_w[cin * K + K + i] = class_w[i]; }
  __syncthreads();
  const float *s_b = s_w + cin * K, *s_cw = s_b + K;
  const int planes = cin / 8;
  const long long hw = (long long)a.h * a.w, total = (long long)a.n * hw;
  float p_db[K], p_loss = 0.f;
#pragma unroll
  for (int k = 0; k < K; ++k) p_db[k] = 0.f;
  for (long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x; pix < total; pix += (long long)gridDim.x * blockDim.x) {
    const long long off = pix % hw;
    const int b = (int)(pix / hw);
    float z[K];
#pragma unroll
    for (int k = 0; k < K; ++k) z[k] = s_b[k];
    for (int pl = 0; pl < planes; ++pl) {
      const Vec8f v = load8(a.ptr + b * a.img_stride + ((long long)pl * hw + off) * 8);
#pragma unroll
      for (int c = 0; c < 8; ++c)
#pragma unroll
        for (int k = 0; k < K; ++k) z[k] = fmaf(v.v[c], s_w[(pl * 8 + c) * K + k], z[k]);
    }
    float mx = z[0];
#pragma unroll
    for (int k = 1; k < K; ++k) mx = fmaxf(mx, z[k]);
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k) { z[k] = expf(z[k] - mx); s += z[k]; }
    const float inv = 1.f / s;
    float S = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k) { z[k] *= inv; S += z[k]; }
    const int t = labels[pix];
    float pt = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k) { z[k] = z[k] / S; if (k == t) pt = z[k]; }
    const float wt = (t < K) ? s_cw[t] : 0.f;
    const bool active = (pt >= 1e-7f) && (pt <= 1.f - 1e-7f);
    p_loss += -wt * logf(fminf(fmaxf(pt, 1e-7f), 1.f - 1e-7f)) * inv_den;
    float dl[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
      dl[k] = active ? wt * (z[k] - (k == t ? 1.f : 0.f)) * inv_den : 0.f;
      p_db[k] += dl[k];
      dlog[pix * K + k] = dl[k];
    }
    for (int pl = 0; pl < planes; ++pl) {
      Vec8f g;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < K; ++k) acc = fmaf(dl[k], s_w[(pl * 8 + c) * K + k], acc);
        g.v[c] = acc;
      }
      store8(da.ptr + b * da.img_stride + ((long long)pl * hw + off) * 8, g);
    }
  }
  __shared__ float s_red[8][K + 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const float r = warp_sum(p_db[k]);
    if (lane == 0) s_red[warp][k] = r;
  }
  p_loss = warp_sum(p_loss);
  if (lane == 0) s_red[warp][K] = p_loss;
  __syncthreads();
  if (threadIdx.x <= K) {
    float r = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) r += s_red[w][threadIdx.x];
    if (threadIdx.x < K) atomicAdd(&d_bias[threadIdx.x], r);
    else atomicAdd(loss_acc, (double)r);
  }
}

template <typename T, int K>
__global__ void __launch_bounds__(256) head_wgrad_wide_kernel(View<const T> a, const float *__restrict__ dlog,
                                                              float *__restrict__ d_wgt) {
  const int pl = blockIdx.y;
  const long long hw = (long long)a.h * a.w, total = (long long)a.n * hw;
  float p[8][K];
#pragma unroll
  for (int c = 0; c < 8; ++c)
#pragma unroll
    for (int k = 0; k < K; ++k) p[c][k] = 0.f;
  for (long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x; pix < total; pix += (long long)gridDim.x * blockDim.x) {
    const long long off = pix % hw;
    const int b = (int)(pix / hw);
    const Vec8f v = load8(a.ptr + b * a.img_stride + ((long long)pl * hw + off) * 8);
    float dl[K];
#pragma unroll
    for (int k = 0; k < K; ++k) dl[k] = dlog[pix * K + k];
#pragma unroll
    for (int c = 0; c < 8; ++c)
#pragma unroll
      for (int k = 0; k < K; ++k) p[c][k] = fmaf(v.v[c], dl[k], p[c][k]);
  }
  __shared__ float s_red[8][8 * K];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int c = 0; c < 8; ++c)
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const float r = warp_sum(p[c][k]);
      if (lane == 0) s_red[warp][c * K + k] = r;
    }
  __syncthreads();
  for (int i = threadIdx.x; i < 8 * K; i += blockDim.x) {
    float r = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) r += s_red[w][i];
    atomicAdd(&d_wgt[(pl * 8) * K + i], r);
  }
}

template <typename T, int K>
static int launch_head_loss_wide(View<const T> a, const float *wgt, const float *bias, int cin, const uint8_t *labels,
                                 const float *class_w, float inv_den, View<T> da, float *d_wgt, float *d_bias,
                                 double *loss_acc, float *dlog_scratch, cudaStream_t st) {
  if (!dlog_scratch) { set_error("training head with more than 16 input channels needs the dlogits scratch buffer"); return 1; }
  const long long total = (long long)a.n * a.h * a.w;
  const unsigned grid = (unsigned)std::min<long long>((total + 255) / 256, 148 * 8);
  const size_t smem = (size_t)(cin * K + 2 * K) * sizeof(float);
  if (smem > 48 * 1024) { set_error("training head: cin * num_classes too large"); return 1; }
  head_loss_wide_kernel<T, K><<<grid, 256, smem, st>>>(a, wgt, bias, cin, labels, class_w, inv_den, da, dlog_scratch, d_bias,
                                                       loss_acc);
  OCTSEG_CUDA(cudaGetLastError());
  dim3 g2((unsigned)std::min<long long>((total + 255) / 256, 148), cin / 8);
  head_wgrad_wide_kernel<T, K><<<g2, 256, 0, st>>>(a, dlog_scratch, d_wgt);
  OCTSEG_CUDA(cudaGetLastError());
  return 0;
}

template <typename T, int K>
static int launch_head_loss_k(View<const T> a, const float *wgt, const float *bias, int cin, const uint8_t *labels,
                              const float *class_w, float inv_den, View<T> da, float *d_wgt, float *d_bias,
                              double *loss_acc, float *dlog_scratch, cudaStream_t st) {
  if (cin == 8) return launch_head_loss_kp<T, K, 1>(a, wgt, bias, labels, class_w, inv_den, da, d_wgt, d_bias, loss_acc, st);
  if (cin == 16 && K <= 8) return launch_head_loss_kp<T, K, 2>(a, wgt, bias, labels, class_w, inv_den, da, d_wgt, d_bias, loss_acc, st);
  return launch_head_loss_wide<T, K>(a, wgt, bias, cin, labels, class_w, inv_den, da, d_wgt, d_bias, loss_acc, dlog_scratch, st);
}

template <typename T>
int launch_head_loss(View<const T> a, const float *wgt, const float *bias, int cin, int K,
                     const uint8_t *labels, const float *class_w, float inv_denominator, View<T> da,
                     float *d_wgt, float *d_bias, double *loss_acc, float *dlog_scratch, cudaStream_t st) {
  switch (K) {
#define HK(k) case k: return launch_head_loss_k<T, k>(a, wgt, bias, cin, labels, class_w, inv_denominator, da, d_wgt, d_bias, loss_acc, dlog_scratch, st);
    HK(2) HK(3) HK(4) HK(5) HK(6) HK(7) HK(8)
#undef HK
  }
  set_error("training supports 2..8 classes");
  return 1;
}

// ---------------------------------------------------------------------------------
// BN + ReLU (+dropout multiplier) backward
// ---------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kRedThreads, 2) bn_bwd_reduce_kernel(
    View<const T> da, View<const T> z, const float *__restrict__ mean, const float *__restrict__ invstd,
    const float *__restrict__ gamma, const float *__restrict__ beta, const T *__restrict__ mask, double *sums,
    int c) {
  const int pl = blockIdx.y, img = blockIdx.z;
  const int hw = z.h * z.w;
  const T *zb = z.ptr + (long long)img * z.img_stride + (long long)pl * hw * 8;
  const T *gb = da.ptr + (long long)img * da.img_stride + (long long)pl * hw * 8;
  const T *mb = mask ? mask + ((long long)img * z.planes + pl) * hw * 8 : nullptr;
  const int v_end = min(hw, (int)(blockIdx.x + 1) * kSpanVecs);
  float mu[8], is[8], ga[8], be[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { mu[i] = mean[pl * 8 + i]; is[i] = invstd[pl * 8 + i]; ga[i] = gamma[pl * 8 + i]; be[i] = beta[pl * 8 + i]; }
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = 0.f;
  // 4 vectors per iteration with every load issued before the first use: the kernel runs at ~37 % occupancy (78
  // registers), so memory-level parallelism has to come from the instruction stream (ncu: DRAM 44 % -> see profiles/)
  constexpr int kU = 4;
  for (int v0 = blockIdx.x * kSpanVecs + threadIdx.x; v0 < v_end; v0 += kRedThreads * kU) {
    Raw8<T> rz[kU], rg[kU], rm[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int v = v0 + u * kRedThreads;
      if (v < v_end) {
        rz[u] = load_raw8(zb + (long long)v * 8);
        rg[u] = load_raw8(gb + (long long)v * 8);
        if (mb) rm[u] = load_raw8(mb + (long long)v * 8);
      }
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int v = v0 + u * kRedThreads;
      if (v >= v_end) break;
      const Vec8f zz = cvt8(rz[u]), g = cvt8(rg[u]);
      Vec8f mk = zero8();
      if (mb) mk = cvt8(rm[u]);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float gi = mb ? g.v[i] * mk.v[i] : g.v[i];
        const float zh = (zz.v[i] - mu[i]) * is[i];
        const float dy = (fmaf(ga[i], zh, be[i]) > 0.f) ? gi : 0.f;
        acc[i] += dy;
        acc[8 + i] = fmaf(dy, zh, acc[8 + i]);
      }
    }
  }
  block_reduce16_to_double(acc, sums, sums + c, pl * 8);
}

template <typename T>
int launch_bn_bwd_reduce(View<const T> da, View<const T> z, const float *mean, const float *invstd,
                         const float *gamma, const float *beta, const T *mask, double *sums,
                         cudaStream_t st) {
  const int c = z.planes * 8;   // `sums` is zeroed by the caller
  const int hw = z.h * z.w;
  dim3 grid((hw + kSpanVecs - 1) / kSpanVecs, z.planes, z.n);
  bn_bwd_reduce_kernel<T><<<grid, kRedThreads, 0, st>>>(da, z, mean, invstd, gamma, beta, mask, sums, c);
  OCTSEG_CUDA(cudaGetLastError());
  return 0;
}

template <typename T>
__global__ void __launch_bounds__(256, 2) bn_bwd_apply_kernel(
    View<const T> da, View<const T> z, const float *__restrict__ mean, const float *__restrict__ invstd,
    const float *__restrict__ gamma, const float *__restrict__ beta, const T *__restrict__ mask,
    const double *__restrict__ sums, long long count, View<T> dz, float *d_gamma, float *d_beta, int c) {
  if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0)
    for (int i = threadIdx.x; i < c; i += blockDim.x) { d_beta[i] = (float)sums[i]; d_gamma[i] = (float)sums[c + i]; }
  const int pl = blockIdx.y, img = blockIdx.z;
  const int hw = z.h * z.w;
  const T *zb = z.ptr + (long long)img * z.img_stride + (long long)pl * hw * 8;
  const T *gb = da.ptr + (long long)img * da.img_stride + (long long)pl * hw * 8;
  T *ob = dz.ptr + (long long)img * dz.img_stride + (long long)pl * hw * 8;
  const T *mb = mask ? mask + ((long long)img * z.planes + pl) * hw * 8 : nullptr;
  const float inv_m = 1.f / (float)count;
  __shared__ float s_par[6][8];
  if (threadIdx.x < 8) {
    const int ch = pl * 8 + threadIdx.x;
    s_par[0][threadIdx.x] = mean[ch]; s_par[1][threadIdx.x] = invstd[ch]; s_par[2][threadIdx.x] = gamma[ch];
    s_par[3][threadIdx.x] = beta[ch];
    s_par[4][threadIdx.x] = (float)sums[ch] * inv_m; s_par[5][threadIdx.x] = (float)sums[c + ch] * inv_m;
  }
  __syncthreads();
  float mu[8], is[8], ga[8], be[8], sdy[8], sdyz[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    mu[k] = s_par[0][k]; is[k] = s_par[1][k]; ga[k] = s_par[2][k]; be[k] = s_par[3][k]; sdy[k] = s_par[4][k]; sdyz[k] = s_par[5][k];
  }
  constexpr int kU = 4;      // see bn_bwd_reduce_kernel: loads of 4 vectors in flight per thread
  const int stride = gridDim.x * blockDim.x;
  for (int v0 = blockIdx.x * blockDim.x + threadIdx.x; v0 < hw; v0 += stride * kU) {
    Raw8<T> rz[kU], rg[kU], rm[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int v = v0 + u * stride;
      if (v < hw) {
        rz[u] = load_raw8(zb + (long long)v * 8);
        rg[u] = load_raw8(gb + (long long)v * 8);
        if (mb) rm[u] = load_raw8(mb + (long long)v * 8);
      }
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int v = v0 + u * stride;
      if (v >= hw) break;
      const Vec8f zz = cvt8(rz[u]), g = cvt8(rg[u]);
      Vec8f mk = zero8();
      if (mb) mk = cvt8(rm[u]);
      Vec8f o;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float gk = mb ? g.v[k] * mk.v[k] : g.v[k];
        const float zh = (zz.v[k] - mu[k]) * is[k];
        const float dy = (fmaf(ga[k], zh, be[k]) > 0.f) ? gk : 0.f;
        o.v[k] = ga[k] * is[k] * (dy - sdy[k] - zh * sdyz[k]);
      }
      store8(ob + (long long)v * 8, o);
    }
  }
}

template <typename T>
int launch_bn_bwd_apply(View<const T> da, View<const T> z, const float *mean, const float *invstd,
                        const float *gamma, const float *beta, const T *mask, const double *sums,
                        long long count, View<T> dz, float *d_gamma, float *d_beta, cudaStream_t st) {
  const int hw = z.h * z.w;
  dim3 grid(std::max(1, std::min((hw + 4095) / 4096, 64)), z.planes, z.n);           // >= 16 vectors per thread
  bn_bwd_apply_kernel<T><<<grid, 256, 0, st>>>(da, z, mean, invstd, gamma, beta, mask, sums, count, dz, d_gamma,
                                               d_beta, z.planes * 8);
  OCTSEG_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------
// max-pool backward + skip add: first max in window scan order gets the pooled gradient
// ---------------------------------------------------------------------------------
// BnRed (optional): the BatchNorm-backward reductions of THIS block (sum dy, sum dy * zhat over the batch, dy = da where the
// ReLU was active) are accumulated from the fp32 totals while they are in registers, so bn_bwd_reduce_kernel -- a
// second read of da and z -- is not launched for the encoder-final blocks.
struct BnRed {
  const float *mean, *invstd, *gamma, *beta;   // [c] of the block whose output gradient this kernel produces
  double *sums;                                // [2c]: sum(dy), sum(dy * zhat); zeroed by the caller
};

template <typename T, bool STATS>
__global__ void __launch_bounds__(256, 2) pool_bwd_add_kernel(View<const T> a, View<const T> d_pooled,
                                                           View<const T> d_skip, View<T> out, int has_skip,
                                                           View<const T> z, BnRed bn, int c) {
  const int Ho = d_pooled.h, Wo = d_pooled.w;
  const int pl = blockIdx.y, b = blockIdx.z;
  // per-channel constants through shared memory (broadcast reads): zhat = z * inv - mean * inv, ReLU active <=> z * sc + sh > 0
  __shared__ float s_inv[8], s_minv[8], s_sc[8], s_sh[8];
  float acc[16];
  if constexpr (STATS) {
    if (threadIdx.x < 8) {
      const int ch = pl * 8 + threadIdx.x;
      const float inv = bn.invstd[ch], sc = bn.gamma[ch] * inv;
      s_inv[threadIdx.x] = inv; s_minv[threadIdx.x] = bn.mean[ch] * inv;
      s_sc[threadIdx.x] = sc; s_sh[threadIdx.x] = bn.beta[ch] - bn.mean[ch] * sc;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) acc[k] = 0.f;
  }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < Ho * Wo; i += gridDim.x * blockDim.x) {
    const int x = i % Wo, y = i / Wo;
    const long long base = (((long long)pl * a.h + 2 * y) * a.w + 2 * x) * 8;
    const long long row = (long long)a.w * 8;
    const T *ap = a.ptr + b * a.img_stride + base;
    const Vec8f a00 = load8(ap), a01 = load8(ap + 8), a10 = load8(ap + row), a11 = load8(ap + row + 8);
    const Vec8f g = load8(d_pooled.ptr + b * d_pooled.img_stride + (((long long)pl * Ho + y) * Wo + x) * 8);
    Vec8f o00 = zero8(), o01 = zero8(), o10 = zero8(), o11 = zero8();
    if (has_skip) {
      const T *sp = d_skip.ptr + b * d_skip.img_stride + base;
      o00 = load8(sp); o01 = load8(sp + 8); o10 = load8(sp + row); o11 = load8(sp + row + 8);
    }
    Raw8<T> rz[4];
    if constexpr (STATS) {
      const T *zp = z.ptr + b * z.img_stride + base;        // z is dense with the same plane grid as `a`
      rz[0] = load_raw8(zp); rz[1] = load_raw8(zp + 8); rz[2] = load_raw8(zp + row); rz[3] = load_raw8(zp + row + 8);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float m = a00.v[k];
      int am = 0;
      if (a01.v[k] > m) { m = a01.v[k]; am = 1; }
      if (a10.v[k] > m) { m = a10.v[k]; am = 2; }
      if (a11.v[k] > m) { m = a11.v[k]; am = 3; }
      if (am == 0) o00.v[k] += g.v[k];
      else if (am == 1) o01.v[k] += g.v[k];
      else if (am == 2) o10.v[k] += g.v[k];
      else o11.v[k] += g.v[k];
    }
    T *op = out.ptr + b * out.img_stride + base;
    store8(op, o00); store8(op + 8, o01); store8(op + row, o10); store8(op + row + 8, o11);
    if constexpr (STATS) {
      const Vec8f *oo[4] = {&o00, &o01, &o10, &o11};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const Vec8f zz = cvt8(rz[j]);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float zh = fmaf(zz.v[k], s_inv[k], -s_minv[k]);
          const float dy = (fmaf(zz.v[k], s_sc[k], s_sh[k]) > 0.f) ? oo[j]->v[k] : 0.f;
          acc[k] += dy;
          acc[8 + k] = fmaf(dy, zh, acc[8 + k]);
        }
      }
    }
  }
  if constexpr (STATS) block_reduce16_to_double(acc, bn.sums, bn.sums + c, pl * 8);
}

template <typename T>
int launch_pool_bwd_add(View<const T> a, View<const T> d_pooled, View<const T> d_skip, View<T> da_total,
                        cudaStream_t st) {
  const int hw = d_pooled.h * d_pooled.w;
  dim3 grid(std::max(1, std::min((hw + 255) / 256, 64)), d_pooled.planes, d_pooled.n);
  pool_bwd_add_kernel<T, false><<<grid, 256, 0, st>>>(a, d_pooled, d_skip, da_total, d_skip.ptr != nullptr, View<const T>{},
                                                      BnRed{}, 0);
  OCTSEG_CUDA(cudaGetLastError());
  return 0;
}

// + the BatchNorm-backward reductions of the block (z dense, no dropout on it)
template <typename T>
int launch_pool_bwd_add_bnred(View<const T> a, View<const T> d_pooled, View<const T> d_skip, View<T> da_total, View<const T> z,
                              const float *mean, const float *invstd, const float *gamma, const float *beta, double *sums,
                              cudaStream_t st) {
  const int hw = d_pooled.h * d_pooled.w;
  // >= 8 windows per thread: every block ends with 16 double atomics on the same 16 words per plane
  dim3 grid(std::max(1, std::min((hw + 2047) / 2048, 64)), d_pooled.planes, d_pooled.n);
  BnRed bn{mean, invstd, gamma, beta, sums};
  pool_bwd_add_kernel<T, true><<<grid, 256, 0, st>>>(a, d_pooled, d_skip, da_total, d_skip.ptr != nullptr, z, bn, z.planes * 8);
  OCTSEG_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------
// nearest x2 up-sampling, materialised only for the weight gradient of the >= 128-channel up-conv (tcgen05 kernel); the
// narrower up-convs read the low-res tensor, and the adjoint (2x2 sum-pool) lives in the data-gradient conv
// ---------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) upsample2x_kernel(View<const T> in, View<T> out) {
  const int pl = blockIdx.y, b = blockIdx.z;
  const int h = in.h, w = in.w;
  const T *ib = in.ptr + b * in.img_stride + (long long)pl * h * w * 8;
  T *ob = out.ptr + b * out.img_stride + (long long)pl * (4LL * h * w) * 8;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < h * w; i += gridDim.x * blockDim.x) {
    const int x = i % w, y = i / w;
    const uint4 v = *reinterpret_cast<const uint4 *>(ib + (long long)i * 8);
    T *o = ob + ((long long)(2 * y) * (2 * w) + 2 * x) * 8;
    if constexpr (sizeof(T) == 2) {
      *reinterpret_cast<uint4 *>(o) = v; *reinterpret_cast<uint4 *>(o + 8) = v;
      *reinterpret_cast<uint4 *>(o + (long long)2 * w * 8) = v; *reinterpret_cast<uint4 *>(o + (long long)2 * w * 8 + 8) = v;
    } else {
      const Vec8f f = load8(ib + (long long)i * 8);
      store8(o, f); store8(o + 8, f); store8(o + (long long)2 * w * 8, f); store8(o + (long long)2 * w * 8 + 8, f);
    }
  }
}
template <typename T>
int launch_upsample2x(View<const T> in, View<T> out, cudaStream_t st) {
  dim3 grid(std::max(1, std::min((in.h * in.w + 255) / 256, 64)), in.planes, in.n);
  upsample2x_kernel<T><<<grid, 256, 0, st>>>(in, out);
  OCTSEG_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------
// weight gradient (CUDA cores): block = one (ci-plane, co-plane) pair x one 32x8 pixel tile
//   thread (co8, ci8, q): 9 tap accumulators over a quarter of the tile's pixels
// ---------------------------------------------------------------------------------
constexpr int kWgTW = 32, kWgTH = 8;

template <typename T>
__global__ void __launch_bounds__(256) wgrad_kernel(View<const T> a_in, View<const T> dz, int kh, int kw, int pt,
                                                    int pl_, int ups, int cin, int cout, float *__restrict__ dW,
                                                    float *__restrict__ db, int tiles_x) {
  extern __shared__ float sm[];
  const int aw = kWgTW + kw - 1, ah = kWgTH + kh - 1;
  float *s_a = sm;                       // [ah][aw][8]
  float *s_d = s_a + ah * aw * 8;        // [TH][TW][8]
  float *s_red = s_d + kWgTH * kWgTW * 8;   // [4][64][9]
  const int co_planes = cout / 8;
  const int cgi = blockIdx.y / co_planes, cgo = blockIdx.y % co_planes;
  const int b = blockIdx.z;
  const int tile_x = blockIdx.x % tiles_x, tile_y = blockIdx.x / tiles_x;
  const int H = dz.h, W = dz.w;          // output grid of the conv
  const int x0 = tile_x * kWgTW, y0 = tile_y * kWgTH;
  // stage the (virtual) input tile with halo; ups: virtual coords index the x2-upsampled input
  for (int i = threadIdx.x; i < ah * aw; i += blockDim.x) {
    const int ty = i / aw, tx = i % aw;
    int vy = y0 + ty - pt, vx = x0 + tx - pl_;
    Vec8f v = zero8();
    if (vy >= 0 && vy < H && vx >= 0 && vx < W) {
      if (ups) { vy >>= 1; vx >>= 1; }
      v = load8(a_in.ptr + b * a_in.img_stride + (((long long)cgi * a_in.h + vy) * a_in.w + vx) * 8);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) s_a[i * 8 + k] = v.v[k];
  }
  for (int i = threadIdx.x; i < kWgTH * kWgTW; i += blockDim.x) {
    const int ty = i / kWgTW, tx = i % kWgTW;
    const int y = y0 + ty, x = x0 + tx;
    Vec8f v = zero8();
    if (y < H && x < W) v = load8(dz.ptr + b * dz.img_stride + (((long long)cgo * H + y) * W + x) * 8);
#pragma unroll
    for (int k = 0; k < 8; ++k) s_d[i * 8 + k] = v.v[k];
  }
  __syncthreads();
  const int co8 = threadIdx.x & 7, ci8 = (threadIdx.x >> 3) & 7, q = threadIdx.x >> 6;
  float acc[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) acc[t] = 0.f;
  float dsum = 0.f;
  for (int p = q * (kWgTH * kWgTW / 4); p < (q + 1) * (kWgTH * kWgTW / 4); ++p) {
    const int ty = p / kWgTW, tx = p % kWgTW;
    const float d = s_d[p * 8 + co8];
    dsum += d;
#pragma unroll
    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
      for (int dx = 0; dx < 3; ++dx)
        if (dy < kh && dx < kw) acc[dy * 3 + dx] = fmaf(s_a[((ty + dy) * aw + tx + dx) * 8 + ci8], d, acc[dy * 3 + dx]);
  }
#pragma unroll
  for (int t = 0; t < 9; ++t) s_red[(q * 64 + (threadIdx.x & 63)) * 9 + t] = acc[t];
  __syncthreads();
  if (q == 0) {
    for (int dy = 0; dy < kh; ++dy)
      for (int dx = 0; dx < kw; ++dx) {
        const int t = dy * 3 + dx;
        const float s = s_red[(0 * 64 + threadIdx.x) * 9 + t] + s_red[(1 * 64 + threadIdx.x) * 9 + t] +
                        s_red[(2 * 64 + threadIdx.x) * 9 + t] + s_red[(3 * 64 + threadIdx.x) * 9 + t];
        atomicAdd(&dW[(((long long)dy * kw + dx) * cin + cgi * 8 + ci8) * cout + cgo * 8 + co8], s);
      }
  }
  if (db && cgi == 0 && ci8 == 0) {
    // bias gradient: sum of dz over pixels (4 quarter-partials per co8)
    atomicAdd(&db[cgo * 8 + co8], dsum);
  }
}

template <typename T>
int launch_wgrad(View<const T> a_in, View<const T> dz, int kh, int kw, int pad_top, int pad_left,
                 int ups, int cin, int cout, float *dW, float *db, cudaStream_t st) {
  if (kh > 3 || kw > 3) { set_error("wgrad: kernel larger than 3x3 not supported"); return 1; }
  const int tiles_x = (dz.w + kWgTW - 1) / kWgTW, tiles_y = (dz.h + kWgTH - 1) / kWgTH;
  dim3 grid(tiles_x * tiles_y, (cin / 8) * (cout / 8), dz.n);
  const int aw = kWgTW + kw - 1, ah = kWgTH + kh - 1;
  size_t smem = (size_t)(ah * aw * 8 + kWgTH * kWgTW * 8 + 4 * 64 * 9) * sizeof(float);
  wgrad_kernel<T><<<grid, 256, smem, st>>>(a_in, dz, kh, kw, pad_top, pad_left, ups, cin, cout, dW, db, tiles_x);
  OCTSEG_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------
template <typename T, typename IMG>
__global__ void image_to_blocked_kernel(const IMG *__restrict__ img, int n, int h, int w, int cin, int pre,
                                        T *__restrict__ out) {
  const long long total = (long long)n * h * w;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < total;
       p += (long long)gridDim.x * blockDim.x) {
    Vec8f v = zero8();
    for (int c = 0; c < cin && c < 8; ++c) {
      const IMG raw = img[p * cin + c];
      if constexpr (sizeof(IMG) == 1) v.v[c] = __fdiv_rn((float)raw, 255.0f);
      else v.v[c] = pre ? (float)raw : (float)((double)raw / 255.0);
    }
    store8(out + p * 8, v);   // single plane: [n][1][h][w][8]
  }
}

template <typename T>
int launch_image_to_blocked(const void *img, int img_dtype, int n, int h, int w, int cin, T *out,
                            cudaStream_t st) {
  if (cin > 8) { set_error("training supports input_channels <= 8"); return 1; }
  const long long total = (long long)n * h * w;
  unsigned grid = (unsigned)std::min<long long>((total + 255) / 256, 148 * 16);
  if (img_dtype == 0)
    image_to_blocked_kernel<T, uint8_t><<<grid, 256, 0, st>>>((const uint8_t *)img, n, h, w, cin, 0, out);
  else
    image_to_blocked_kernel<T, float><<<grid, 256, 0, st>>>((const float *)img, n, h, w, cin, img_dtype == 2, out);
  OCTSEG_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------
__global__ void flip_transpose_kernel(const float *__restrict__ w, int kh, int kw, int cin, int cout,
                                      float *__restrict__ out) {
  const int total = kh * kw * cin * cout;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int co = i % cout;
    const int ci = (i / cout) % cin;
    const int b = (i / (cout * cin)) % kw;
    const int a = i / (cout * cin * kw);
    out[(((kh - 1 - a) * kw + (kw - 1 - b)) * cout + co) * cin + ci] = w[i];
  }
}
int launch_flip_transpose(const float *w, int kh, int kw, int cin, int cout, float *out, cudaStream_t st) {
  flip_transpose_kernel<<<64, 256, 0, st>>>(w, kh, kw, cin, cout, out);
  OCTSEG_CUDA(cudaGetLastError());
  return 0;
}

// Wd[r][c][co][ci] = sum_{py,a: py-a+kh-1 == r} sum_{px,b: px-b+kw-1 == c} w[a][b][ci][co]
__global__ void upconv_dgrad_weights_kernel(const float *__restrict__ w, int kh, int kw, int cin, int cout,
                                            float *__restrict__ out) {
  const int R = kh + 1, C = kw + 1;
  const int total = R * C * cout * cin;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int ci = i % cin;
    const int co = (i / cin) % cout;
    const int c = (i / (cin * cout)) % C;
    const int r = i / (cin * cout * C);
    float s = 0.f;
    for (int py = 0; py < 2; ++py) {
      const int a = py + kh - 1 - r;
      if (a < 0 || a >= kh) continue;
      for (int px = 0; px < 2; ++px) {
        const int b = px + kw - 1 - c;
        if (b < 0 || b >= kw) continue;
        s += w[((a * kw + b) * cin + ci) * cout + co];
      }
    }
    out[i] = s;
  }
}
int launch_upconv_dgrad_weights(const float *w, int kh, int kw, int cin, int cout, float *out,
                                cudaStream_t st) {
  upconv_dgrad_weights_kernel<<<64, 256, 0, st>>>(w, kh, kw, cin, cout, out);
  OCTSEG_CUDA(cudaGetLastError());
  return 0;
}

__global__ void stem_wgrad_extract_kernel(const float *__restrict__ tmp, int taps, int cin, int cout,
                                          float *__restrict__ dW) {
  const int total = taps * cin * cout;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int co = i % cout, ci = (i / cout) % cin, t = i / (cout * cin);
    dW[i] = tmp[(t * 8 + ci) * cout + co];
  }
}
int launch_stem_wgrad_extract(const float *tmp, int taps, int cin, int cout, float *dW, cudaStream_t st) {
  stem_wgrad_extract_kernel<<<8, 256, 0, st>>>(tmp, taps, cin, cout, dW);
  OCTSEG_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------
__global__ void step_advance_kernel(StepState *s, float lr, float b1, float b2) {
  s->step += 1;
  const double t = (double)s->step;
  s->lr_t = (float)((double)lr * sqrt(1.0 - pow((double)b2, t)) / (1.0 - pow((double)b1, t)));   // Keras optimizer_v2
}
int launch_step_advance(StepState *s, float lr, float b1, float b2, cudaStream_t st) {
  step_advance_kernel<<<1, 1, 0, st>>>(s, lr, b1, b2);
  OCTSEG_CUDA(cudaGetLastError());
  return 0;
}

__global__ void __launch_bounds__(256) adam_kernel(float4 *__restrict__ p, const float4 *__restrict__ g,
                                                   float4 *__restrict__ m, float4 *__restrict__ v, long long n4,
                                                   const StepState *__restrict__ state, float b1, float b2, float eps) {
  const float lr_t = state->lr_t;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x) {
    float4 pp = p[i], gg = g[i], mm = m[i], vv = v[i];
#define ADAM1(f)                                            \
  mm.f = b1 * mm.f + (1.f - b1) * gg.f;                     \
  vv.f = b2 * vv.f + (1.f - b2) * gg.f * gg.f;              \
  pp.f -= lr_t * mm.f / (sqrtf(vv.f) + eps);
    ADAM1(x) ADAM1(y) ADAM1(z) ADAM1(w)
#undef ADAM1
    p[i] = pp; m[i] = mm; v[i] = vv;
  }
}
int launch_adam(float *p, const float *g, float *m, float *v, long long n, const StepState *state, float b1, float b2,
                float eps, cudaStream_t st) {
  const long long n4 = n / 4;   // flat buffers are padded to 16 floats per tensor
  unsigned grid = (unsigned)std::min<long long>((n4 + 255) / 256, 148 * 8);
  adam_kernel<<<grid, 256, 0, st>>>((float4 *)p, (const float4 *)g, (float4 *)m, (float4 *)v, n4, state, b1, b2, eps);
  OCTSEG_CUDA(cudaGetLastError());
  return 0;
}

#define INST(T)                                                                                             \
  template int launch_bn_stats<T>(View<const T>, double *, cudaStream_t);                                   \
  template int launch_bn_finalize_apply<T>(View<const T>, const double *, long long, float, float, const float *, \
                                           const float *, float *, float *, float *, float *, float *, float *, \
                                           const T *, View<T>, View<T>, cudaStream_t);                      \
  template int launch_bn_apply_relu<T>(View<const T>, const float *, const float *, const T *, View<T>,     \
                                       cudaStream_t);                                                       \
  template int launch_dropout_mask<T>(const uint8_t *, unsigned long long, const StepState *, float, int, int, \
                                      int, int, T *, cudaStream_t);                                         \
  template int launch_head_loss<T>(View<const T>, const float *, const float *, int, int, const uint8_t *,  \
                                   const float *, float, View<T>, float *, float *, double *, float *, cudaStream_t); \
  template int launch_bn_bwd_reduce<T>(View<const T>, View<const T>, const float *, const float *,          \
                                       const float *, const float *, const T *, double *, cudaStream_t);    \
  template int launch_bn_bwd_apply<T>(View<const T>, View<const T>, const float *, const float *,           \
                                      const float *, const float *, const T *, const double *, long long,   \
                                      View<T>, float *, float *, cudaStream_t);                             \
  template int launch_pool_bwd_add<T>(View<const T>, View<const T>, View<const T>, View<T>, cudaStream_t);  \
  template int launch_pool_bwd_add_bnred<T>(View<const T>, View<const T>, View<const T>, View<T>, View<const T>, const float *,  \
                                            const float *, const float *, const float *, double *, cudaStream_t);  \
  template int launch_wgrad<T>(View<const T>, View<const T>, int, int, int, int, int, int, int, float *,    \
                               float *, cudaStream_t);                                                      \
  template int launch_image_to_blocked<T>(const void *, int, int, int, int, int, T *, cudaStream_t);         \
  template int launch_upsample2x<T>(View<const T>, View<T>, cudaStream_t);
INST(float)
INST(__nv_bfloat16)

}  // namespace octseg
