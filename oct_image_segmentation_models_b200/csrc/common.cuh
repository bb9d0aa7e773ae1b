// Shared declarations for the liboctseg kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <atomic>
#include <string>

namespace octseg {

// ---------------------------------------------------------------------------------
// Device activation layout: channel-blocked [N][C/8][H][W][8]  ("plane" = 8 channels).
// A view may address a slice of planes inside a wider buffer (skip-concat written in
// place: reference models/unet.py:52 concatenates [up, skip]; here both producers
// write straight into their plane ranges of one buffer).
// ---------------------------------------------------------------------------------
template <typename T>
struct View {
  T *ptr;            // first element of plane 0 of image 0 of this view
  int n, planes, h, w;
  long long img_stride;   // elements between consecutive images
};

template <typename T>
__host__ __device__ inline View<T> make_view(T *base, int n, int planes_total, int plane0, int planes,
                                             int h, int w) {
  View<T> v;
  v.ptr = base + (long long)plane0 * h * w * 8;
  v.n = n; v.planes = planes; v.h = h; v.w = w;
  v.img_stride = (long long)planes_total * h * w * 8;
  return v;
}

struct Vec8f { float v[8]; };

__device__ __forceinline__ Vec8f load8(const float *p) {
  Vec8f r;
  float4 a = *reinterpret_cast<const float4 *>(p);
  float4 b = *reinterpret_cast<const float4 *>(p + 4);
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
__device__ __forceinline__ Vec8f load8(const __nv_bfloat16 *p) {
  Vec8f r;
  uint4 u = *reinterpret_cast<const uint4 *>(p);
  const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __bfloat1622float2(h[i]);
    r.v[2 * i] = f.x; r.v[2 * i + 1] = f.y;
  }
  return r;
}
__device__ __forceinline__ Vec8f load8(const __half *p) {
  Vec8f r;
  uint4 u = *reinterpret_cast<const uint4 *>(p);
  const __half2 *h = reinterpret_cast<const __half2 *>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __half22float2(h[i]);
    r.v[2 * i] = f.x; r.v[2 * i + 1] = f.y;
  }
  return r;
}
__device__ __forceinline__ void store8(__half *p, const Vec8f &r) {
  uint4 u;
  __half2 *h = reinterpret_cast<__half2 *>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(r.v[2 * i], r.v[2 * i + 1]);
  *reinterpret_cast<uint4 *>(p) = u;
}
__device__ __forceinline__ void store8(float *p, const Vec8f &r) {
  *reinterpret_cast<float4 *>(p) = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
  *reinterpret_cast<float4 *>(p + 4) = make_float4(r.v[4], r.v[5], r.v[6], r.v[7]);
}
__device__ __forceinline__ void store8(__nv_bfloat16 *p, const Vec8f &r) {
  uint4 u;
  __nv_bfloat162 *h = reinterpret_cast<__nv_bfloat162 *>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(r.v[2 * i], r.v[2 * i + 1]);
  *reinterpret_cast<uint4 *>(p) = u;
}
// Storage tag of the fp32-accurate tensor-core mode: every logical 8-channel plane is a PAIR of fp16 planes,
// hi = rn(a) and lo' = rn((a - hi) * 2^11), so a == hi + lo' * 2^-11 to 2^-22 relative (|a| <= 65504).
struct SplitHalf { uint16_t bits; };
template <typename T> struct PlaneMul { static constexpr int v = 1; };
template <> struct PlaneMul<SplitHalf> { static constexpr int v = 2; };

// store the 8 channels of one pixel of a LOGICAL plane; p = that pixel in the plane's first physical plane
template <typename T>
__device__ __forceinline__ void store_plane8(T *p, long long plane_elems, const Vec8f &r) { (void)plane_elems; store8(p, r); }
template <>
__device__ __forceinline__ void store_plane8<SplitHalf>(SplitHalf *p, long long plane_elems, const Vec8f &r) {
  uint4 hi, lo;
  uint32_t *h = reinterpret_cast<uint32_t *>(&hi), *l = reinterpret_cast<uint32_t *>(&lo);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __half2 h2 = __floats2half2_rn(r.v[2 * i], r.v[2 * i + 1]);
    const float2 hf = __half22float2(h2);
    const __half2 l2 = __floats2half2_rn((r.v[2 * i] - hf.x) * 2048.f, (r.v[2 * i + 1] - hf.y) * 2048.f);
    h[i] = *reinterpret_cast<const uint32_t *>(&h2);
    l[i] = *reinterpret_cast<const uint32_t *>(&l2);
  }
  *reinterpret_cast<uint4 *>(p) = hi;
  *reinterpret_cast<uint4 *>(p + plane_elems) = lo;
}

// Raw (unconverted) 8-channel vector: lets a streaming kernel keep several loads in flight without paying 8 fp32
// registers per vector before the data is needed.
template <typename T> struct Raw8 { uint4 u; };
template <> struct Raw8<float> { float4 a, b; };
template <typename T>
__device__ __forceinline__ Raw8<T> load_raw8(const T *p) { Raw8<T> r; r.u = *reinterpret_cast<const uint4 *>(p); return r; }
template <>
__device__ __forceinline__ Raw8<float> load_raw8<float>(const float *p) {
  Raw8<float> r; r.a = *reinterpret_cast<const float4 *>(p); r.b = *reinterpret_cast<const float4 *>(p + 4); return r;
}
__device__ __forceinline__ Vec8f cvt8(const Raw8<__nv_bfloat16> &r) {
  Vec8f o;
  const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&r.u);
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float2 f = __bfloat1622float2(h[i]); o.v[2 * i] = f.x; o.v[2 * i + 1] = f.y; }
  return o;
}
__device__ __forceinline__ Vec8f cvt8(const Raw8<__half> &r) {
  Vec8f o;
  const __half2 *h = reinterpret_cast<const __half2 *>(&r.u);
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float2 f = __half22float2(h[i]); o.v[2 * i] = f.x; o.v[2 * i + 1] = f.y; }
  return o;
}
__device__ __forceinline__ Vec8f cvt8(const Raw8<float> &r) {
  Vec8f o;
  o.v[0] = r.a.x; o.v[1] = r.a.y; o.v[2] = r.a.z; o.v[3] = r.a.w; o.v[4] = r.b.x; o.v[5] = r.b.y; o.v[6] = r.b.z; o.v[7] = r.b.w;
  return o;
}

__device__ __forceinline__ Vec8f zero8() {
  Vec8f r;
#pragma unroll
  for (int i = 0; i < 8; ++i) r.v[i] = 0.f;
  return r;
}

// One-time per-DEVICE initialisation (e.g. cudaFuncSetAttribute, which is a per-device setting): a process may
// drive several GPUs with one host thread each, so a plain `static bool` is both wrong and a data race.
struct PerDeviceOnce {
  std::atomic<bool> done[64];
  // returns the current device index when its initialisation has not run yet, else -1
  int pending() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return -1;
    return done[dev].load(std::memory_order_acquire) ? -1 : dev;
  }
  void mark(int dev) { done[dev].store(true, std::memory_order_release); }
};

// error plumbing -------------------------------------------------------------------
void set_error(const std::string &msg);
#define OCTSEG_CUDA(expr)                                                              \
  do {                                                                                 \
    cudaError_t _e = (expr);                                                           \
    if (_e != cudaSuccess) {                                                           \
      octseg::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e) + " at " +  \
                        __FILE__ + ":" + std::to_string(__LINE__));                    \
      return 1;                                                                        \
    }                                                                                  \
  } while (0)

}  // namespace octseg
