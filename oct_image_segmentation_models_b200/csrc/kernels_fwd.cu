// CUDA-core forward kernels: first conv (raw image in), generic direct conv (fp32-exact
// path and fallback), max-pool, 1x1+softmax head, BN fold.  All tensors use the blocked
// [N][C/8][H][W][8] layout so every pixel access is one 16 B (bf16) / 32 B (fp32) vector
// and consecutive threads touch consecutive vectors (fully coalesced).
#include "kernels.cuh"

namespace octseg {

int init_preprocess_lut() { return 0; }   // x/255 is computed in-kernel (see conv_first_kernel)

// ---------------------------------------------------------------------------------
// first conv: thread = pixel; the kh*kw*cin (<= 27) preprocessed inputs are loaded once
// into registers and reused for every output plane.  x/255: for u8 inputs
// float32(u)/255f is bit-identical to float32(float64(u)/255) for all 256 values
// (checked in tests/test_oracle.py), so no table and no fp64 is needed.
// ---------------------------------------------------------------------------------
constexpr int kFirstMaxTaps = 27;

// KH/KW/CIN > 0: compile-time filter geometry (the 3x3x1 stem of every OCT model); 0 = runtime
template <typename T, typename IMG, int KH, int KW, int CIN>
__global__ void __launch_bounds__(256) conv_first_kernel(
    const IMG *__restrict__ img, int n, int h, int w, int cin, const float *__restrict__ wgt, int kh,
    int kw, int cout, const float *__restrict__ scale, const float *__restrict__ shift, int relu,
    View<T> out, int pre) {
  extern __shared__ float wsm[];  // [kh*kw*cin][cout] + scale[cout] + shift[cout]
  if constexpr (KH > 0) { kh = KH; kw = KW; cin = CIN; }
  const int taps = kh * kw * cin;
  for (int i = threadIdx.x; i < taps * cout; i += blockDim.x) wsm[i] = wgt[i];
  for (int i = threadIdx.x; i < cout; i += blockDim.x) {
    wsm[taps * cout + i] = scale[i];
    wsm[taps * cout + cout + i] = shift[i];
  }
  __syncthreads();
  const long long total = (long long)n * h * w;
  const int pt = (kh - 1) / 2, pl = (kw - 1) / 2;
  for (long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x; pix < total;
       pix += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(pix % w);
    const int y = (int)((pix / w) % h);
    const int b = (int)(pix / ((long long)w * h));
    float in[kFirstMaxTaps];
#pragma unroll
    for (int t = 0; t < kFirstMaxTaps; ++t) {
      float v = 0.f;
      if (t < taps) {
        const int ci = t % cin, dx = (t / cin) % kw, dy = t / (cin * kw);
        const int iy = y + dy - pt, ix = x + dx - pl;
        if (iy >= 0 && iy < h && ix >= 0 && ix < w) {
          const IMG raw = img[(((long long)b * h + iy) * w + ix) * cin + ci];
          if constexpr (sizeof(IMG) == 1) v = __fdiv_rn((float)raw, 255.0f);
          else v = pre ? (float)raw : (float)((double)raw / 255.0);
        }
      }
      in[t] = v;
    }
    for (int cog = 0; cog < cout / 8; ++cog) {
      Vec8f acc = zero8();
#pragma unroll
      for (int t = 0; t < kFirstMaxTaps; ++t) {
        if (t < taps) {
          const float4 wa = *reinterpret_cast<const float4 *>(wsm + t * cout + cog * 8);
          const float4 wb = *reinterpret_cast<const float4 *>(wsm + t * cout + cog * 8 + 4);
          acc.v[0] = fmaf(in[t], wa.x, acc.v[0]); acc.v[1] = fmaf(in[t], wa.y, acc.v[1]);
          acc.v[2] = fmaf(in[t], wa.z, acc.v[2]); acc.v[3] = fmaf(in[t], wa.w, acc.v[3]);
          acc.v[4] = fmaf(in[t], wb.x, acc.v[4]); acc.v[5] = fmaf(in[t], wb.y, acc.v[5]);
          acc.v[6] = fmaf(in[t], wb.z, acc.v[6]); acc.v[7] = fmaf(in[t], wb.w, acc.v[7]);
        }
      }
#pragma unroll
      for (int co = 0; co < 8; ++co) {
        float v = fmaf(acc.v[co], wsm[taps * cout + cog * 8 + co], wsm[taps * cout + cout + cog * 8 + co]);
        acc.v[co] = relu ? fmaxf(v, 0.f) : v;
      }
      T *dst = out.ptr + b * out.img_stride + (((long long)cog * PlaneMul<T>::v * h + y) * w + x) * 8;
      store_plane8(dst, (long long)h * w * 8, acc);
    }
  }
}

// ---------------------------------------------------------------------------------
// 3x3 / 1-channel stem, the first conv of every OCT U-Net: thread = 4 consecutive pixels
// of a row.  3 rows x 6 raw pixels are fetched once (one aligned 4-byte word + two edge
// bytes per row), preprocessed once, and reused for 4 px x all output planes; stores are
// 64 contiguous bytes per thread per plane.
// ---------------------------------------------------------------------------------
template <typename T, typename IMG>
__global__ void __launch_bounds__(256) conv_stem3x3_kernel(
    const IMG *__restrict__ img, int n, int h, int w, const float *__restrict__ wgt, int cout,
    const float *__restrict__ scale, const float *__restrict__ shift, int relu, View<T> out, int pre) {
  extern __shared__ float wsm[];  // [9][cout] + scale[cout] + shift[cout]
  for (int i = threadIdx.x; i < 9 * cout; i += blockDim.x) wsm[i] = wgt[i];
  for (int i = threadIdx.x; i < cout; i += blockDim.x) {
    wsm[9 * cout + i] = scale[i];
    wsm[10 * cout + i] = shift[i];
  }
  __syncthreads();
  const int w4 = w >> 2;                       // launcher guarantees w % 4 == 0
  const long long total = (long long)n * h * w4;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < total;
       q += (long long)gridDim.x * blockDim.x) {
    const int x0 = (int)(q % w4) * 4;
    const int y = (int)((q / w4) % h);
    const int b = (int)(q / ((long long)w4 * h));
    float in[3][6];
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      const int iy = y + dy - 1;
      const bool rowok = (iy >= 0 && iy < h);
      const IMG *row = img + ((long long)b * h + (rowok ? iy : 0)) * w;
      IMG raw[6];
      if constexpr (sizeof(IMG) == 1) {
        const uint32_t word = rowok ? *reinterpret_cast<const uint32_t *>(row + x0) : 0u;
        raw[1] = (IMG)(word & 0xFF); raw[2] = (IMG)((word >> 8) & 0xFF);
        raw[3] = (IMG)((word >> 16) & 0xFF); raw[4] = (IMG)(word >> 24);
      } else {
        const float4 word = rowok ? *reinterpret_cast<const float4 *>(row + x0) : make_float4(0, 0, 0, 0);
        raw[1] = word.x; raw[2] = word.y; raw[3] = word.z; raw[4] = word.w;
      }
      raw[0] = (rowok && x0 > 0) ? row[x0 - 1] : (IMG)0;
      raw[5] = (rowok && x0 + 4 < w) ? row[x0 + 4] : (IMG)0;
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        if constexpr (sizeof(IMG) == 1) in[dy][i] = __fdiv_rn((float)raw[i], 255.0f);
        else in[dy][i] = pre ? (float)raw[i] : (float)((double)raw[i] / 255.0);
      }
    }
    for (int cog = 0; cog < cout / 8; ++cog) {
      Vec8f acc[4];
#pragma unroll
      for (int p = 0; p < 4; ++p) acc[p] = zero8();
#pragma unroll
      for (int dy = 0; dy < 3; ++dy)
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          const float *wr = wsm + (dy * 3 + dx) * cout + cog * 8;
          const float4 wa = *reinterpret_cast<const float4 *>(wr);
          const float4 wb = *reinterpret_cast<const float4 *>(wr + 4);
#pragma unroll
          for (int p = 0; p < 4; ++p) {
            const float v = in[dy][p + dx];
            acc[p].v[0] = fmaf(v, wa.x, acc[p].v[0]); acc[p].v[1] = fmaf(v, wa.y, acc[p].v[1]);
            acc[p].v[2] = fmaf(v, wa.z, acc[p].v[2]); acc[p].v[3] = fmaf(v, wa.w, acc[p].v[3]);
            acc[p].v[4] = fmaf(v, wb.x, acc[p].v[4]); acc[p].v[5] = fmaf(v, wb.y, acc[p].v[5]);
            acc[p].v[6] = fmaf(v, wb.z, acc[p].v[6]); acc[p].v[7] = fmaf(v, wb.w, acc[p].v[7]);
          }
        }
      T *dst = out.ptr + b * out.img_stride + (((long long)cog * PlaneMul<T>::v * h + y) * w + x0) * 8;
#pragma unroll
      for (int p = 0; p < 4; ++p) {
#pragma unroll
        for (int co = 0; co < 8; ++co) {
          float v = fmaf(acc[p].v[co], wsm[9 * cout + cog * 8 + co], wsm[10 * cout + cog * 8 + co]);
          acc[p].v[co] = relu ? fmaxf(v, 0.f) : v;
        }
        store_plane8(dst + p * 8, (long long)h * w * 8, acc[p]);
      }
    }
  }
}

// ---------------------------------------------------------------------------------
// Tiled 3x3 / 1-channel stem (the default path): block = 128 x 16 output pixels.  The raw
// halo tile is fetched and preprocessed ONCE per pixel into shared memory; every warp then
// owns two rows and every lane the pixels lane, lane+32, lane+64, lane+96 of a row, so each
// 16-byte (bf16) / 32-byte (fp32) store instruction of a warp covers one contiguous run and
// every shared-memory read is conflict-free.  The 72 weights of an output plane sit in
// registers as float2 pairs and the MACs are packed FFMA2 (sm_100 f32x2 pipe).
// ---------------------------------------------------------------------------------
constexpr int kStemTW = 128, kStemTH = 16, kStemPitch = kStemTW + 8, kStemX = 4;   // body starts at column 4 (16 B aligned)

template <typename IMG>
__device__ __forceinline__ float stem_preprocess(IMG raw, int pre) {
  if constexpr (sizeof(IMG) == 1) return __fdiv_rn((float)raw, 255.0f);   // == float32(float64(u) / 255)
  else return pre ? (float)raw : (float)((double)raw / 255.0);
}

template <typename T, typename IMG>
__global__ void __launch_bounds__(256, 2) conv_stem3x3_tile_kernel(
    const IMG *__restrict__ img, int h, int w, const float *__restrict__ wgt, int cout,
    const float *__restrict__ scale, const float *__restrict__ shift, int relu, View<T> out, int pre) {
  __shared__ __align__(16) float tile[kStemTH + 2][kStemPitch];
  __shared__ float2 ss[2][64];                 // scale / shift pairs, cout <= 128
  for (int i = threadIdx.x; i < cout / 2; i += 256) {
    ss[0][i] = *reinterpret_cast<const float2 *>(scale + 2 * i);
    ss[1][i] = *reinterpret_cast<const float2 *>(shift + 2 * i);
  }
  const int x0 = blockIdx.x * kStemTW, y0 = blockIdx.y * kStemTH, b = blockIdx.z;
  const IMG *src = img + (long long)b * h * w;
  // halo tile: 16-byte vector loads, all issued before the first use (one exposed latency per block);
  // launcher guarantees w % 16 == 0 and a 16-byte aligned image, so a vector is inside the row or outside it
  constexpr int kPerVec = 16 / (int)sizeof(IMG);                 // pixels per 16-byte load
  constexpr int kVecPerRow = kStemTW / kPerVec;
  constexpr int kVecs = (kStemTH + 2) * kVecPerRow;
  constexpr int kIter = (kVecs + 255) / 256;
  uint4 raw[kIter];
#pragma unroll
  for (int it = 0; it < kIter; ++it) {
    const int i = threadIdx.x + it * 256;
    const int r = i / kVecPerRow, v = i - r * kVecPerRow;
    const int gy = y0 - 1 + r, gx = x0 + v * kPerVec;
    raw[it] = make_uint4(0, 0, 0, 0);
    if (i < kVecs && gy >= 0 && gy < h && gx < w) raw[it] = *reinterpret_cast<const uint4 *>(src + (long long)gy * w + gx);
  }
  if (threadIdx.x < 2 * (kStemTH + 2)) {        // left / right halo columns
    const int r = threadIdx.x >> 1, side = threadIdx.x & 1;
    const int gy = y0 - 1 + r, gx = side ? x0 + kStemTW : x0 - 1;
    float v = 0.f;
    if (gy >= 0 && gy < h && gx >= 0 && gx < w) v = stem_preprocess<IMG>(src[(long long)gy * w + gx], pre);
    tile[r][side ? kStemX + kStemTW : kStemX - 1] = v;
  }
#pragma unroll
  for (int it = 0; it < kIter; ++it) {
    const int i = threadIdx.x + it * 256;
    const int r = i / kVecPerRow, v = i - r * kVecPerRow;
    if (i < kVecs) {
      float *dst = &tile[r][kStemX + v * kPerVec];
      if constexpr (sizeof(IMG) == 1) {
        const uint32_t wds[4] = {raw[it].x, raw[it].y, raw[it].z, raw[it].w};
#pragma unroll
        for (int q = 0; q < 4; ++q)
          *reinterpret_cast<float4 *>(dst + 4 * q) =
              make_float4(stem_preprocess<uint8_t>(wds[q] & 0xFF, 0), stem_preprocess<uint8_t>((wds[q] >> 8) & 0xFF, 0),
                          stem_preprocess<uint8_t>((wds[q] >> 16) & 0xFF, 0), stem_preprocess<uint8_t>(wds[q] >> 24, 0));
      } else {
        *reinterpret_cast<float4 *>(dst) =
            make_float4(stem_preprocess<float>(__uint_as_float(raw[it].x), pre), stem_preprocess<float>(__uint_as_float(raw[it].y), pre),
                        stem_preprocess<float>(__uint_as_float(raw[it].z), pre), stem_preprocess<float>(__uint_as_float(raw[it].w), pre));
      }
    }
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, wy = threadIdx.x >> 5;
  for (int cog = 0; cog < cout / 8; ++cog) {
    float2 wr[9][4];
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
      for (int k = 0; k < 4; ++k) wr[t][k] = *reinterpret_cast<const float2 *>(wgt + t * cout + cog * 8 + 2 * k);
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int r = wy * 2 + j, y = y0 + r;
      if (y >= h) break;
      T *row = out.ptr + b * out.img_stride + ((long long)cog * PlaneMul<T>::v * h + y) * w * 8;
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const int xl = lane + 32 * p, x = x0 + xl;
        float2 acc[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[k] = make_float2(0.f, 0.f);
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
          for (int dx = 0; dx < 3; ++dx) {
            const float v = tile[r + dy][xl + dx + kStemX - 1];
            const float2 vv = make_float2(v, v);
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[k] = __ffma2_rn(vv, wr[dy * 3 + dx][k], acc[k]);
          }
        Vec8f o;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 y2 = __ffma2_rn(acc[k], ss[0][cog * 4 + k], ss[1][cog * 4 + k]);
          o.v[2 * k] = relu ? fmaxf(y2.x, 0.f) : y2.x;
          o.v[2 * k + 1] = relu ? fmaxf(y2.y, 0.f) : y2.y;
        }
        if (x < w) store_plane8(row + (long long)x * 8, (long long)h * w * 8, o);
      }
    }
  }
}

template <typename T>
int launch_conv_first(const void *img, int img_dtype, int n, int h, int w, int cin_img,
                      const float *wgt, int kh, int kw, int cout, const float *scale,
                      const float *shift, int relu, View<T> out, cudaStream_t st) {
  const int pre = (img_dtype == 2) ? 1 : 0;   // 2 = float32 already divided by 255 on the host
  if (kh * kw * cin_img > kFirstMaxTaps) { set_error("first conv: kh*kw*input_channels > 27 not supported"); return 1; }
  const long long total = (long long)n * h * w;
  unsigned grid = (unsigned)std::min<long long>((total + 255) / 256, 148 * 32);
  size_t smem = ((size_t)kh * kw * cin_img * cout + 2 * cout) * sizeof(float);
  if (kh == 3 && kw == 3 && cin_img == 1 && n <= 65535 && (cout % 8) == 0 && cout <= 128 && (w % 16) == 0 &&
      ((uintptr_t)img % 16) == 0) {
    dim3 gt((w + kStemTW - 1) / kStemTW, (h + kStemTH - 1) / kStemTH, n);
    if (img_dtype == 0)
      conv_stem3x3_tile_kernel<T, uint8_t><<<gt, 256, 0, st>>>((const uint8_t *)img, h, w, wgt, cout, scale, shift,
                                                               relu, out, 0);
    else
      conv_stem3x3_tile_kernel<T, float><<<gt, 256, 0, st>>>((const float *)img, h, w, wgt, cout, scale, shift, relu,
                                                             out, pre);
    OCTSEG_CUDA(cudaGetLastError());
    return 0;
  }
  if (kh == 3 && kw == 3 && cin_img == 1 && (w % 4) == 0 && ((uintptr_t)img % 16) == 0) {
    const long long quads = total / 4;
    unsigned g4 = (unsigned)std::min<long long>((quads + 255) / 256, 148 * 16);
    size_t sm4 = (size_t)11 * cout * sizeof(float);
    if (img_dtype == 0)
      conv_stem3x3_kernel<T, uint8_t><<<g4, 256, sm4, st>>>((const uint8_t *)img, n, h, w, wgt, cout, scale,
                                                            shift, relu, out, 0);
    else
      conv_stem3x3_kernel<T, float><<<g4, 256, sm4, st>>>((const float *)img, n, h, w, wgt, cout, scale, shift,
                                                          relu, out, pre);
    OCTSEG_CUDA(cudaGetLastError());
    return 0;
  }
  const bool stem331 = (kh == 3 && kw == 3 && cin_img == 1);
  if (img_dtype == 0) {
    if (stem331)
      conv_first_kernel<T, uint8_t, 3, 3, 1><<<grid, 256, smem, st>>>((const uint8_t *)img, n, h, w, cin_img,
                                                                      wgt, kh, kw, cout, scale, shift, relu, out, pre);
    else
      conv_first_kernel<T, uint8_t, 0, 0, 0><<<grid, 256, smem, st>>>((const uint8_t *)img, n, h, w, cin_img,
                                                                      wgt, kh, kw, cout, scale, shift, relu, out, pre);
  } else {
    if (stem331)
      conv_first_kernel<T, float, 3, 3, 1><<<grid, 256, smem, st>>>((const float *)img, n, h, w, cin_img, wgt,
                                                                    kh, kw, cout, scale, shift, relu, out, pre);
    else
      conv_first_kernel<T, float, 0, 0, 0><<<grid, 256, smem, st>>>((const float *)img, n, h, w, cin_img, wgt,
                                                                    kh, kw, cout, scale, shift, relu, out, pre);
  }
  OCTSEG_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------
// TC stem helpers.  The tensor-core stem (net.cu) treats 8 adjacent pixels of a row as the 8 input
// channels of one GEMM row, so its A operand is just the image widened to 16 bits: integers 0..255
// are exact in bf16 and fp16, and the 1/255 of the reference's preprocessing is folded into the
// per-column scale.
// ---------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) u8_to_act_kernel(const uint8_t *__restrict__ img, long long n16, T *__restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (long long)gridDim.x * blockDim.x) {
    const uint4 raw = *reinterpret_cast<const uint4 *>(img + i * 16);
    const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      Vec8f v;
#pragma unroll
      for (int k = 0; k < 8; ++k) v.v[k] = (float)((w[hh * 2 + (k >> 2)] >> ((k & 3) * 8)) & 0xFFu);
      store8(out + i * 16 + hh * 8, v);
    }
  }
}
template <typename T>
int launch_u8_to_act(const uint8_t *img, long long count, T *out, cudaStream_t st) {
  if (count % 16 || ((uintptr_t)img % 16)) { set_error("u8_to_act: count and pointer must be 16-aligned"); return 1; }
  const long long n16 = count / 16;
  unsigned grid = (unsigned)std::min<long long>((n16 + 255) / 256, 148 * 16);
  u8_to_act_kernel<T><<<grid, 256, 0, st>>>(img, n16, out);
  OCTSEG_CUDA(cudaGetLastError());
  return 0;
}
template int launch_u8_to_act<__nv_bfloat16>(const uint8_t *, long long, __nv_bfloat16 *, cudaStream_t);
template int launch_u8_to_act<__half>(const uint8_t *, long long, __half *, cudaStream_t);
template int launch_u8_to_act<float>(const uint8_t *, long long, float *, cudaStream_t);

__global__ void stem_rep_kernel(const float *__restrict__ scale, const float *__restrict__ shift, int cout,
                                float *__restrict__ rep_scale, float *__restrict__ rep_shift) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;       // col = plane*64 + pixel*8 + c8
  if (col >= cout * 8) return;
  const int c = (col >> 6) * 8 + (col & 7);
  rep_scale[col] = scale[c] / 255.0f;
  rep_shift[col] = shift[c];
}
int launch_stem_rep(const float *scale, const float *shift, int cout, float *rep_scale, float *rep_shift, cudaStream_t st) {
  stem_rep_kernel<<<(cout * 8 + 127) / 128, 128, 0, st>>>(scale, shift, cout, rep_scale, rep_shift);
  OCTSEG_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------
// generic direct conv: block = 32x4 threads, tile = 64 px x 4 rows, 2 px / thread,
// one output plane (8 couts) per blockIdx.y; weights of the plane staged in smem per
// 64-channel input chunk.
// ---------------------------------------------------------------------------------
constexpr int kDirectChunk = 64;

template <typename T>
__global__ void __launch_bounds__(128) conv_direct_kernel(
    View<const T> in, const float *__restrict__ wgt, int kh, int kw, int cin, int cout, int ups,
    const float *__restrict__ scale, const float *__restrict__ shift, int relu, View<T> out,
    int tiles_x, int stride, int pt, int pl) {
  extern __shared__ float wsm[];  // [kh*kw][chunk][8]
  const int cog = blockIdx.y;
  const int b = blockIdx.z;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int tile_x = blockIdx.x % tiles_x, tile_y = blockIdx.x / tiles_x;
  const int x0 = tile_x * 64 + tx * 2;
  const int y = tile_y * 4 + ty;
  const int H = out.h, W = out.w;           // output grid
  const int Hin = in.h, Win = in.w;         // input grid (H/2 when ups, H*stride when strided)
  const int Hv = ups ? H : Hin, Wv = ups ? W : Win;   // extent of the (virtual) input the taps index
  const int ntap = kh * kw;
  const bool active = (y < H) && (x0 < W);
  Vec8f acc0 = zero8(), acc1 = zero8();

  for (int c0 = 0; c0 < cin; c0 += kDirectChunk) {
    const int chunk = min(kDirectChunk, cin - c0);
    __syncthreads();
    for (int i = threadIdx.x; i < ntap * chunk * 8; i += blockDim.x) {
      const int co = i & 7;
      const int ci = (i >> 3) % chunk;
      const int t = (i >> 3) / chunk;
      wsm[i] = wgt[((long long)t * cin + c0 + ci) * cout + cog * 8 + co];
    }
    __syncthreads();
    if (!active) continue;
    for (int cgl = 0; cgl < chunk / 8; ++cgl) {
      const T *plane = in.ptr + b * in.img_stride + (long long)(c0 / 8 + cgl) * Hin * Win * 8;
      for (int dy = 0; dy < kh; ++dy) {
        int iy = y * stride + dy - pt;
        if (iy < 0 || iy >= Hv) continue;
        if (ups) iy >>= 1;
        for (int dx = 0; dx < kw; ++dx) {
          int ix0 = x0 * stride + dx - pl, ix1 = ix0 + stride;
          const bool v0ok = (ix0 >= 0 && ix0 < Wv);
          const bool v1ok = (ix1 >= 0 && ix1 < Wv) && (x0 + 1 < W);
          if (ups) { ix0 >>= 1; ix1 >>= 1; }
          Vec8f v0 = v0ok ? load8(plane + ((long long)iy * Win + ix0) * 8) : zero8();
          Vec8f v1 = v1ok ? load8(plane + ((long long)iy * Win + ix1) * 8) : zero8();
          const float *wt = wsm + ((dy * kw + dx) * chunk + cgl * 8) * 8;
#pragma unroll
          for (int ci = 0; ci < 8; ++ci) {
            const float4 wa = *reinterpret_cast<const float4 *>(wt + ci * 8);
            const float4 wb = *reinterpret_cast<const float4 *>(wt + ci * 8 + 4);
            const float a = v0.v[ci], c = v1.v[ci];
            acc0.v[0] = fmaf(a, wa.x, acc0.v[0]); acc0.v[1] = fmaf(a, wa.y, acc0.v[1]);
            acc0.v[2] = fmaf(a, wa.z, acc0.v[2]); acc0.v[3] = fmaf(a, wa.w, acc0.v[3]);
            acc0.v[4] = fmaf(a, wb.x, acc0.v[4]); acc0.v[5] = fmaf(a, wb.y, acc0.v[5]);
            acc0.v[6] = fmaf(a, wb.z, acc0.v[6]); acc0.v[7] = fmaf(a, wb.w, acc0.v[7]);
            acc1.v[0] = fmaf(c, wa.x, acc1.v[0]); acc1.v[1] = fmaf(c, wa.y, acc1.v[1]);
            acc1.v[2] = fmaf(c, wa.z, acc1.v[2]); acc1.v[3] = fmaf(c, wa.w, acc1.v[3]);
            acc1.v[4] = fmaf(c, wb.x, acc1.v[4]); acc1.v[5] = fmaf(c, wb.y, acc1.v[5]);
            acc1.v[6] = fmaf(c, wb.z, acc1.v[6]); acc1.v[7] = fmaf(c, wb.w, acc1.v[7]);
          }
        }
      }
    }
  }
  if (!active) return;
#pragma unroll
  for (int co = 0; co < 8; ++co) {
    const float s = scale[cog * 8 + co], t = shift[cog * 8 + co];
    float a = fmaf(acc0.v[co], s, t), c = fmaf(acc1.v[co], s, t);
    acc0.v[co] = relu ? fmaxf(a, 0.f) : a;
    acc1.v[co] = relu ? fmaxf(c, 0.f) : c;
  }
  T *dst = out.ptr + b * out.img_stride + (((long long)cog * H + y) * W + x0) * 8;
  store8(dst, acc0);
  if (x0 + 1 < W) store8(dst + 8, acc1);
}

template <typename T>
int launch_conv_direct_ex(View<const T> in, const float *wgt, int kh, int kw, int cin, int cout,
                          int ups, int stride, int pad_top, int pad_left, const float *scale,
                          const float *shift, int relu, View<T> out, cudaStream_t st) {
  const int tiles_x = (out.w + 63) / 64, tiles_y = (out.h + 3) / 4;
  dim3 grid(tiles_x * tiles_y, cout / 8, out.n);
  size_t smem = (size_t)kh * kw * std::min(cin, kDirectChunk) * 8 * sizeof(float);
  conv_direct_kernel<T><<<grid, 128, smem, st>>>(in, wgt, kh, kw, cin, cout, ups, scale, shift, relu, out,
                                               tiles_x, stride, pad_top, pad_left);
  OCTSEG_CUDA(cudaGetLastError());
  return 0;
}

template <typename T>
int launch_conv_direct(View<const T> in, const float *wgt, int kh, int kw, int cin, int cout,
                       int ups, const float *scale, const float *shift, int relu, View<T> out,
                       cudaStream_t st) {
  return launch_conv_direct_ex<T>(in, wgt, kh, kw, cin, cout, ups, 1, (kh - 1) / 2, (kw - 1) / 2, scale, shift,
                                  relu, out, st);
}

// ---------------------------------------------------------------------------------
// 2x2 max pool: thread = output (pixel, plane) vector
// ---------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) maxpool2_kernel(View<const T> in, View<T> out) {
  const int Ho = out.h, Wo = out.w;
  const long long total = (long long)out.n * out.planes * Ho * Wo;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % Wo);
    const int y = (int)((i / Wo) % Ho);
    const int pl = (int)((i / ((long long)Wo * Ho)) % out.planes);
    const int b = (int)(i / ((long long)Wo * Ho * out.planes));
    const T *src = in.ptr + b * in.img_stride + (((long long)pl * in.h + 2 * y) * in.w + 2 * x) * 8;
    Vec8f a = load8(src), c = load8(src + 8);
    Vec8f d = load8(src + (long long)in.w * 8), e = load8(src + (long long)in.w * 8 + 8);
#pragma unroll
    for (int k = 0; k < 8; ++k) a.v[k] = fmaxf(fmaxf(a.v[k], c.v[k]), fmaxf(d.v[k], e.v[k]));
    store8(out.ptr + b * out.img_stride + (((long long)pl * Ho + y) * Wo + x) * 8, a);
  }
}

// fp32 mode on tensor cores: activations are fp16 (hi, lo') plane pairs; the pool acts on hi + lo' * 2^-11
__global__ void __launch_bounds__(256) maxpool2_split_kernel(View<const SplitHalf> in, View<SplitHalf> out) {
  const int Ho = out.h, Wo = out.w, lp = out.planes >> 1;          // logical planes
  const long long total = (long long)out.n * lp * Ho * Wo;
  const long long ipe = (long long)in.h * in.w * 8, ope = (long long)Ho * Wo * 8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % Wo);
    const int y = (int)((i / Wo) % Ho);
    const int pl = (int)((i / ((long long)Wo * Ho)) % lp);
    const int b = (int)(i / ((long long)Wo * Ho * lp));
    const __half *src = reinterpret_cast<const __half *>(in.ptr) + b * in.img_stride + (2LL * pl) * ipe + ((long long)(2 * y) * in.w + 2 * x) * 8;
    Vec8f m;
#pragma unroll
    for (int k = 0; k < 8; ++k) m.v[k] = -3.0e38f;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const long long o = (long long)(q >> 1) * in.w * 8 + (q & 1) * 8;
      const Vec8f hi = load8(src + o), lo = load8(src + ipe + o);
#pragma unroll
      for (int k = 0; k < 8; ++k) m.v[k] = fmaxf(m.v[k], fmaf(lo.v[k], 4.8828125e-4f, hi.v[k]));
    }
    SplitHalf *dst = out.ptr + b * out.img_stride + (2LL * pl) * ope + ((long long)y * Wo + x) * 8;
    store_plane8(dst, ope, m);
  }
}
template <>
int launch_maxpool2<SplitHalf>(View<const SplitHalf> in, View<SplitHalf> out, cudaStream_t st) {
  const long long total = (long long)out.n * (out.planes >> 1) * out.h * out.w;
  unsigned grid = (unsigned)std::min<long long>((total + 255) / 256, 148 * 32);
  maxpool2_split_kernel<<<grid, 256, 0, st>>>(in, out);
  OCTSEG_CUDA(cudaGetLastError());
  return 0;
}

template <typename T>
int launch_maxpool2(View<const T> in, View<T> out, cudaStream_t st) {
  const long long total = (long long)out.n * out.planes * out.h * out.w;
  unsigned grid = (unsigned)std::min<long long>((total + 255) / 256, 148 * 32);
  maxpool2_kernel<T><<<grid, 256, 0, st>>>(in, out);
  OCTSEG_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------
// head: 1x1 conv + softmax (+ first-max argmax).  thread = pixel.
// ---------------------------------------------------------------------------------

template <typename T, int K>
__global__ void __launch_bounds__(256) head_kernel(View<const T> in, const float *__restrict__ wgt,
                                                   const float *__restrict__ bias, int cin,
                                                   float *__restrict__ probs,
                                                   uint8_t *__restrict__ labels) {
  extern __shared__ float wsm[];  // [cin][K] + [K]
  for (int i = threadIdx.x; i < cin * K; i += blockDim.x) wsm[i] = wgt[i];
  for (int i = threadIdx.x; i < K; i += blockDim.x) wsm[cin * K + i] = bias[i];
  __syncthreads();
  const int H = in.h, W = in.w;
  const long long total = (long long)in.n * H * W;
  for (long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x; pix < total;
       pix += (long long)gridDim.x * blockDim.x) {
    const long long hw = pix % ((long long)H * W);
    const int b = (int)(pix / ((long long)H * W));
    float z[K];
#pragma unroll
    for (int k = 0; k < K; ++k) z[k] = wsm[cin * K + k];
    for (int pl = 0; pl < cin / 8; ++pl) {
      Vec8f v = load8(in.ptr + b * in.img_stride + ((long long)pl * H * W + hw) * 8);
#pragma unroll
      for (int ci = 0; ci < 8; ++ci) {
        const float *wr = wsm + (pl * 8 + ci) * K;
#pragma unroll
        for (int k = 0; k < K; ++k) z[k] = fmaf(v.v[ci], wr[k], z[k]);
      }
    }
    float m = z[0];
#pragma unroll
    for (int k = 1; k < K; ++k) m = fmaxf(m, z[k]);
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k) { z[k] = expf(z[k] - m); s += z[k]; }
    const float inv = 1.f / s;
    // np.argmax runs on the float32 probabilities (reference common/utils.py:104): take the
    // first max of p itself so ties created by rounding resolve as they would on the host
    float pm = -1.f;
    int pa = 0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      z[k] *= inv;
      if (z[k] > pm) { pm = z[k]; pa = k; }
    }
    if (probs) {
      float *dst = probs + pix * K;
      if constexpr (K == 4) *reinterpret_cast<float4 *>(dst) = make_float4(z[0], z[1], z[2], z[3]);
      else {
#pragma unroll
        for (int k = 0; k < K; ++k) dst[k] = z[k];
      }
    }
    if (labels) labels[pix] = (uint8_t)pa;
  }
}

template <typename T, int K>
static int launch_head_k(View<const T> in, const float *wgt, const float *bias, int cin, float *probs,
                         uint8_t *labels, cudaStream_t st) {
  const long long total = (long long)in.n * in.h * in.w;
  unsigned grid = (unsigned)std::min<long long>((total + 255) / 256, 148 * 32);
  size_t smem = (size_t)(cin * K + K) * sizeof(float);
  head_kernel<T, K><<<grid, 256, smem, st>>>(in, wgt, bias, cin, probs, labels);
  OCTSEG_CUDA(cudaGetLastError());
  return 0;
}

template <typename T>
int launch_head(View<const T> in, const float *wgt, const float *bias, int cin, int K, float *probs,
                uint8_t *labels, cudaStream_t st) {
  switch (K) {
#define HK(k) case k: return launch_head_k<T, k>(in, wgt, bias, cin, probs, labels, st);
    HK(1) HK(2) HK(3) HK(4) HK(5) HK(6) HK(7) HK(8) HK(9) HK(10) HK(11) HK(12) HK(13) HK(14) HK(15) HK(16)
#undef HK
  }
  set_error("num_classes must be 1..16");
  return 1;
}

// ---------------------------------------------------------------------------------
// labels -> boundary probability maps.  Integer restatement of
// perform_argmax(bin=True) + convert_predictions_to_maps_semantic (reference common/utils.py:73-168):
// with c = one-hot plane (region above for the ILM/CSI flags, else region k), along rows:
//   g = +-np.gradient(c)  -> clamp<0 -> *2      == s[i] in {0,1,2}: interior max(+-(c[i+1]-c[i-1]),0),
//                                                  edges 2*max(+-(one-sided diff),0)
//   g -= np.roll(g,-1) (wraps) -> clamp<0 -> *255 -> uint8   == (max(s[i]-s[(i+1)%H],0)*255) & 0xFF
// (510 wraps to 254 exactly as the float->uint8 cast does in the reference.)
// ---------------------------------------------------------------------------------
__device__ __forceinline__ int bm_s(const uint8_t *col, int stride, int i, int H, int cls, int sign) {
  // c[j] = (label[j] == cls)
  auto c = [&](int j) { return (int)(col[(long long)j * stride] == cls); };
  int d;
  if (H == 1) return 0;
  if (i == 0) d = 2 * (c(1) - c(0));
  else if (i == H - 1) d = 2 * (c(H - 1) - c(H - 2));
  else d = c(i + 1) - c(i - 1);
  d *= sign;
  return d > 0 ? d : 0;
}

__global__ void __launch_bounds__(256) boundary_maps_kernel(const uint8_t *__restrict__ labels, int n, int H, int W,
                                                            int K, int bg_ilm, int bg_csi, int transposed,
                                                            uint8_t *__restrict__ maps) {
  const long long total = (long long)n * (K - 1) * H * W;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    // idx enumerates OUTPUT elements so that stores are coalesced
    int i, x;
    long long t = idx;
    if (transposed) { i = (int)(t % H); t /= H; x = (int)(t % W); t /= W; }
    else { x = (int)(t % W); t /= W; i = (int)(t % H); t /= H; }
    const int m = (int)(t % (K - 1)) + 1;
    const int b = (int)(t / (K - 1));
    const bool above = (m == 1 && bg_ilm) || (m == K - 1 && bg_csi);
    const int cls = above ? m - 1 : m, sign = above ? -1 : 1;
    const uint8_t *col = labels + (long long)b * H * W + x;
    const int s0 = bm_s(col, W, i, H, cls, sign);
    const int s1 = bm_s(col, W, (i + 1) % H, H, cls, sign);
    const int v = s0 - s1;
    maps[idx] = (uint8_t)(((v > 0 ? v : 0) * 255) & 0xFF);
  }
}

// Row-major maps, 4 pixels per thread: the four label rows a pixel quad needs (i-1 .. i+2, with the
// reference's one-sided differences at the first/last row and its wrap-around of the last row) are
// fetched once as 32-bit words and reused for all K-1 maps; one 32-bit store per map.
__global__ void __launch_bounds__(256) boundary_maps_quad_kernel(const uint8_t *__restrict__ labels, int H, int W, int K,
                                                                 int bg_ilm, int bg_csi, uint8_t *__restrict__ maps) {
  const int xq = blockIdx.x * 64 + (threadIdx.x & 63);
  const int i = blockIdx.y * 4 + (threadIdx.x >> 6);
  const int b = blockIdx.z;
  if (xq * 4 >= W || i >= H) return;
  const uint8_t *img = labels + (long long)b * H * W + xq * 4;
  auto row = [&](int j) { return *reinterpret_cast<const uint32_t *>(img + (long long)j * W); };
  // s(i) = relu(sign * mult * (c(ja) - c(jb)))
  auto pick = [&](int r, int &ja, int &jb, int &mult) {
    if (r == 0) { ja = 1; jb = 0; mult = 2; }
    else if (r == H - 1) { ja = H - 1; jb = H - 2; mult = 2; }
    else { ja = r + 1; jb = r - 1; mult = 1; }
  };
  int a0, b0, m0, a1, b1, m1;
  pick(i, a0, b0, m0);
  pick((i + 1) % H, a1, b1, m1);
  const uint32_t wa0 = row(a0), wb0 = row(b0), wa1 = row(a1), wb1 = row(b1);
  for (int m = 1; m < K; ++m) {
    const bool above = (m == 1 && bg_ilm) || (m == K - 1 && bg_csi);
    const uint32_t cls = above ? m - 1 : m;
    const int sign = above ? -1 : 1;
    uint32_t outw = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      auto c = [&](uint32_t wv) { return (int)(((wv >> (8 * k)) & 0xFFu) == cls); };
      int s0 = sign * m0 * (c(wa0) - c(wb0)), s1 = sign * m1 * (c(wa1) - c(wb1));
      s0 = s0 > 0 ? s0 : 0; s1 = s1 > 0 ? s1 : 0;
      const int v = s0 - s1;
      outw |= (uint32_t)(((v > 0 ? v : 0) * 255) & 0xFF) << (8 * k);
    }
    *reinterpret_cast<uint32_t *>(maps + (((long long)b * (K - 1) + (m - 1)) * H + i) * W + xq * 4) = outw;
  }
}

int launch_boundary_maps(const uint8_t *labels, int n, int h, int w, int K, int bg_ilm, int bg_csi, int transposed,
                         uint8_t *maps, cudaStream_t st) {
  const long long total = (long long)n * (K - 1) * h * w;
  if (total <= 0) return 0;
  if (!transposed && h >= 2 && (w % 4) == 0 && n <= 65535 && ((uintptr_t)labels % 4) == 0 && ((uintptr_t)maps % 4) == 0) {
    dim3 grid((w / 4 + 63) / 64, (h + 3) / 4, n);
    boundary_maps_quad_kernel<<<grid, 256, 0, st>>>(labels, h, w, K, bg_ilm, bg_csi, maps);
    OCTSEG_CUDA(cudaGetLastError());
    return 0;
  }
  unsigned grid = (unsigned)std::min<long long>((total + 255) / 256, 148 * 32);
  boundary_maps_kernel<<<grid, 256, 0, st>>>(labels, n, h, w, K, bg_ilm, bg_csi, transposed, maps);
  OCTSEG_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------
// validation counts: grid = (spans, images); per-thread register counters, one block reduction, few atomics
// ---------------------------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(256) eval_counts_kernel(const float *__restrict__ probs, const uint8_t *__restrict__ labels,
                                                          int hw, const float *__restrict__ class_w,
                                                          unsigned long long *__restrict__ counts, double *__restrict__ loss) {
  const int b = blockIdx.y;
  const float *pb = probs + (long long)b * hw * K;
  const uint8_t *lb = labels + (long long)b * hw;
  int c_int[K], c_pred[K], c_true[K];
  float l_sum = 0.f;
#pragma unroll
  for (int k = 0; k < K; ++k) { c_int[k] = 0; c_pred[k] = 0; c_true[k] = 0; }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += gridDim.x * blockDim.x) {
    float p[K];
    if constexpr (K == 4) {
      const float4 v = *reinterpret_cast<const float4 *>(pb + (long long)i * 4);
      p[0] = v.x; p[1] = v.y; p[2] = v.z; p[3] = v.w;
    } else {
#pragma unroll
      for (int k = 0; k < K; ++k) p[k] = pb[(long long)i * K + k];
    }
    const int t = lb[i];
    float s = 0.f, pt = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const int on = p[k] > 0.5f, is = (k == t);
      c_pred[k] += on; c_true[k] += is; c_int[k] += on & is;
      s += p[k];
      if (is) pt = p[k];
    }
    // weighted CE exactly as the loss: renormalise, clip to [1e-7, 1 - 1e-7], -w_t log p_t
    const float q = fminf(fmaxf(pt / s, 1e-7f), 1.f - 1e-7f);
    l_sum += -(class_w && t < K ? class_w[t] : 1.f) * logf(q);
  }
  __shared__ int s_cnt[8][3 * K];
  __shared__ float s_loss[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    int a = c_int[k], c = c_pred[k], d = c_true[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o); c += __shfl_xor_sync(0xffffffffu, c, o); d += __shfl_xor_sync(0xffffffffu, d, o);
    }
    if (lane == 0) { s_cnt[warp][3 * k] = a; s_cnt[warp][3 * k + 1] = c; s_cnt[warp][3 * k + 2] = d; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) l_sum += __shfl_xor_sync(0xffffffffu, l_sum, o);
  if (lane == 0) s_loss[warp] = l_sum;
  __syncthreads();
  if (threadIdx.x < 3 * K) {
    unsigned long long t = 0;
    for (int wv = 0; wv < 8; ++wv) t += (unsigned long long)s_cnt[wv][threadIdx.x];
    if (t) atomicAdd(counts + (long long)b * 3 * K + threadIdx.x, t);
  }
  if (threadIdx.x == 32) {
    double t = 0;
    for (int wv = 0; wv < 8; ++wv) t += (double)s_loss[wv];
    atomicAdd(loss + b, t);
  }
}

int launch_eval_counts(const float *probs, const uint8_t *labels, int n, int h, int w, int K, const float *class_w,
                       unsigned long long *counts, double *loss, cudaStream_t st) {
  const int hw = h * w;
  dim3 grid(std::max(1, std::min((hw + 4095) / 4096, 64)), n);
  switch (K) {
#define EK(k) case k: eval_counts_kernel<k><<<grid, 256, 0, st>>>(probs, labels, hw, class_w, counts, loss); break;
    EK(2) EK(3) EK(4) EK(5) EK(6) EK(7) EK(8) EK(9) EK(10) EK(11) EK(12) EK(13) EK(14) EK(15) EK(16)
#undef EK
    default: set_error("eval counts: num_classes must be 2..16"); return 1;
  }
  OCTSEG_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------
__global__ void bn_fold_kernel(const float *bias, const float *gamma, const float *beta,
                               const float *mean, const float *var, float eps, int c, float *scale,
                               float *shift) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c) return;
  // same operation order as the oracle: inv = rsqrt(var+eps); y = (z-mean)*(inv*gamma)+beta
  float s = (1.0f / sqrtf(var[i] + eps)) * gamma[i];
  scale[i] = s;
  shift[i] = (bias[i] - mean[i]) * s + beta[i];
}

int launch_bn_fold(const float *bias, const float *gamma, const float *beta, const float *mean,
                   const float *var, float eps, int c, float *scale, float *shift, cudaStream_t st) {
  bn_fold_kernel<<<(c + 127) / 128, 128, 0, st>>>(bias, gamma, beta, mean, var, eps, c, scale, shift);
  OCTSEG_CUDA(cudaGetLastError());
  return 0;
}

// explicit instantiations
#define INST(T)                                                                                        \
  template int launch_conv_first<T>(const void *, int, int, int, int, int, const float *, int, int,    \
                                    int, const float *, const float *, int, View<T>, cudaStream_t);    \
  template int launch_conv_direct<T>(View<const T>, const float *, int, int, int, int, int,            \
                                     const float *, const float *, int, View<T>, cudaStream_t);        \
  template int launch_conv_direct_ex<T>(View<const T>, const float *, int, int, int, int, int, int,    \
                                        int, int, const float *, const float *, int, View<T>,          \
                                        cudaStream_t);                                                 \
  template int launch_maxpool2<T>(View<const T>, View<T>, cudaStream_t);                               \
  template int launch_head<T>(View<const T>, const float *, const float *, int, int, float *,          \
                              uint8_t *, cudaStream_t);
INST(float)
INST(__nv_bfloat16)
INST(__half)
template int launch_conv_first<SplitHalf>(const void *, int, int, int, int, int, const float *, int, int, int,
                                          const float *, const float *, int, View<SplitHalf>, cudaStream_t);

}  // namespace octseg
