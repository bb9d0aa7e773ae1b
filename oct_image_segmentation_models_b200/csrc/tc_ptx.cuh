// PTX wrappers shared by the tcgen05 kernels (conv_tc.cu, wgrad_tc.cu): mbarrier, TMA, tcgen05.mma / ld / commit.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace octseg {

// ----------------------------------------------------------------------------------
// PTX wrappers
// ----------------------------------------------------------------------------------
static __device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
static __device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
static __device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
static __device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Bounded wait: a broken pipeline must end in an error code, never in a hung GPU.
static __device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity, int *status, int code,
                                          long long *waited = nullptr) {
  const long long t0 = clock64();
  for (uint32_t it = 1;; ++it) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) {
      if (waited) *waited += clock64() - t0;
      return true;
    }
    if ((it & 0x3FFu) == 0) {
      // ~2 s at 2 GHz, or another role already gave up
      if (clock64() - t0 > 4000000000ll || *reinterpret_cast<volatile int *>(status) != 0) break;
    }
  }
  atomicCAS(status, 0, code);
  return false;
}
static __device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0,
                                            int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
static __device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0,
                                            int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
static __device__ __forceinline__ void bulk_load(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}
// One lane of a fully converged warp.  tcgen05.mma / TMA take their operands in UNIFORM registers:
// the issuing code must stay warp-uniform (all 32 lanes compute the same descriptors) and only the
// instruction itself is predicated on the elected lane -- issuing from inside an `if (lane == 0)`
// region makes the compiler wrap every MMA in a VOTEU/ELECT/R2UR waterfall loop (~20 extra
// instructions, measured 220 clk per MMA instead of 39).
static __device__ __forceinline__ uint32_t elect_one_sync() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
      "elect.sync rx|px, %1;\n\t"
      "@px mov.s32 %0, 1;\n\t}"
      : "+r"(pred)
      : "r"(0xFFFFFFFFu));
  return pred;
}
static __device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                          uint32_t idesc, uint32_t accumulate) {
  if (elect_one_sync()) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
static __device__ __forceinline__ void umma_commit(uint32_t bar) {
  if (elect_one_sync()) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
                 : "memory");
  }
}
static __device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
                 "=r"(r[7])
               : "r"(taddr));
}
static __device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr));
}
// fp32 -> error-compensated fp16 pair: hi = rn(o), lo' = rn((o - hi) * 2^11); o == hi + lo' * 2^-11 to 2^-22 relative
// (o - hi is exact in fp32, so is the power-of-two scaling)
static __device__ __forceinline__ void split_pack8(const float (&o)[8], uint4 &hi, uint4 &lo) {
  uint32_t *h = reinterpret_cast<uint32_t *>(&hi), *l = reinterpret_cast<uint32_t *>(&lo);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const __half2 h2 = __floats2half2_rn(o[2 * k], o[2 * k + 1]);
    const float2 hf = __half22float2(h2);
    const __half2 l2 = __floats2half2_rn((o[2 * k] - hf.x) * 2048.f, (o[2 * k + 1] - hf.y) * 2048.f);
    h[k] = *reinterpret_cast<const uint32_t *>(&h2);
    l[k] = *reinterpret_cast<const uint32_t *>(&l2);
  }
}
// Sum of 8 per-thread values over the 32 lanes of a warp with 7 shuffles (instead of 40): every step halves the
// number of values a lane carries.  Returns, in EVERY lane, the warp total of value index
// ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1)   (the four lanes of a quad hold the same sum).
static __device__ __forceinline__ float warp_sum8(const float (&v)[8], int lane) {
  bool hi = (lane & 16) != 0;
  float a[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float send = hi ? v[i] : v[i + 4], keep = hi ? v[i + 4] : v[i];
    a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
  hi = (lane & 8) != 0;
  float b[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float send = hi ? a[i] : a[i + 2], keep = hi ? a[i + 2] : a[i];
    b[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
  hi = (lane & 4) != 0;
  const float send = hi ? b[0] : b[1], keep = hi ? b[1] : b[0];
  float c = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  c += __shfl_xor_sync(0xffffffffu, c, 2);
  c += __shfl_xor_sync(0xffffffffu, c, 1);
  return c;
}
static __device__ __forceinline__ int warp_sum8_index(int lane) { return ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1); }
static __device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
static __device__ __forceinline__ float fast_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
static __device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
static __device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
static __device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }


}  // namespace octseg
