// Micro-benchmark: cycles per tcgen05.mma (M=128, K=16, bf16) as a function of N, smem layout
// (no-swizzle core-matrix vs 128B swizzle), strides and number of independent accumulators.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/mma_bench.cu -o tools/mma_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct P { int n_cols, layout, lbo, sbo, n_acc, reps, a_step, issuers; };

__global__ void __launch_bounds__(256, 1) k(P p, long long *out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[8];
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) ((uint32_t *)smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar[i])), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;
  if (lane == 0 && warp < p.issuers) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.n_cols >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t a_base = smem_u32(smem), b_base = smem_u32(smem + 128 * 1024);
    const uint64_t hi = ((uint64_t)((p.lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((p.sbo >> 4) & 0x3FFF) << 32) | (1ull << 46) |
                        ((uint64_t)p.layout << 61);
    const uint64_t db = (uint64_t)((b_base >> 4) & 0x3FFF) | ((uint64_t)((p.n_cols * 16 >> 4) & 0x3FFF) << 16) |
                        ((uint64_t)(128 >> 4) << 32) | (1ull << 46);
    const int acc_per = p.n_acc / p.issuers > 0 ? p.n_acc / p.issuers : 1;
    long long t0 = clock64();
    for (int r = 0; r < p.reps; ++r) {
      const int acc = warp * acc_per + (r % acc_per);
      const uint64_t da = hi | (uint64_t)(((a_base + (uint32_t)((r * p.a_step) & 0xFFFF)) >> 4) & 0x3FFF);
      const uint32_t d = tmem_base + (uint32_t)(acc * p.n_cols);
      asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, q;\n\t}"
                   ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(r >= acc_per ? 1u : 0u) : "memory");
    }
    long long t1 = clock64();
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[warp])) : "memory");
    uint32_t done = 0;
    while (!done)
      asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.u32 %0, 1, 0, q;\n\t}"
                   : "=r"(done) : "r"(smem_u32(&bar[warp])), "r"(0) : "memory");
    long long t2 = clock64();
    if (blockIdx.x == 0 && warp == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
}

int main() {
  long long *d_out, h[2];
  cudaMalloc(&d_out, 16);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  struct { const char *name; P p; } cases[] = {
      // name                         n   lay lbo    sbo   acc reps a_step issuers
      {"nosw N16 16acc 1 issuer",    {16, 0, 16,   1056, 16, 512, 128,  1}},
      {"nosw N16 16acc 2 issuers",   {16, 0, 16,   1056, 16, 512, 128,  2}},
      {"nosw N16 16acc 4 issuers",   {16, 0, 16,   1056, 16, 512, 128,  4}},
      {"nosw N16 16acc 8 issuers",   {16, 0, 16,   1056, 16, 512, 128,  8}},
      {"nosw N16 8acc 8 issuers",    {16, 0, 16,   1056, 8,  512, 128,  8}},
      {"nosw N32 8acc 8 issuers",    {32, 0, 16,   1056, 8,  512, 128,  8}},
      {"nosw N32 8acc 4 issuers",    {32, 0, 16,   1056, 8,  512, 128,  4}},
      {"nosw N64 8acc 8 issuers",    {64, 0, 16,   1056, 8,  512, 128,  8}},
      {"nosw N64 4acc 4 issuers",    {64, 0, 16,   1056, 4,  512, 128,  4}},
      {"nosw N128 4acc 4 issuers",   {128,0, 16,   1056, 4,  512, 128,  4}},
      {"nosw N256 2acc 2 issuers",   {256,0, 16,   1056, 2,  512, 128,  2}},
      {"sw128 N16 16acc 8 issuers",  {16, 2, 16,   1024, 16, 512, 32,   8}},
      {"sw128 N64 8acc 8 issuers",   {64, 2, 16,   1024, 8,  512, 32,   8}},
  };
  for (auto &c : cases) {
    for (int rep = 0; rep < 2; ++rep) {
      k<<<148, 256, 192 * 1024>>>(c.p, d_out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%s: CUDA error %s\n", c.name, cudaGetErrorString(e)); return 1; }
    }
    cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost);
    printf("%-30s issue %6.1f clk/mma   complete %6.1f clk/mma\n", c.name, (double)h[0] / c.p.reps, (double)h[1] / c.p.reps);
  }
  return 0;
}
