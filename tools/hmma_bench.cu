// Micro-benchmark: warp-level mma.sync.m16n8k16 bf16 (SASS HMMA.16816) throughput per SM on sm_100a, with and without
// the ldmatrix.x4 that feeds it -- the bound of the narrow-layer weight-gradient kernels (csrc/wgrad_mma.cu).
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/hmma_bench.cu -o tools/hmma_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int NACC, int LD>
__global__ void __launch_bounds__(1024, 1) k(int reps, float *sink, long long *clk) {
  __shared__ __align__(16) uint8_t sm[32 * 1024];
  for (int i = threadIdx.x; i < 32 * 1024 / 4; i += blockDim.x) ((uint32_t *)sm)[i] = 0x3c003c00u;
  __syncthreads();
  float acc[NACC][4];
#pragma unroll
  for (int i = 0; i < NACC; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  uint32_t a[4] = {0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u}, b[2] = {0x3f803f80u, 0x3f803f80u};
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(sm) + (threadIdx.x & 1023) * 16;
  const long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
      if (LD) asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                           : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]) : "r"(base + (uint32_t)(((r + i) & 15) * 1024)));
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(acc[i][0]), "+f"(acc[i][1]), "+f"(acc[i][2]), "+f"(acc[i][3])
                   : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += acc[i][0] + acc[i][3];
  if (s == 12345.f) sink[0] = s;
  if (blockIdx.x == 0 && threadIdx.x == 0) clk[0] = t1 - t0;
}

template <int NACC, int LD>
void run(int warps, const char *what) {
  float *sink; long long *clk, h;
  cudaMalloc(&sink, 4); cudaMalloc(&clk, 8);
  const int reps = 2000;
  k<NACC, LD><<<148, warps * 32>>>(reps, sink, clk);
  k<NACC, LD><<<148, warps * 32>>>(reps, sink, clk);
  cudaDeviceSynchronize();
  cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
  const double per = (double)h / ((double)reps * NACC * warps);
  printf("%-28s warps/SM %2d  acc %d : %.2f clk per MMA per SM  (%.0f bf16 FMA/clk/SM)\n", what, warps, NACC, per, 2048.0 / per);
  cudaFree(sink); cudaFree(clk);
}

int main() {
  for (int w : {4, 8, 16, 32}) run<8, 0>(w, "mma.sync only");
  for (int w : {4, 8, 16, 32}) run<8, 1>(w, "ldmatrix.x4 + mma.sync");
  for (int w : {16}) run<4, 1>(w, "ldmatrix.x4 + mma.sync");
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
