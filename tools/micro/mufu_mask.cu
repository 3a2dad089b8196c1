// Does a MUFU instruction with few active lanes occupy the pipe for fewer cycles?  Tools only.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(512) k(uint32_t* out, int iters, int active, long long* cyc) {
  uint32_t a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = __float_as_uint(-0.001f * (threadIdx.x + i + 1));
  __syncthreads();
  long long t0 = clock64();
  if ((threadIdx.x & 31) < active) {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+r"(a[i]));
    }
  }
  __syncthreads();
  long long t1 = clock64();
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s ^= a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  uint32_t* out; long long* cyc; cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 8);
  const int iters = 4096;
  for (int active : {32, 16, 8, 4, 2, 1}) {
    k<<<148, 512>>>(out, iters, active, cyc); cudaDeviceSynchronize();
    k<<<148, 512>>>(out, iters, active, cyc); cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    double instr = 16.0 * iters * 8;
    printf("active lanes %2d: %8lld cycles  %.2f cycles/warp-instr/SM (4 schedulers)\n", active, c, c / instr);
  }
  return 0;
}
