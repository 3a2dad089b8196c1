// MUFU throughput probe: ex2.approx f32 vs f16x2 vs bf16x2 (results per clock per SM).  Tools only.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(512) k(uint32_t* out, int iters, long long* cyc) {
  uint32_t a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = (MODE == 0) ? __float_as_uint(-0.001f * (threadIdx.x + i + 1)) : 0xB000B400u + i + threadIdx.x;  // small negative halves
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+r"(a[i]));
      if (MODE == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(a[i]));
      if (MODE == 2) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(a[i]));
      if (MODE == 3) asm volatile("tanh.approx.f32 %0, %0;" : "+r"(a[i]));
      if (MODE == 4) asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(a[i]));
    }
  }
  long long t1 = clock64();
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s ^= a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int MODE> void run(const char* name, int per_instr) {
  uint32_t* out; long long* cyc; cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 8);
  const int iters = 4096;
  k<MODE><<<148, 512>>>(out, iters, cyc); cudaDeviceSynchronize();
  k<MODE><<<148, 512>>>(out, iters, cyc); cudaDeviceSynchronize();
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  double instr = 16.0 * iters * 8;  // warp instructions per SM
  printf("%-22s %8lld cycles  %.2f cycles/warp-instr/SM  %.1f results/clk/SM  (%s)\n", name, c, c / instr, instr * 32 * per_instr / c, cudaGetErrorString(cudaGetLastError()));
}
int main() {
  run<0>("ex2.f32", 1); run<1>("ex2.f16x2", 2); run<2>("ex2.ftz.bf16x2", 2); run<3>("tanh.f32", 1); run<4>("tanh.bf16x2", 2);
  return 0;
}
