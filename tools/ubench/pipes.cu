// Issue-rate microbenchmark for the softmax instruction mix (one CTA of `warps` warps per SM):
// cycles per warp-instruction group for MUFU.EX2 alone and combined with F2FP / FFMA2 / FADD2 / FMNMX3.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t cvt2(float a, float b) { uint32_t r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a)); return r; }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float max3(float a, float b, float c) { float r; asm volatile("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
template <int MODE>
__global__ void k(float* out, int iters, long long* cyc) {
  float x[8], acc = 0.f; uint32_t pk = 0; uint64_t s2 = 0, y2 = 0x3f8000003f800000ull; float mx = 0.f;
  for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 1e-3f + i;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
      float a = x[i], b = x[i + 1];
      if (MODE != 5) { a = ex2(a); b = ex2(b); }           // 2 MUFU
      if (MODE == 1 || MODE == 4) pk ^= cvt2(a, b);      // + 1 F2FP
      if (MODE == 2 || MODE == 4) {                      // + 1 FFMA2 + 1 FADD2-like
        uint64_t p; asm volatile("mov.b64 %0, {%1, %2};" : "=l"(p) : "f"(a), "f"(b));
        s2 = fma2(p, y2, s2); y2 = fma2(p, y2, y2);
      }
      if (MODE == 3 || MODE == 4) mx = max3(mx, a, b);   // + 1 FMNMX3
      if (MODE == 5) { pk ^= cvt2(a, b); pk += cvt2(b, a); }  // F2FP only (2 per pair)
      x[i] = a * 0.5f; x[i + 1] = b * 0.25f;
    }
  }
  long long t1 = clock64();
  for (int i = 0; i < 8; ++i) acc += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + pk + (float)s2 + (float)y2 + mx;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int MODE> void run(const char* name, int warps) {
  float* out; long long* cyc; cudaMalloc(&out, 4 * 148 * 1024); cudaMalloc(&cyc, 8);
  int iters = 4000;
  k<MODE><<<148, warps * 32>>>(out, 10, cyc); cudaDeviceSynchronize();
  k<MODE><<<148, warps * 32>>>(out, iters, cyc); cudaDeviceSynchronize();
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  // per SMSP: warps/4 warps each issue iters*4 pairs
  double pairs_per_smsp = (double)(warps / 4.0) * iters * 4;
  printf("%-34s warps/SM=%2d  %.2f cycles per (2 MUFU [+extras]) per scheduler  [%s]\n", name, warps, c / pairs_per_smsp, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out); cudaFree(cyc);
}
int main() {
  for (int w : {4, 8, 16}) {
    run<0>("2 MUFU", w); run<1>("2 MUFU + F2FP", w); run<2>("2 MUFU + 2 FFMA2", w); run<3>("2 MUFU + FMNMX3", w);
    run<4>("2 MUFU + F2FP + 2 FFMA2 + FMNMX3", w); run<5>("2 F2FP (no MUFU)", w);
  }
  return 0;
}
