// Micro-benchmarks of per-SM instruction throughput on B200 (ex2, bf16 pack, FFMA, TMEM ld/st).
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdio.h>
#include <stdint.h>

#define ITERS 4096
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, float seed) {
  float a[8];
  for (int i = 0; i < 8; ++i) a[i] = seed + threadIdx.x * 1e-3f + i;
  uint32_t acc = 0;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) {  // ex2 only
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      } else if (MODE == 1) {  // ffma only
        a[i] = fmaf(a[i], 1.0001f, 0.5f);
      } else if (MODE == 2) {  // ffma + ex2
        a[i] = fmaf(a[i], 1.0001f, -0.5f);
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      } else if (MODE == 3) {  // softmax-like: ffma, ex2, fadd, and a pack per pair
        a[i] = fmaf(a[i], 1.0001f, -0.5f);
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      } else if (MODE == 5) {  // packed half2 ex2: two exponentials per MUFU op
        uint32_t h = __float_as_uint(a[i]);
        asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h));
        a[i] = __uint_as_float(h);
      } else if (MODE == 6) {  // softmax-like with packed ex2: 2 ffma + cvt.f16x2 + ex2.f16x2 + hadd2
        float x0 = fmaf(a[i], 1.0001f, -0.5f), x1 = fmaf(a[i], 0.9999f, -0.25f);
        uint32_t h;
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(x1), "f"(x0));
        asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h));
        asm volatile("add.rn.f16x2 %0, %0, %1;" : "+r"(acc) : "r"(h));
        a[i] += 1e-3f;
      } else if (MODE == 7) {  // tanh.approx.f32
        asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a[i]));
      } else if (MODE == 8) {  // tanh.approx.f16x2
        uint32_t h = __float_as_uint(a[i]);
        asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(h));
        a[i] = __uint_as_float(h);
      } else if (MODE == 4) {  // pack only
        __nv_bfloat162 p = __floats2bfloat162_rn(a[i], a[(i + 1) & 7]);
        acc ^= *reinterpret_cast<uint32_t*>(&p);
        a[i] += 1.0f;
      }
    }
    if (MODE == 3) {
#pragma unroll
      for (int i = 0; i < 8; i += 2) {
        __nv_bfloat162 p = __floats2bfloat162_rn(a[i], a[i + 1]);
        acc ^= *reinterpret_cast<uint32_t*>(&p);
      }
    }
  }
  float s = 0;
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + __uint_as_float(acc & 0xff);
}

template <int MODE>
void run(const char* name, int blocks_per_sm, float ops_per_iter_per_thread) {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  float* out; cudaMalloc(&out, sizeof(float) * sms * blocks_per_sm * 256);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<sms * blocks_per_sm, 256>>>(out, 0.5f);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  for (int r = 0; r < 5; ++r) k<MODE><<<sms * blocks_per_sm, 256>>>(out, 0.5f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
  double ops = (double)sms * blocks_per_sm * 256 * ITERS * ops_per_iter_per_thread;
  printf("%-28s blocks/SM=%d  %.3f ms  %.2f Gop/s  = %.2f ops/clk/SM @%.2f GHz(nominal max)\n", name, blocks_per_sm, ms,
         ops / ms / 1e6, ops / (ms * 1e-3) / sms / (clk * 1e3), clk / 1e6);
  cudaFree(out);
}

int main() {
  for (int b : {2}) {
    run<0>("ex2", b, 8);
    run<1>("ffma", b, 8);
    run<2>("ffma+ex2 (per pair)", b, 8);
    run<3>("ffma+ex2+pack (per exp)", b, 8);
    run<4>("bf16x2 pack (per pack)", b, 8);
    run<5>("ex2.f16x2 (per MUFU op)", b, 8);
    run<6>("2ffma+cvt+ex2.f16x2+hadd2 (per pair)", b, 8);
    run<7>("tanh.f32", b, 8);
    run<8>("tanh.f16x2 (per MUFU op)", b, 8);
  }
  return 0;
}
