// Shared helpers for the pcd_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/pcd_b200.h"

namespace pcd {

void set_error(const char* fmt, ...);
// number of kernels launched through this library since load (bench.py's gpu_launches)
extern unsigned long long g_launch_count;

#define PCD_CHECK_ARG(cond, ...)                 \
  do {                                           \
    if (!(cond)) {                               \
      ::pcd::set_error(__VA_ARGS__);             \
      return PCD_ERR_INVALID;                    \
    }                                            \
  } while (0)

#define PCD_CHECK_LAUNCH(name)                                              \
  do {                                                                      \
    ++::pcd::g_launch_count;                                                \
    cudaError_t e__ = cudaGetLastError();                                   \
    if (e__ != cudaSuccess) {                                               \
      ::pcd::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__)); \
      return PCD_ERR_CUDA;                                                  \
    }                                                                       \
  } while (0)

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

int num_sms();

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// exact-erf GELU, nn.GELU() default (reference models/transformer.py:57)
__device__ __forceinline__ float gelu_erf(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}

// exact-erf GELU evaluated as 0.5 x (1 + tanh(z P(z^2))), z = x/sqrt(2) clamped to +-3.3,
// with P fitted so that tanh(z P(z^2)) == erf(z) to 1.3e-4 (max |gelu error| 6.3e-5 over all x
// with an exact tanh) and the hardware tanh.approx.f32 (rel. error 2^-11).  Used only by the
// bf16 tensor-core GEMM epilogue, whose output is rounded to bf16 (rel. 2^-9) anyway: one MUFU
// and 6 FMA-pipe instructions per element instead of erff's ~25.
__device__ __forceinline__ float gelu_fast(float x) {
  // the polynomial in x^2 (clamped at 2 * 3.3^2) with 1/sqrt(2) folded into its coefficients; see gelu_fast2
  float xx = fminf(x * x, 21.78f);
  float p = fmaf(fmaf(-0.00204817f * 0.70710678118654752440f * 0.25f, xx, 0.10449843f * 0.70710678118654752440f * 0.5f),
                 xx, 1.12819195f * 0.70710678118654752440f);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(x * p));
  float hx = 0.5f * x;
  return fmaf(hx, t, hx);
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

}  // namespace pcd
