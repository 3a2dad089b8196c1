// TMEM <-> register bandwidth of tcgen05.ld / tcgen05.st (32x32b.x32) on one SM and chip-wide.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
#define LD32(taddr, r) asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];" \
  : "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]),"=r"(r[4]),"=r"(r[5]),"=r"(r[6]),"=r"(r[7]),"=r"(r[8]),"=r"(r[9]),"=r"(r[10]),"=r"(r[11]),"=r"(r[12]),"=r"(r[13]),"=r"(r[14]),"=r"(r[15]),"=r"(r[16]),"=r"(r[17]),"=r"(r[18]),"=r"(r[19]),"=r"(r[20]),"=r"(r[21]),"=r"(r[22]),"=r"(r[23]),"=r"(r[24]),"=r"(r[25]),"=r"(r[26]),"=r"(r[27]),"=r"(r[28]),"=r"(r[29]),"=r"(r[30]),"=r"(r[31]) : "r"(taddr) : "memory")
#define ST32(taddr, r) asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" \
  :: "r"(taddr),"r"(r[0]),"r"(r[1]),"r"(r[2]),"r"(r[3]),"r"(r[4]),"r"(r[5]),"r"(r[6]),"r"(r[7]),"r"(r[8]),"r"(r[9]),"r"(r[10]),"r"(r[11]),"r"(r[12]),"r"(r[13]),"r"(r[14]),"r"(r[15]),"r"(r[16]),"r"(r[17]),"r"(r[18]),"r"(r[19]),"r"(r[20]),"r"(r[21]),"r"(r[22]),"r"(r[23]),"r"(r[24]),"r"(r[25]),"r"(r[26]),"r"(r[27]),"r"(r[28]),"r"(r[29]),"r"(r[30]),"r"(r[31]) : "memory")

template <int MODE>  // 0: ld, 1: st, 2: ld with wait after each
__global__ void __launch_bounds__(256) k(uint32_t* out, int iters, int nwarps) {
  __shared__ uint32_t slot;
  int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  uint32_t base = slot;
  uint32_t r[32];
  for (int i = 0; i < 32; ++i) r[i] = threadIdx.x + i;
  uint32_t acc = 0;
  if (warp < nwarps) {
    uint32_t taddr = base + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 64;
    for (int it = 0; it < iters; ++it) {
      if (MODE == 1) {
        ST32(taddr + (it & 1) * 32, r);
      } else {
        LD32(taddr + (it & 1) * 32, r);
        if (MODE == 2) asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      }
      if ((it & 7) == 7) {
        if (MODE == 1) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        else asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        acc += r[it & 31];
      }
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + r[3];
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(256));
}
template <int MODE>
void run(const char* name, int blocks, int nwarps) {
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  uint32_t* out; cudaMalloc(&out, 4 * blocks * 256);
  int iters = 20000;
  k<MODE><<<blocks, 256>>>(out, 100, nwarps);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<blocks, 256>>>(out, iters, nwarps);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double bytes_per_sm = (double)nwarps * iters * 4096;  // one CTA per SM
  printf("%-22s warps=%d  %.3f ms  %.1f B/clk/SM (at %.2f GHz nominal)  err=%s\n", name, nwarps, ms,
         bytes_per_sm / (ms * 1e-3 * clk * 1e3), clk / 1e6, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
}
int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  for (int w : {1, 4, 8}) {
    run<0>("tcgen05.ld x32", sms, w);
    run<2>("tcgen05.ld x32 +wait", sms, w);
    run<1>("tcgen05.st x32", sms, w);
  }
  return 0;
}
