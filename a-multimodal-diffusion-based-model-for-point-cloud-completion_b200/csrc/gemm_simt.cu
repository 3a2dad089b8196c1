// fp32 CUDA-core GEMM: C[M,N] = epi(A[M,K] W[N,K]^T + bias).
// Used by the fp32 parity mode (north-star tolerance 1e-4) and for the tiny
// per-sequence layers (time MLP, CLIP projections) where M is the batch size.
// 128x128x16 block tile, 8x8 register tile, 256 threads, register-prefetched
// double buffering; operands are both K-contiguous (nn.Linear layout) and are
// transposed into shared memory as [k][m] so the inner product reads are
// conflict-free float4 broadcasts.
#include "common.cuh"

namespace pcd {

constexpr int BM = 128, BN = 128, BK = 16, TM = 8, TN = 8;

template <int EPI>
__global__ void __launch_bounds__(256) gemm_f32_kernel(const float* __restrict__ A, int lda,
                                                       const float* __restrict__ W, int ldw,
                                                       const float* __restrict__ bias,
                                                       const float* residual, int ldr,
                                                       float* C, int ldc, int M, int N,
                                                       int K) {
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Ws[2][BK][BN + 4];
  int tid = threadIdx.x;
  int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  // loader mapping: 128 rows x 16 k = 512 float4; each thread loads 2 float4 per operand
  int lrow = tid >> 2;          // 0..63
  int lk = (tid & 3) * 4;       // 0,4,8,12
  float4 ra[2], rw[2];
  auto gload = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      int r = lrow + i * 64;
      int gm = m0 + r, gn = n0 + r, gk = k0 + lk;
      ra[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      rw[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gm < M && gk < K) ra[i] = *reinterpret_cast<const float4*>(A + (size_t)gm * lda + gk);
      if (gn < N && gk < K) rw[i] = *reinterpret_cast<const float4*>(W + (size_t)gn * ldw + gk);
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      int r = lrow + i * 64;
      As[buf][lk + 0][r] = ra[i].x; As[buf][lk + 1][r] = ra[i].y;
      As[buf][lk + 2][r] = ra[i].z; As[buf][lk + 3][r] = ra[i].w;
      Ws[buf][lk + 0][r] = rw[i].x; Ws[buf][lk + 1][r] = rw[i].y;
      Ws[buf][lk + 2][r] = rw[i].z; Ws[buf][lk + 3][r] = rw[i].w;
    }
  };
  int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads, each 8x8 outputs (split 4+4 for banks)
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  int nk = (K + BK - 1) / BK;
  gload(0);
  sstore(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    int buf = kt & 1;
    if (kt + 1 < nk) gload((kt + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[TM], w[TN];
      float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      float4 w0 = *reinterpret_cast<const float4*>(&Ws[buf][k][tx * 4]);
      float4 w1 = *reinterpret_cast<const float4*>(&Ws[buf][k][64 + tx * 4]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
      w[0] = w0.x; w[1] = w0.y; w[2] = w0.z; w[3] = w0.w; w[4] = w1.x; w[5] = w1.y; w[6] = w1.z; w[7] = w1.w;
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      sstore(buf ^ 1);
      __syncthreads();
    }
  }
  // epilogue
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int gm = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (gm >= M) continue;
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      int gn = n0 + jh * 64 + tx * 4;
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float t = acc[i][jh * 4 + j];
        int n = gn + j;
        if (n < N) {
          if (bias) t += bias[n];
          if (EPI == PCD_EPI_BIAS_GELU) t = gelu_erf(t);
          if (EPI == PCD_EPI_BIAS_RESIDUAL) t += residual[(size_t)gm * ldr + n];
        }
        v[j] = t;
      }
      if (gn + 3 < N && (ldc & 3) == 0) {
        *reinterpret_cast<float4*>(C + (size_t)gm * ldc + gn) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (gn + j < N) C[(size_t)gm * ldc + gn + j] = v[j];
      }
    }
  }
}

}  // namespace pcd

using namespace pcd;

extern "C" int pcd_gemm_f32(const float* A, int lda, const float* W, int ldw, const float* bias,
                            const float* residual, int ldr, float* C, int ldc, int M, int N, int K,
                            int epilogue, void* stream) {
  PCD_CHECK_ARG(M > 0 && N > 0 && K > 0, "gemm_f32: empty problem");
  PCD_CHECK_ARG(K % 4 == 0 && lda % 4 == 0 && ldw % 4 == 0, "gemm_f32: K, lda, ldw must be multiples of 4 (K=%d lda=%d ldw=%d)", K, lda, ldw);
  PCD_CHECK_ARG(epilogue != PCD_EPI_BIAS_RESIDUAL || residual != nullptr, "gemm_f32: residual missing");
  dim3 grid(ceil_div(N, BN), ceil_div(M, BM)), block(256);
  cudaStream_t st = (cudaStream_t)stream;
  switch (epilogue) {
    case PCD_EPI_BIAS: gemm_f32_kernel<PCD_EPI_BIAS><<<grid, block, 0, st>>>(A, lda, W, ldw, bias, residual, ldr, C, ldc, M, N, K); break;
    case PCD_EPI_BIAS_GELU: gemm_f32_kernel<PCD_EPI_BIAS_GELU><<<grid, block, 0, st>>>(A, lda, W, ldw, bias, residual, ldr, C, ldc, M, N, K); break;
    case PCD_EPI_BIAS_RESIDUAL: gemm_f32_kernel<PCD_EPI_BIAS_RESIDUAL><<<grid, block, 0, st>>>(A, lda, W, ldw, bias, residual, ldr, C, ldc, M, N, K); break;
    default: PCD_CHECK_ARG(false, "gemm_f32: unknown epilogue %d", epilogue);
  }
  PCD_CHECK_LAUNCH("gemm_f32");
  return PCD_OK;
}
