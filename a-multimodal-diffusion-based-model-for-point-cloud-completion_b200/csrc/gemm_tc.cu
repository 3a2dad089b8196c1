// bf16 GEMM on the 5th-generation tensor cores:  C[M,N] = epi(A[M,K] W[N,K]^T + bias)
//
// Persistent, warp-specialised kernel (one CTA per SM, 320 threads):
//   warp 0  : TMA producer   -- A (128 x 64) and W (BN x 64) bf16 tiles, 128B swizzle,
//                               STAGES-deep mbarrier ring
//   warp 1  : MMA issuer     -- one elected lane issues tcgen05.mma (M=128, N=BN, K=16),
//                               fp32 accumulators double-buffered in TMEM (2 x BN columns)
//   warps 2-9: epilogue      -- 8 warps: TMEM lane quarter = warp % 4, column half = (warp-2) / 4;
//                               double-buffered tcgen05.ld 32x32b.x32 (thread = row), bias staged
//                               in shared memory per tile, GELU / fp32 residual, 16-byte stores
//                               straight from registers
// Both operands are K-contiguous (activations [M,K], nn.Linear weights [N,K]) so no
// transposes are needed.  Tails in M, N, K are handled by TMA zero fill + store predicates.
#include "common.cuh"
#include "tc_sm100.cuh"

namespace pcd {

using namespace tc;

constexpr int G_BM = 128, G_BK = 64;
constexpr int G_THREADS = 320, G_EPI_WARPS = 8;

template <int BN>
struct GemmCfg {
  static constexpr int STAGES = (BN == 256) ? 4 : 6;
  static constexpr int A_BYTES = G_BM * G_BK * 2;   // 16 KB
  static constexpr int B_BYTES = BN * G_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int TMEM_COLS = 2 * BN;          // 512 / 256 / 128: powers of two
  static constexpr int BIAS_BYTES = 2 * BN * 4;      // per accumulator stage
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + BIAS_BYTES;
};

template <int BN, int EPI, bool OUT_BF16>
__global__ void __launch_bounds__(G_THREADS, 1)
gemm_bf16_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                    const float* __restrict__ bias, const float* residual, int ldr,
                    void* Cout, int ldc, int M, int N, int K) {
  using Cfg = GemmCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ unsigned char smem_raw[];
  // 1024-byte alignment is required by the 128B swizzle atom (8 rows x 128 B)
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* tiles = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* full = bars;                    // [STAGES]  TMA -> MMA
  uint64_t* empty = bars + STAGES;          // [STAGES]  MMA -> TMA
  uint64_t* acc_full = bars + 2 * STAGES;   // [2]       MMA -> epilogue
  uint64_t* acc_empty = acc_full + 2;       // [2]       epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* sbias = reinterpret_cast<float*>(smem + STAGES * Cfg::STAGE_BYTES + 256);  // [2][BN]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_m = (M + G_BM - 1) / G_BM, num_n = (N + BN - 1) / BN;
  const int num_tiles = num_m * num_n;
  const int num_kb = (K + G_BK - 1) / G_BK;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmA);
    prefetch_tensormap(&tmW);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], G_EPI_WARPS);  // one arrival per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------- TMA producer -------------------------
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        int m_blk = t / num_n, n_blk = t % num_n;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          unsigned char* sa = tiles + stage * Cfg::STAGE_BYTES;
          unsigned char* sb = sa + Cfg::A_BYTES;
          mbar_expect_tx(&full[stage], Cfg::STAGE_BYTES);
          tma_load_2d(sa, &tmA, &full[stage], kb * G_BK, m_blk * G_BM);
          tma_load_2d(sb, &tmW, &full[stage], kb * G_BK, n_blk * BN);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // -------------------------- MMA issuer --------------------------
    constexpr uint32_t idesc = idesc_bf16_f32(G_BM, BN, 0);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      mbar_wait(&acc_empty[as], aphase ^ 1);
      tcgen05_fence_after();
      const uint32_t d_tmem = tmem_base + as * BN;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full[stage], phase);
        tcgen05_fence_after();
        if (lane == 0) {
          const uint32_t sa = smem_u32(tiles + stage * Cfg::STAGE_BYTES);
          const uint64_t adesc = smem_desc_sw128(sa);
          const uint64_t bdesc = smem_desc_sw128(sa + Cfg::A_BYTES);
#pragma unroll
          for (int k = 0; k < G_BK / 16; ++k) {
            // advance 16 bf16 = 32 B along K inside the 128B swizzle row: +2 in (addr >> 4)
            umma_bf16_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          }
          umma_commit(&empty[stage]);                       // smem slot free when MMAs retire
          if (kb == num_kb - 1) umma_commit(&acc_full[as]); // accumulator ready
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // --------------------------- epilogue ----------------------------
    const int ew = warp - 2;           // 0..7
    const int quarter = warp & 3;      // TMEM lanes [32*quarter, +32) are accessible to this warp
    const int half = ew >> 2;          // which half of the tile's columns this warp drains
    constexpr int HALF_N = BN / 2;
    constexpr int NCH = HALF_N / 32;   // 32-column chunks per thread (BN=64 -> one 32-col chunk)
    static_assert(BN % 64 == 0, "BN must be a multiple of 64");
    const int etid = threadIdx.x - 64;  // 0..255
    int it = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int m_blk = t / num_n, n_blk = t % num_n;
      const int row = m_blk * G_BM + quarter * 32 + lane;
      // stage this tile's bias slice in shared memory (double-buffered by accumulator stage)
      float* sb = sbias + as * BN;
      if (etid < BN) {
        const int n = n_blk * BN + etid;
        sb[etid] = (bias != nullptr && n < N) ? __ldg(bias + n) : 0.f;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      mbar_wait(&acc_full[as], aphase);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + as * BN + half * HALF_N;
      uint32_t rbuf[2][32];
      tmem_ld_32x32b_x32(taddr, rbuf[0]);
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        uint32_t* r = rbuf[c & 1];
        const int ccol = half * HALF_N + c * 32;      // column inside the tile
        const int col0 = n_blk * BN + ccol;           // global column
        const bool active = row < M && col0 < N;
        const bool full_cols = (col0 + 32 <= N);
        float4 res[8];
        if (EPI == PCD_EPI_BIAS_RESIDUAL && active && full_cols) {
          // issue the residual loads before waiting on TMEM so both latencies overlap
          const float4* rp = reinterpret_cast<const float4*>(residual + (size_t)row * ldr + col0);
#pragma unroll
          for (int j = 0; j < 8; ++j) res[j] = rp[j];
        }
        tmem_ld_wait();
        if (c + 1 < NCH) tmem_ld_32x32b_x32(taddr + (c + 1) * 32, rbuf[(c + 1) & 1]);
        if (active) {
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 b4 = *reinterpret_cast<const float4*>(sb + ccol + j);
            v[j] = __uint_as_float(r[j]) + b4.x;
            v[j + 1] = __uint_as_float(r[j + 1]) + b4.y;
            v[j + 2] = __uint_as_float(r[j + 2]) + b4.z;
            v[j + 3] = __uint_as_float(r[j + 3]) + b4.w;
          }
          if (EPI == PCD_EPI_BIAS_GELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = gelu_fast(v[j]);
          }
          if (full_cols) {
            if (EPI == PCD_EPI_BIAS_RESIDUAL) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                v[4 * j] += res[j].x; v[4 * j + 1] += res[j].y; v[4 * j + 2] += res[j].z; v[4 * j + 3] += res[j].w;
              }
            }
            if (OUT_BF16) {
              uint16_t* cp = reinterpret_cast<uint16_t*>(Cout) + (size_t)row * ldc + col0;
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                uint4 p = make_uint4(pack_bf16x2(v[j], v[j + 1]), pack_bf16x2(v[j + 2], v[j + 3]),
                                     pack_bf16x2(v[j + 4], v[j + 5]), pack_bf16x2(v[j + 6], v[j + 7]));
                *reinterpret_cast<uint4*>(cp + j) = p;
              }
            } else {
              float* cp = reinterpret_cast<float*>(Cout) + (size_t)row * ldc + col0;
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4*>(cp + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            }
          } else {
            // ragged N tail: predicated scalar path (fully unrolled: registers only)
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              if (col0 + j < N) {
                float t2 = v[j];
                if (EPI == PCD_EPI_BIAS_RESIDUAL) t2 += residual[(size_t)row * ldr + col0 + j];
                if (OUT_BF16)
                  reinterpret_cast<__nv_bfloat16*>(Cout)[(size_t)row * ldc + col0 + j] = __float2bfloat16_rn(t2);
                else
                  reinterpret_cast<float*>(Cout)[(size_t)row * ldc + col0 + j] = t2;
              }
            }
          }
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[as]);
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

template <int BN, int EPI, bool OUT_BF16>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmW, const float* bias,
                       const float* residual, int ldr, void* C, int ldc, int M, int N, int K,
                       cudaStream_t st) {
  using Cfg = GemmCfg<BN>;
  auto kern = gemm_bf16_tc_kernel<BN, EPI, OUT_BF16>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) {
      set_error("gemm_bf16: cudaFuncSetAttribute(%d): %s", Cfg::SMEM_BYTES, cudaGetErrorString(e));
      return PCD_ERR_CUDA;
    }
    attr_set = true;
  }
  int tiles = ceil_div(M, G_BM) * ceil_div(N, BN);
  int grid = tiles < num_sms() ? tiles : num_sms();
  kern<<<grid, G_THREADS, Cfg::SMEM_BYTES, st>>>(tmA, tmW, bias, residual, ldr, C, ldc, M, N, K);
  PCD_CHECK_LAUNCH("gemm_bf16");
  return PCD_OK;
}

template <int BN>
static int dispatch_epi(const CUtensorMap& a, const CUtensorMap& w, const float* bias,
                        const float* residual, int ldr, void* C, int ldc, int out_prec, int M, int N,
                        int K, int epi, cudaStream_t st) {
#define PCD_GEMM_CASE(E)                                                                          \
  case E:                                                                                         \
    return out_prec == PCD_BF16 ? launch_gemm<BN, E, true>(a, w, bias, residual, ldr, C, ldc, M, N, K, st) \
                                : launch_gemm<BN, E, false>(a, w, bias, residual, ldr, C, ldc, M, N, K, st);
  switch (epi) {
    PCD_GEMM_CASE(PCD_EPI_BIAS)
    PCD_GEMM_CASE(PCD_EPI_BIAS_GELU)
    PCD_GEMM_CASE(PCD_EPI_BIAS_RESIDUAL)
  }
#undef PCD_GEMM_CASE
  set_error("gemm_bf16: unknown epilogue %d", epi);
  return PCD_ERR_INVALID;
}

}  // namespace pcd

using namespace pcd;

extern "C" int pcd_gemm_bf16(const uint16_t* A, int lda, const uint16_t* W, int ldw, const float* bias,
                             const float* residual, int ldr, void* C, int ldc, int out_precision,
                             int M, int N, int K, int epilogue, void* stream) {
  PCD_CHECK_ARG(M > 0 && N > 0 && K > 0, "gemm_bf16: empty problem");
  PCD_CHECK_ARG(K % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0, "gemm_bf16: K, lda, ldw must be multiples of 8 (K=%d lda=%d ldw=%d)", K, lda, ldw);
  PCD_CHECK_ARG((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0, "gemm_bf16: operands must be 16-byte aligned");
  PCD_CHECK_ARG(ldc % 8 == 0 && (reinterpret_cast<uintptr_t>(C) & 15) == 0, "gemm_bf16: C must be 16-byte aligned with ldc %% 8 == 0");
  PCD_CHECK_ARG(epilogue != PCD_EPI_BIAS_RESIDUAL || (residual != nullptr && ldr % 4 == 0), "gemm_bf16: residual missing or misaligned");
  const int BN = (N % 256 == 0 || N > 1024) ? 256 : (N > 64 ? 128 : 64);
  CUtensorMap tmA, tmW;
  uint64_t dimsA[2] = {(uint64_t)K, (uint64_t)M}, strA[1] = {(uint64_t)lda * 2};
  uint32_t boxA[2] = {G_BK, G_BM};
  uint64_t dimsW[2] = {(uint64_t)K, (uint64_t)N}, strW[1] = {(uint64_t)ldw * 2};
  uint32_t boxW[2] = {G_BK, (uint32_t)BN};
  int rc = encode_tmap_bf16(&tmA, A, 2, dimsA, strA, boxA);
  if (rc != PCD_OK) return rc;
  rc = encode_tmap_bf16(&tmW, W, 2, dimsW, strW, boxW);
  if (rc != PCD_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (BN == 256) return dispatch_epi<256>(tmA, tmW, bias, residual, ldr, C, ldc, out_precision, M, N, K, epilogue, st);
  if (BN == 128) return dispatch_epi<128>(tmA, tmW, bias, residual, ldr, C, ldc, out_precision, M, N, K, epilogue, st);
  return dispatch_epi<64>(tmA, tmW, bias, residual, ldr, C, ldc, out_precision, M, N, K, epilogue, st);
}
