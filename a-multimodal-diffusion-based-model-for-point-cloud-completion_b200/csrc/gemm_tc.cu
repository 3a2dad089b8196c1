// bf16 GEMM on the 5th-generation tensor cores:  C[M,N] = epi(A[M,K] W[N,K]^T + bias)
//
// Persistent, warp-specialised kernel (one CTA per SM, 320 threads):
//   warp 0   : TMA producer -- A (128 x 64) and W (BN x 64) bf16 tiles, 128B swizzle,
//              STAGES-deep mbarrier ring
//   warp 1   : MMA issuer   -- one elected lane issues tcgen05.mma (M=128, N=BN, K=16),
//              fp32 accumulators double-buffered in TMEM (2 x BN columns)
//   warps 2-9: epilogue     -- TMEM lane quarter = warp % 4, column half = (warp-2) / 4.
//              Double-buffered tcgen05.ld 32x32b.x32 (thread = output row), bias from a
//              per-tile shared-memory slice, GELU / fp32 residual, then the 32-row x 128-byte
//              chunk is staged in 128B-swizzled shared memory and written with ONE TMA store
//              (full 128-byte lines; M/N tails are clipped by the tensor map).  A residual
//              chunk, when requested, is TMA-loaded into the same buffer and updated in place.
// Measured with tools/gemm_probe.py: the MMA loop alone sustains ~1.7 PFLOP/s and the
// TMA-fed main loop ~1.4 PFLOP/s; register->global stores from the row-per-thread TMEM layout
// (32 distinct 128-byte lines per warp instruction) were the bottleneck of the first version,
// hence the staged TMA-store epilogue.
// Both operands are K-contiguous (activations [M,K], nn.Linear weights [N,K]) so no
// transposes are needed.  Tails in M, N, K are handled by TMA zero fill / clipping.
#include "common.cuh"
#include "tc_sm100.cuh"

namespace pcd {

using namespace tc;

constexpr int G_BM = 128, G_BK = 64;
constexpr int G_THREADS = 320, G_EPI_WARPS = 8;
constexpr int G_STAGE_TILE = 32 * 128;  // epilogue staging tile: 32 rows x 128 bytes

// Timeline of one epilogue warp and of the MMA issuer of CTA 0 (tools build only: -DPCD_GEMM_TRACE, tools/gemm_trace.py)
#ifdef PCD_GEMM_TRACE
constexpr int GT_TILES = 24, GT_PTS = 24;
__device__ unsigned long long g_gemm_trace[2 * GT_TILES * GT_PTS];
#define PCD_GTRACE(who, pt)                                                                        \
  do {                                                                                             \
    if (blockIdx.x == 0 && lane == 0 && it < GT_TILES) g_gemm_trace[((who) * GT_TILES + it) * GT_PTS + (pt)] = clock64(); \
  } while (0)
#else
#define PCD_GTRACE(who, pt) do { } while (0)
#endif


template <int BN, int EPI>
struct GemmCfg {
  static constexpr int NBUF = 1;  // staging tiles per epilogue warp
  static constexpr int A_BYTES = G_BM * G_BK * 2;   // 16 KB
  static constexpr int B_BYTES = BN * G_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int OUT_BYTES = G_EPI_WARPS * NBUF * G_STAGE_TILE;
  static constexpr int MISC_BYTES = 512 + 2 * BN * 4;  // barriers + bias slices
  static constexpr int BUDGET = 232448 - OUT_BYTES - MISC_BYTES;
  static constexpr int STAGES = (BUDGET / STAGE_BYTES) > 6 ? 6 : (BUDGET / STAGE_BYTES);
  static constexpr int TMEM_COLS = 2 * BN;          // 512 / 256: powers of two
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + OUT_BYTES + MISC_BYTES;
  static_assert(STAGES >= 3, "pipeline too shallow");
};

template <int BN, int EPI, bool OUT_BF16>
__global__ void __launch_bounds__(G_THREADS, 1)
gemm_bf16_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                    const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR,
                    const float* __restrict__ bias, int M, int N, int K, int dbg) {
  using Cfg = GemmCfg<BN, EPI>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int NBUF = Cfg::NBUF;
  // 1024-byte alignment is required by the 128B swizzle atom (8 rows x 128 B); the dynamic
  // shared window starts at offset 0 of the CTA (no static __shared__ in this kernel)
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  unsigned char* tiles = smem;
  unsigned char* stage_out = smem + STAGES * Cfg::STAGE_BYTES;             // [8 warps][NBUF][4 KB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(stage_out + Cfg::OUT_BYTES);
  uint64_t* full = bars;                    // [STAGES]  TMA -> MMA
  uint64_t* empty = bars + STAGES;          // [STAGES]  MMA -> TMA
  uint64_t* acc_full = bars + 2 * STAGES;   // [2]       MMA -> epilogue
  uint64_t* acc_empty = acc_full + 2;       // [2]       epilogue -> MMA
  uint64_t* res_bar = acc_empty + 2;        // [8 warps][2] residual chunk landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + 2 * G_EPI_WARPS);
  float* sbias = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(bars) + 512);  // [2][BN]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_m = (M + G_BM - 1) / G_BM, num_n = (N + BN - 1) / BN;
  const int num_tiles = num_m * num_n;
  const int num_kb = (K + G_BK - 1) / G_BK;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmA);
    prefetch_tensormap(&tmW);
    prefetch_tensormap(&tmC);
    if (EPI == PCD_EPI_BIAS_RESIDUAL) prefetch_tensormap(&tmR);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], G_EPI_WARPS);  // one arrival per epilogue warp
    }
    for (int s = 0; s < 2 * G_EPI_WARPS; ++s) mbar_init(&res_bar[s], 1);
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
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        int m_blk = t / num_n, n_blk = t % num_n;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          unsigned char* sa = tiles + stage * Cfg::STAGE_BYTES;
          unsigned char* sb = sa + Cfg::A_BYTES;
          if (dbg & 2) {  // profiling aid: no TMA traffic, the MMAs run on stale shared memory
            mbar_arrive(&full[stage]);
          } else {
            mbar_expect_tx(&full[stage], Cfg::STAGE_BYTES);
            tma_load_2d(sa, &tmA, &full[stage], kb * G_BK, m_blk * G_BM);
            tma_load_2d(sb, &tmW, &full[stage], kb * G_BK, n_blk * BN);
          }
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
        if (elect_one()) {
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
    constexpr int NCH = HALF_N / 32;   // 32-column TMEM chunks per thread
    // one staged store covers 128 bytes per row: 64 bf16 columns (2 chunks) or 32 fp32 columns
    constexpr int CH_PER_STORE = OUT_BF16 ? 2 : 1;
    static_assert(BN % 128 == 0, "BN must be a multiple of 128");
    static_assert(!(EPI == PCD_EPI_BIAS_RESIDUAL && OUT_BF16), "the residual stream is fp32");
    const int etid = threadIdx.x - 64;  // 0..255
    unsigned char* my_buf = stage_out + ew * NBUF * G_STAGE_TILE;
    uint64_t* my_res_bar = res_bar + 2 * ew;
    const int sw = lane & 7;            // 128B swizzle: 16-byte piece index ^= (row & 7)
    unsigned char* my_row0 = my_buf + lane * 128;
    uint32_t res_uses[2] = {0, 0};      // per-buffer residual-load counters (mbarrier parity)
    int sidx = 0;                       // running store-group index of this warp
    int it = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int m_blk = t / num_n, n_blk = t % num_n;
      const int row0 = m_blk * G_BM + quarter * 32;       // first row of this warp
      const int colbase = n_blk * BN + half * HALF_N;     // first column of this warp
      // stage this tile's bias slice in shared memory (double-buffered by accumulator stage)
      float* sb = sbias + as * BN;
      if (etid < BN) {
        const int n = n_blk * BN + etid;
        sb[etid] = (bias != nullptr && n < N) ? __ldg(bias + n) : 0.f;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      mbar_wait(&acc_full[as], aphase);
      tcgen05_fence_after();
      if (dbg & 1) {  // profiling aid: drain nothing, release the accumulator immediately
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[as]);
        continue;
      }
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + as * BN + half * HALF_N;
      uint32_t rbuf[2][32];
      tmem_ld_32x32b_x32(taddr, rbuf[0]);
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        uint32_t* r = rbuf[c & 1];
        const int b = sidx % NBUF;
        unsigned char* buf_row = my_row0 + b * G_STAGE_TILE;
        const bool first_of_store = (c % CH_PER_STORE) == 0;
        const bool last_of_store = (c % CH_PER_STORE) == CH_PER_STORE - 1;
        if (first_of_store) {
          if (elect_one()) {
            tma_store_wait_read<0>();  // staging buffer free again
            if (EPI == PCD_EPI_BIAS_RESIDUAL) {
              // residual chunk -> staging buffer (updated in place below).  Not prefetched: the
              // hot path adds residuals in the fused add+LayerNorm kernel instead (elementwise.cu).
              mbar_expect_tx(&my_res_bar[0], G_STAGE_TILE);
              tma_load_2d(my_buf, &tmR, &my_res_bar[0], colbase + c * 32, row0);
            }
          }
          __syncwarp();
          if (EPI == PCD_EPI_BIAS_RESIDUAL) {
            mbar_wait(&my_res_bar[0], res_uses[0] & 1);
            res_uses[0]++;
          }
        }
        tmem_ld_wait();
        if (c + 1 < NCH) tmem_ld_32x32b_x32(taddr + (c + 1) * 32, rbuf[(c + 1) & 1]);
        const float* sbc = sb + half * HALF_N + c * 32;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 b4 = *reinterpret_cast<const float4*>(sbc + j);
          v[j] = __uint_as_float(r[j]) + b4.x;
          v[j + 1] = __uint_as_float(r[j + 1]) + b4.y;
          v[j + 2] = __uint_as_float(r[j + 2]) + b4.z;
          v[j + 3] = __uint_as_float(r[j + 3]) + b4.w;
        }
        if (EPI == PCD_EPI_BIAS_GELU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = gelu_fast(v[j]);
        }
        if (OUT_BF16) {
          // 32 columns = 64 bytes = pieces [4*(c&1), +4) of the 128-byte staging row
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int piece = ((c % CH_PER_STORE) * 4 + i) ^ sw;
            *reinterpret_cast<uint4*>(buf_row + piece * 16) =
                make_uint4(pack_bf16x2(v[8 * i], v[8 * i + 1]), pack_bf16x2(v[8 * i + 2], v[8 * i + 3]),
                           pack_bf16x2(v[8 * i + 4], v[8 * i + 5]), pack_bf16x2(v[8 * i + 6], v[8 * i + 7]));
          }
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float4* pp = reinterpret_cast<float4*>(buf_row + ((i ^ sw) * 16));
            float4 o = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
            if (EPI == PCD_EPI_BIAS_RESIDUAL) {
              const float4 rr = *pp;  // residual chunk landed here by TMA; update in place
              o.x += rr.x; o.y += rr.y; o.z += rr.z; o.w += rr.w;
            }
            *pp = o;
          }
        }
        if (ew == 0) PCD_GTRACE(0, 5 + 4 * c);  // chunk c: math + staging stores issued
        if (last_of_store) {
          fence_proxy_async_smem();
          __syncwarp();
          if (elect_one()) {
            const int col = colbase + (c / CH_PER_STORE) * (OUT_BF16 ? 64 : 32);
            tma_store_2d(&tmC, my_buf + b * G_STAGE_TILE, col, row0);
            tma_store_commit();
          }
          ++sidx;
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[as]);
    }
    if (elect_one()) tma_store_wait_all<0>();  // all output tiles are globally visible
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------
// CTA-pair kernel (cta_group::2): two CTAs of a cluster compute one 256 x 256 tile.  Each CTA
// stages its own 128 rows of A and HALF of the W tile (128 of the 256 N rows); the leader's MMA
// thread issues M=256 instructions that read both halves from both shared memories, so every CTA
// writes and reads 32 KB per k-block instead of 48 KB.  tools/gemm_probe.py shows the single-CTA
// kernel bound by shared-memory bandwidth (TMA writes + UMMA operand reads + epilogue staging =
// ~1 MB per 128x256 tile at 128 B/clk), which is what this layout relieves.
// ---------------------------------------------------------------------------
// DEEPK (residual epilogues, K >= 1024): the main loop is the critical path and wants pipeline stages more than
// the epilogue wants a second staging tile -- one staging tile per warp, residual chunks fetched on demand.
template <int EPI, bool DEEPK = false>
struct Gemm2Cfg {
  static constexpr int BN = 256;                     // N of the pair tile (each CTA stages 128 rows of W)
  static constexpr int A_BYTES = G_BM * G_BK * 2;    // 16 KB
  static constexpr int B_BYTES = (BN / 2) * G_BK * 2;  // 16 KB
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  // two staging tiles per epilogue warp: a TMA store only has to have finished READING its tile
  // by the time the warp comes back to it two store groups later (the round trip is ~2k cycles)
  static constexpr int NBUF = DEEPK ? 1 : 2;
  // the residual + row-statistics epilogue writes TWO tensors (fp32 residual stream and its bf16
  // copy): a second set of staging tiles
  static constexpr int OUT1_BYTES = G_EPI_WARPS * NBUF * G_STAGE_TILE;
  static constexpr int OUT_BYTES = OUT1_BYTES + (EPI == PCD_EPI_RESIDUAL_STATS ? G_EPI_WARPS * G_STAGE_TILE : 0);
  static constexpr int MISC_BYTES = 512 + 4 * BN * 4;  // barriers + per-tile bias and column-sum slices (two tiles each)
  static constexpr int BUDGET = 232448 - OUT_BYTES - MISC_BYTES;
  static constexpr int STAGES = (BUDGET / STAGE_BYTES) > 6 ? 6 : (BUDGET / STAGE_BYTES);
  static constexpr int TMEM_COLS = 2 * BN;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + OUT_BYTES + MISC_BYTES;
};

template <int EPI, bool OUT_BF16, bool DEEPK = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(G_THREADS, 1)
gemm_bf16_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                     const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR,
                     const __grid_constant__ CUtensorMap tmC2, const float* __restrict__ bias,
                     const float* __restrict__ colsum, const float2* __restrict__ stats_in, int stats_in_slots,
                     float2* __restrict__ stats_out, float ln_eps, int M, int N, int K, int dbg) {
  using Cfg = Gemm2Cfg<EPI, DEEPK>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int BN = Cfg::BN;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  unsigned char* tiles = smem;
  unsigned char* stage_out = smem + STAGES * Cfg::STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(stage_out + Cfg::OUT_BYTES);
  uint64_t* full = bars;                    // [STAGES]  both producers -> leader MMA (leader copy used)
  uint64_t* empty = bars + STAGES;          // [STAGES]  leader MMA -> each CTA's producer (multicast)
  uint64_t* acc_full = bars + 2 * STAGES;   // [2]       leader MMA -> each CTA's epilogue (multicast)
  uint64_t* acc_empty = acc_full + 2;       // [2]       both epilogues -> leader MMA (leader copy used)
  uint64_t* res_bar = acc_empty + 2;        // [8 warps][2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + 2 * G_EPI_WARPS);
  float* sbias = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(bars) + 512);  // [2][BN]
  float* scolsum = sbias + 2 * BN;                                                         // [2][BN] (LN-folded epilogues)
  constexpr bool kResid = (EPI == PCD_EPI_BIAS_RESIDUAL || EPI == PCD_EPI_RESIDUAL_STATS);
  constexpr bool kStats = (EPI == PCD_EPI_RESIDUAL_STATS);
  constexpr bool kLnFold = (EPI == PCD_EPI_LN_BIAS || EPI == PCD_EPI_LN_BIAS_GELU);
  constexpr bool kGelu = (EPI == PCD_EPI_BIAS_GELU || EPI == PCD_EPI_LN_BIAS_GELU);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();       // 0 = leader
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const int num_m = (M + 2 * G_BM - 1) / (2 * G_BM), num_n = (N + BN - 1) / BN;
  const int num_tiles = num_m * num_n;
  const int num_kb = (K + G_BK - 1) / G_BK;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmA);
    prefetch_tensormap(&tmW);
    prefetch_tensormap(&tmC);
    if (kResid) prefetch_tensormap(&tmR);
    if (kStats) prefetch_tensormap(&tmC2);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], 2 * G_EPI_WARPS);  // one arrival per epilogue warp of BOTH CTAs
    }
    for (int s = 0; s < 2 * G_EPI_WARPS; ++s) mbar_init(&res_bar[s], 1);
    fence_barrier_init();
  }
  cluster_sync_all();  // barriers of both CTAs initialised before any remote signal
  if (warp == 1) {
    tmem_alloc_2sm(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish_2sm();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------- TMA producer (both CTAs) -------------------------
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = pair; t < num_tiles; t += num_pairs) {
        const int m_blk = t / num_n, n_blk = t % num_n;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          unsigned char* sa = tiles + stage * Cfg::STAGE_BYTES;
          unsigned char* sb = sa + Cfg::A_BYTES;
          const uint32_t leader_full = mapa_shared(smem_u32(&full[stage]), 0);
          if (rank == 0) mbar_expect_tx(&full[stage], 2 * Cfg::STAGE_BYTES);  // bytes of both CTAs
          tma_load_2d_2sm(sa, &tmA, leader_full, kb * G_BK, m_blk * 2 * G_BM + (int)rank * G_BM);
          tma_load_2d_2sm(sb, &tmW, leader_full, kb * G_BK, n_blk * BN + (int)rank * (BN / 2));
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // -------------------------- MMA issuer (leader CTA) --------------------------
    if (rank == 0) {
      constexpr uint32_t idesc = idesc_bf16_f32(2 * G_BM, BN, 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int t = pair; t < num_tiles; t += num_pairs, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        PCD_GTRACE(1, 0);  // issuer: tile begins
        mbar_wait(&acc_empty[as], aphase ^ 1);
        PCD_GTRACE(1, 1);  // issuer: accumulator buffer free
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full[stage], phase);
          if (kb == 0) PCD_GTRACE(1, 2);  // issuer: first operands landed
          if (kb == num_kb - 1) PCD_GTRACE(1, 3);  // issuer: last operands landed
          tcgen05_fence_after();
          if (elect_one()) {
            const uint32_t sa = smem_u32(tiles + stage * Cfg::STAGE_BYTES);
            const uint64_t adesc = smem_desc_sw128(sa);
            const uint64_t bdesc = smem_desc_sw128(sa + Cfg::A_BYTES);
            if (!(dbg & 8)) {  // profiling aid: bit 3 skips the MMAs (epilogue timed without tensor traffic)
#pragma unroll
              for (int k = 0; k < G_BK / 16; ++k)
                umma_bf16_ss_2sm(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
            }
            umma_commit_2sm(&empty[stage], 3);                       // frees the slot in both CTAs
            if (kb == num_kb - 1) umma_commit_2sm(&acc_full[as], 3); // accumulators ready in both CTAs
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // --------------------------- epilogue (both CTAs) ----------------------------
    const int ew = warp - 2;
    const int quarter = warp & 3;
    const int half = ew >> 2;
    constexpr int HALF_N = BN / 2;
    constexpr int NCH = HALF_N / 32;
    constexpr int CH_PER_STORE = OUT_BF16 ? 2 : 1;
    static_assert(!(kResid && OUT_BF16), "the residual stream is fp32");
    static_assert(!kLnFold || OUT_BF16, "the LayerNorm-folded projections write bf16");
    const int etid = threadIdx.x - 64;
    constexpr int NBUF = Cfg::NBUF;
    unsigned char* my_buf0 = stage_out + ew * NBUF * G_STAGE_TILE;
    unsigned char* my_bbuf = stage_out + Cfg::OUT1_BYTES + ew * G_STAGE_TILE;  // bf16 copy (kStats), one tile
    int sidx = 0;   // running store-group index of this warp
    uint64_t* my_res_bar = res_bar + 2 * ew;  // [2]: residual chunk landed in staging tile 0 / 1
    const int sw = lane & 7;
    uint32_t res_ph0 = 0, res_ph1 = 0;
    constexpr bool kPrefetch = kResid && NBUF == 2;  // residual chunks requested ahead (staging tile = chunk & 1)
    static_assert(!kPrefetch || NCH % 2 == 0, "residual prefetch assumes staging tile = chunk & 1");
    // LayerNorm statistics (LN-folded projections): the producer wrote one (mean, M2) pair per 128
    // columns of the normalised vector; they are combined with Chan's parallel-variance formula
    constexpr int LN_MAXS = 8;
    float2 ln_raw[LN_MAXS];
    float ln_mu = 0.f, ln_rstd = 0.f;
    auto ln_issue = [&](int row) {
#pragma unroll
      for (int s2 = 0; s2 < LN_MAXS; ++s2)
        ln_raw[s2] = (s2 < stats_in_slots && row < M) ? __ldg(stats_in + (size_t)row * stats_in_slots + s2) : make_float2(0.f, 0.f);
    };
    auto ln_combine = [&]() {
      float msum = 0.f;
#pragma unroll
      for (int s2 = 0; s2 < LN_MAXS; ++s2) msum += ln_raw[s2].x;  // (absent slots hold zeros)
      ln_mu = msum / (float)stats_in_slots;
      float m2 = 0.f;
#pragma unroll
      for (int s2 = 0; s2 < LN_MAXS; ++s2) {
        if (s2 < stats_in_slots) {
          const float dm = ln_raw[s2].x - ln_mu;
          m2 += ln_raw[s2].y + 128.f * dm * dm;
        }
      }
      ln_rstd = rsqrtf(m2 / (128.f * (float)stats_in_slots) + ln_eps);
    };
    if (kLnFold && stats_in_slots <= LN_MAXS && pair < num_tiles) {
      ln_issue((pair / num_n) * 2 * G_BM + (int)rank * G_BM + quarter * 32 + lane);
      ln_combine();
    }
    int it = 0;
    for (int t = pair; t < num_tiles; t += num_pairs, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int m_blk = t / num_n, n_blk = t % num_n;
      const int row0 = m_blk * 2 * G_BM + (int)rank * G_BM + quarter * 32;
      const int colbase = n_blk * BN + half * HALF_N;
      float* sb = sbias + as * BN;
      float* scs = scolsum + as * BN;
      if (etid < BN) {
        const int n = n_blk * BN + etid;
        sb[etid] = (bias != nullptr && n < N) ? __ldg(bias + n) : 0.f;
        // the column sums of the folded weights go through shared memory like the bias: with the whole carve-out
        // given to shared memory there is no L1 to speak of, and eight uniform global loads per 32-column chunk
        // were L2 round trips inside the epilogue's critical path (tools/gemm_trace.py)
        if (kLnFold) scs[etid] = (n < N) ? __ldg(colsum + n) : 0.f;
      }
      // LayerNorm statistics of this thread's row (LN-folded projections): normally prefetched during
      // the previous tile (below); wide rows (> 8 slots) are loaded here
      if (kLnFold && stats_in_slots > LN_MAXS) {
        const int row = row0 + lane;
        ln_mu = 0.f;
        ln_rstd = 0.f;
        if (row < M) {
          const float2* sp = stats_in + (size_t)row * stats_in_slots;
          float msum = 0.f;
          for (int s2 = 0; s2 < stats_in_slots; ++s2) msum += sp[s2].x;
          ln_mu = msum / (float)stats_in_slots;
          float m2 = 0.f;
          for (int s2 = 0; s2 < stats_in_slots; ++s2) {
            const float2 v2 = sp[s2];
            const float dm = v2.x - ln_mu;
            m2 += v2.y + 128.f * dm * dm;
          }
          ln_rstd = rsqrtf(m2 / (128.f * (float)stats_in_slots) + ln_eps);
        }
      }
      const float cur_mu = ln_mu, cur_rstd = ln_rstd;
      if (kLnFold && stats_in_slots <= LN_MAXS && t + num_pairs < num_tiles) {
        // next tile's statistics: loads in flight while this tile is written out (the epilogue is the
        // critical path of these GEMMs, so a load-use stall at tile start would cost ~1 us per tile)
        const int tn = t + num_pairs;
        ln_issue((tn / num_n) * 2 * G_BM + (int)rank * G_BM + quarter * 32 + lane);
      }
      if (kPrefetch && !(dbg & 1)) {
        // residual chunks 0 and 1 of this tile -> the two staging tiles, while the accumulator is still
        // being computed (the stores of the previous tile have long finished reading them)
        if (elect_one()) {
          tma_store_wait_read<0>();
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            mbar_expect_tx(&my_res_bar[c], G_STAGE_TILE);
            tma_load_2d(my_buf0 + c * G_STAGE_TILE, &tmR, &my_res_bar[c], colbase + c * 32, row0);
          }
        }
        __syncwarp();
      }
      if (ew == 0) PCD_GTRACE(0, 0);  // epilogue: tile begins (bias slice written, residual prefetch issued)
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (ew == 0) PCD_GTRACE(0, 1);  // epilogue: past the per-tile CTA barrier
      mbar_wait(&acc_full[as], aphase);
      if (ew == 0) PCD_GTRACE(0, 2);  // epilogue: accumulators complete
      tcgen05_fence_after();
      const uint32_t leader_acc_empty = mapa_shared(smem_u32(&acc_empty[as]), 0);
      if (dbg & 1) {
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(leader_acc_empty);
        continue;
      }
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + as * BN + half * HALF_N;
      uint32_t rbuf[2][32];
      tmem_ld_32x32b_x32(taddr, rbuf[0]);
      // running (count, mean, M2) of the bf16-rounded row over this thread's 128 columns (kStats)
      float st_mean = 0.f, st_m2 = 0.f;
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        uint32_t* r = rbuf[c & 1];
        const bool first_of_store = (c % CH_PER_STORE) == 0;
        const bool last_of_store = (c % CH_PER_STORE) == CH_PER_STORE - 1;
        unsigned char* my_buf = my_buf0 + (sidx % NBUF) * G_STAGE_TILE;
        unsigned char* buf_row = my_buf + lane * 128;
        if (kResid && !kPrefetch) {
          // one staging tile: wait until its previous stores have read it, then fetch this chunk's residual
          if (elect_one()) {
            tma_store_wait_read<0>();
            mbar_expect_tx(&my_res_bar[0], G_STAGE_TILE);
            tma_load_2d(my_buf, &tmR, &my_res_bar[0], colbase + c * 32, row0);
          }
          __syncwarp();
          mbar_wait(&my_res_bar[0], res_ph0);
          res_ph0 ^= 1;
        } else if (kResid) {
          // this chunk's residual was requested one chunk (or one tile) ahead
          if (c & 1) {
            mbar_wait(&my_res_bar[1], res_ph1);
            res_ph1 ^= 1;
          } else {
            mbar_wait(&my_res_bar[0], res_ph0);
            res_ph0 ^= 1;
          }
        } else if (first_of_store) {
          if (elect_one()) tma_store_wait_read<NBUF - 1>();  // the store that used this tile two groups ago has read it
          __syncwarp();
        }
        if (ew == 0) PCD_GTRACE(0, 3 + 4 * c);  // chunk c: staging tile free / residual landed
        tmem_ld_wait();
        if (ew == 0) PCD_GTRACE(0, 4 + 4 * c);  // chunk c: accumulators in registers
        if (c + 1 < NCH) {
          tmem_ld_32x32b_x32(taddr + (c + 1) * 32, rbuf[(c + 1) & 1]);
        } else {
          // the last chunk of this warp's accumulator slice is in registers: hand the TMEM buffer back to the MMA
          // issuer now, a quarter of an epilogue earlier than after the math and stores of this chunk (the MMA
          // stream and the epilogue are nearly balanced at K = 512, so slack on the hand-off is what absorbs jitter)
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(leader_acc_empty);
        }
        const float* sbc = sb + half * HALF_N + c * 32;
        float v[32];
        {
          // bias / LayerNorm fold / GELU on pairs (fma.rn.f32x2: half the FMA-pipe issue slots -- the
          // epilogue warps, two per scheduler, are issue-latency bound).  LayerNorm folded into the
          // projection: with W' = gamma o W (bf16), s_n = sum_k W'[n,k], c_n = beta . W[n,:] + b_n:
          //   LN(x) W^T + b = rstd (x W'^T - mu s) + c
          const float* scc = scs + half * HALF_N + c * 32;  // warp-uniform: broadcast loads from shared memory
          const uint64_t nmu2 = pack2(-cur_mu, -cur_mu), rstd2 = pack2(cur_rstd, cur_rstd);
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 b4 = *reinterpret_cast<const float4*>(sbc + j);
            uint64_t a01 = pack2(__uint_as_float(r[j]), __uint_as_float(r[j + 1]));
            uint64_t a23 = pack2(__uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
            if (kLnFold) {
              const float4 s4 = *reinterpret_cast<const float4*>(scc + j);
              a01 = fma2(rstd2, fma2(nmu2, pack2(s4.x, s4.y), a01), pack2(b4.x, b4.y));
              a23 = fma2(rstd2, fma2(nmu2, pack2(s4.z, s4.w), a23), pack2(b4.z, b4.w));
            } else {
              a01 = add2(a01, pack2(b4.x, b4.y));
              a23 = add2(a23, pack2(b4.z, b4.w));
            }
            if (kGelu) {
              a01 = gelu_fast2(a01);
              a23 = gelu_fast2(a23);
            }
            unpack2(a01, v[j], v[j + 1]);
            unpack2(a23, v[j + 2], v[j + 3]);
          }
        }
        if (OUT_BF16) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int piece = ((c % CH_PER_STORE) * 4 + i) ^ sw;
            *reinterpret_cast<uint4*>(buf_row + piece * 16) =
                make_uint4(pack_bf16x2(v[8 * i], v[8 * i + 1]), pack_bf16x2(v[8 * i + 2], v[8 * i + 3]),
                           pack_bf16x2(v[8 * i + 4], v[8 * i + 5]), pack_bf16x2(v[8 * i + 6], v[8 * i + 7]));
          }
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float4* pp = reinterpret_cast<float4*>(buf_row + ((i ^ sw) * 16));
            float4 o = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
            if (kResid) {
              const float4 rr = *pp;
              o.x += rr.x; o.y += rr.y; o.z += rr.z; o.w += rr.w;
              if (kStats) {
                v[4 * i] = o.x; v[4 * i + 1] = o.y; v[4 * i + 2] = o.z; v[4 * i + 3] = o.w;
              }
            }
            *pp = o;
          }
        }
        if (kStats) {
          // bf16 copy of the updated residual rows (the A operand of the next, LN-folded projection)
          // and the row statistics of exactly those rounded values
          unsigned char* bbuf_row = my_bbuf + lane * 128;
          float q[32];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint32_t w4[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const uint32_t pkd = pack_bf16x2(v[8 * i + 2 * u], v[8 * i + 2 * u + 1]);
              w4[u] = pkd;
              q[8 * i + 2 * u] = __uint_as_float(pkd << 16);
              q[8 * i + 2 * u + 1] = __uint_as_float(pkd & 0xffff0000u);
            }
            const int piece = ((c & 1) * 4 + i) ^ sw;
            *reinterpret_cast<uint4*>(bbuf_row + piece * 16) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
          }
          float csum = 0.f;
#pragma unroll
          for (int j = 0; j < 32; ++j) csum += q[j];
          const float cmean = csum * (1.f / 32.f);
          float cm2 = 0.f;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float dq = q[j] - cmean;
            cm2 = fmaf(dq, dq, cm2);
          }
          // merge (32 values) into the running statistics of 32*c values
          const float na = 32.f * c, nb = 32.f;
          const float delta = cmean - st_mean;
          st_mean += delta * (nb / (na + nb));
          st_m2 += cm2 + delta * delta * (na * nb / (na + nb));
        }
        if (ew == 0) PCD_GTRACE(0, 5 + 4 * c);  // chunk c: math + staging stores issued
        if (last_of_store) {
          fence_proxy_async_smem();
          __syncwarp();
          if (elect_one()) {
            const int col = colbase + (c / CH_PER_STORE) * (OUT_BF16 ? 64 : 32);
            tma_store_2d(&tmC, my_buf, col, row0);
            tma_store_commit();
            if (kStats && (c & 1)) {
              tma_store_2d(&tmC2, my_bbuf, colbase + (c >> 1) * 64, row0);
              tma_store_commit();
            }
            if (kPrefetch) {
              // the stores just issued must have read their tiles before (a) the residual of chunk
              // c + 2 lands in this staging tile and (b) the next chunk writes the bf16 tile
              tma_store_wait_read<0>();
              if (c + 2 < NCH) {
                mbar_expect_tx(&my_res_bar[c & 1], G_STAGE_TILE);
                tma_load_2d(my_buf, &tmR, &my_res_bar[c & 1], colbase + (c + 2) * 32, row0);
              }
            }
          }
          if (kPrefetch) __syncwarp();
          ++sidx;
        }
        if (ew == 0) PCD_GTRACE(0, 6 + 4 * c);  // chunk c: TMA store(s) issued
      }
      if (kLnFold && stats_in_slots <= LN_MAXS && t + num_pairs < num_tiles) ln_combine();
      if (kStats) {
        const int row = row0 + lane;
        if (row < M) stats_out[(size_t)row * (N / HALF_N) + n_blk * 2 + half] = make_float2(st_mean, st_m2);
      }
    }
    if (elect_one()) tma_store_wait_all<0>();
  }

  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();  // neither CTA may release TMEM / exit while its peer can still signal it
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc_2sm(tmem_base, Cfg::TMEM_COLS);
  }
}

#ifdef PCD_GEMM_TRACE
extern "C" __attribute__((visibility("default"))) int pcd_gemm_trace_read(unsigned long long* dst, int n) {
  const int total = 2 * GT_TILES * GT_PTS;
  return cudaMemcpyFromSymbol(dst, g_gemm_trace, sizeof(unsigned long long) * (n < total ? n : total)) == cudaSuccess ? total : -1;
}
#endif

struct LnArgs {  // extra operands of the residual+statistics and LayerNorm-folded epilogues
  CUtensorMap tmC2;
  const float* colsum;
  const float2* stats_in;
  int stats_in_slots;
  float2* stats_out;
  float ln_eps;
};

template <int EPI, bool OUT_BF16, bool DEEPK = false>
static int launch_gemm2(const CUtensorMap& tmA, const CUtensorMap& tmW, const CUtensorMap& tmC,
                        const CUtensorMap& tmR, const float* bias, int M, int N, int K, cudaStream_t st,
                        const LnArgs* ln, int dbg) {
  using Cfg = Gemm2Cfg<EPI, DEEPK>;
  auto kern = gemm_bf16_tc2_kernel<EPI, OUT_BF16, DEEPK>;
  {  // per device, not per process: set on every launch (a host-side table lookup)
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) {
      set_error("gemm_bf16(pair): cudaFuncSetAttribute(%d): %s", Cfg::SMEM_BYTES, cudaGetErrorString(e));
      return PCD_ERR_CUDA;
    }
  }
  const int tiles = ceil_div(M, 2 * G_BM) * ceil_div(N, Cfg::BN);
  int pairs = num_sms() / 2;
  if (tiles < pairs) pairs = tiles;
  static const LnArgs none = {};
  const LnArgs& x = ln ? *ln : none;
  kern<<<2 * pairs, G_THREADS, Cfg::SMEM_BYTES, st>>>(tmA, tmW, tmC, tmR, ln ? x.tmC2 : tmC, bias, x.colsum, x.stats_in,
                                                       x.stats_in_slots, x.stats_out, x.ln_eps, M, N, K, dbg);
  PCD_CHECK_LAUNCH("gemm_bf16(pair)");
  return PCD_OK;
}

static int dispatch_epi2(const CUtensorMap& a, const CUtensorMap& w, const CUtensorMap& c, const CUtensorMap& r,
                         const float* bias, int out_prec, int M, int N, int K, int epi, cudaStream_t st,
                         const LnArgs* ln, int dbg) {
  const bool ob = out_prec == PCD_BF16;
  switch (epi) {
    case PCD_EPI_RESIDUAL_STATS:
      if (ob || ln == nullptr) break;
      if (K >= 1024) return launch_gemm2<PCD_EPI_RESIDUAL_STATS, false, true>(a, w, c, r, bias, M, N, K, st, ln, dbg);
      return launch_gemm2<PCD_EPI_RESIDUAL_STATS, false>(a, w, c, r, bias, M, N, K, st, ln, dbg);
    case PCD_EPI_LN_BIAS:
      if (!ob || ln == nullptr) break;
      return launch_gemm2<PCD_EPI_LN_BIAS, true>(a, w, c, r, bias, M, N, K, st, ln, dbg);
    case PCD_EPI_LN_BIAS_GELU:
      if (!ob || ln == nullptr) break;
      return launch_gemm2<PCD_EPI_LN_BIAS_GELU, true>(a, w, c, r, bias, M, N, K, st, ln, dbg);
    case PCD_EPI_BIAS:
      return ob ? launch_gemm2<PCD_EPI_BIAS, true>(a, w, c, r, bias, M, N, K, st, nullptr, dbg)
                : launch_gemm2<PCD_EPI_BIAS, false>(a, w, c, r, bias, M, N, K, st, nullptr, dbg);
    case PCD_EPI_BIAS_GELU:
      return ob ? launch_gemm2<PCD_EPI_BIAS_GELU, true>(a, w, c, r, bias, M, N, K, st, nullptr, dbg)
                : launch_gemm2<PCD_EPI_BIAS_GELU, false>(a, w, c, r, bias, M, N, K, st, nullptr, dbg);
    case PCD_EPI_BIAS_RESIDUAL:
      if (ob) break;
      return launch_gemm2<PCD_EPI_BIAS_RESIDUAL, false>(a, w, c, r, bias, M, N, K, st, nullptr, dbg);
  }
  set_error("gemm_bf16: unsupported epilogue %d / output precision %d", epi, out_prec);
  return PCD_ERR_INVALID;
}

template <int BN, int EPI, bool OUT_BF16>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmW, const CUtensorMap& tmC,
                       const CUtensorMap& tmR, const float* bias, int M, int N, int K, cudaStream_t st, int dbg) {
  using Cfg = GemmCfg<BN, EPI>;
  auto kern = gemm_bf16_tc_kernel<BN, EPI, OUT_BF16>;
  {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) {
      set_error("gemm_bf16: cudaFuncSetAttribute(%d): %s", Cfg::SMEM_BYTES, cudaGetErrorString(e));
      return PCD_ERR_CUDA;
    }
  }
  int tiles = ceil_div(M, G_BM) * ceil_div(N, BN);
  int grid = tiles < num_sms() ? tiles : num_sms();
  kern<<<grid, G_THREADS, Cfg::SMEM_BYTES, st>>>(tmA, tmW, tmC, tmR, bias, M, N, K, dbg);
  PCD_CHECK_LAUNCH("gemm_bf16");
  return PCD_OK;
}

template <int BN>
static int dispatch_epi(const CUtensorMap& a, const CUtensorMap& w, const CUtensorMap& c, const CUtensorMap& r,
                        const float* bias, int out_prec, int M, int N, int K, int epi, cudaStream_t st, int dbg) {
  const bool ob = out_prec == PCD_BF16;
  switch (epi) {
    case PCD_EPI_BIAS:
      return ob ? launch_gemm<BN, PCD_EPI_BIAS, true>(a, w, c, r, bias, M, N, K, st, dbg)
                : launch_gemm<BN, PCD_EPI_BIAS, false>(a, w, c, r, bias, M, N, K, st, dbg);
    case PCD_EPI_BIAS_GELU:
      return ob ? launch_gemm<BN, PCD_EPI_BIAS_GELU, true>(a, w, c, r, bias, M, N, K, st, dbg)
                : launch_gemm<BN, PCD_EPI_BIAS_GELU, false>(a, w, c, r, bias, M, N, K, st, dbg);
    case PCD_EPI_BIAS_RESIDUAL:
      if (ob) break;
      return launch_gemm<BN, PCD_EPI_BIAS_RESIDUAL, false>(a, w, c, r, bias, M, N, K, st, dbg);
  }
  set_error("gemm_bf16: unsupported epilogue %d / output precision %d", epi, out_prec);
  return PCD_ERR_INVALID;
}

}  // namespace pcd

using namespace pcd;

static int gemm_bf16_impl(const pcd_gemm_args& g, void* stream) {
  const uint16_t* A = (const uint16_t*)g.A;
  const uint16_t* W = (const uint16_t*)g.W;
  const int M = g.M, N = g.N, K = g.K, lda = g.lda, ldw = g.ldw, ldc = g.ldc, ldr = g.ldr;
  const int epilogue = g.epilogue, out_precision = g.out_precision;
  void* C = g.C;
  const float* residual = g.residual;
  PCD_CHECK_ARG(M > 0 && N > 0 && K > 0, "gemm_bf16: empty problem");
  PCD_CHECK_ARG(K % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0, "gemm_bf16: K, lda, ldw must be multiples of 8 (K=%d lda=%d ldw=%d)", K, lda, ldw);
  PCD_CHECK_ARG((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0, "gemm_bf16: operands must be 16-byte aligned");
  const bool ob = out_precision == PCD_BF16;
  PCD_CHECK_ARG(out_precision == PCD_BF16 || out_precision == PCD_F32, "gemm_bf16: bad output precision");
  PCD_CHECK_ARG(ldc % (ob ? 8 : 4) == 0 && (reinterpret_cast<uintptr_t>(C) & 15) == 0, "gemm_bf16: C must be 16-byte aligned with a 16-byte row pitch (ldc=%d)", ldc);
  const bool resid = epilogue == PCD_EPI_BIAS_RESIDUAL || epilogue == PCD_EPI_RESIDUAL_STATS;
  const bool lnfold = epilogue == PCD_EPI_LN_BIAS || epilogue == PCD_EPI_LN_BIAS_GELU;
  PCD_CHECK_ARG(!resid || (residual != nullptr && ldr % 4 == 0 && (reinterpret_cast<uintptr_t>(residual) & 15) == 0 && !ob),
                "gemm_bf16: residual epilogue needs an fp32, 16-byte aligned residual and fp32 output");
  const int dbg = g.debug;
  const bool use_pair = (N % 256 == 0) && M >= 512 && !(dbg & 4);
  if (epilogue == PCD_EPI_RESIDUAL_STATS) {
    PCD_CHECK_ARG(use_pair, "gemm_bf16: the residual+statistics epilogue needs N %% 256 == 0 and M >= 512 (N=%d M=%d)", N, M);
    PCD_CHECK_ARG(g.C2 != nullptr && g.ldc2 % 8 == 0 && (reinterpret_cast<uintptr_t>(g.C2) & 15) == 0 && g.stats_out != nullptr,
                  "gemm_bf16: the residual+statistics epilogue needs a 16-byte aligned bf16 C2 and a stats_out buffer");
  }
  if (lnfold) {
    PCD_CHECK_ARG(use_pair && ob, "gemm_bf16: LayerNorm-folded epilogues need N %% 256 == 0, M >= 512 and bf16 output (N=%d M=%d)", N, M);
    PCD_CHECK_ARG(K % 128 == 0 && g.stats_in != nullptr && g.colsum != nullptr && g.bias != nullptr,
                  "gemm_bf16: LayerNorm-folded epilogues need K %% 128 == 0, stats_in, colsum and the folded bias");
  }
  const int BN = use_pair ? 128 /* W rows staged per CTA */ : ((N % 256 == 0 || N > 1024) ? 256 : 128);
  CUtensorMap tmA, tmW, tmC, tmR;
  uint64_t dimsA[2] = {(uint64_t)K, (uint64_t)M}, strA[1] = {(uint64_t)lda * 2};
  uint32_t boxA[2] = {G_BK, G_BM};
  uint64_t dimsW[2] = {(uint64_t)K, (uint64_t)N}, strW[1] = {(uint64_t)ldw * 2};
  uint32_t boxW[2] = {G_BK, (uint32_t)BN};
  uint64_t dimsC[2] = {(uint64_t)N, (uint64_t)M};
  int rc;
  if ((rc = encode_tmap_bf16(&tmA, A, 2, dimsA, strA, boxA)) != PCD_OK) return rc;
  if ((rc = encode_tmap_bf16(&tmW, W, 2, dimsW, strW, boxW)) != PCD_OK) return rc;
  if (ob) {
    uint64_t strC[1] = {(uint64_t)ldc * 2};
    uint32_t boxC[2] = {64, 32};
    if ((rc = encode_tmap_bf16(&tmC, C, 2, dimsC, strC, boxC)) != PCD_OK) return rc;
  } else {
    uint64_t strC[1] = {(uint64_t)ldc * 4};
    uint32_t boxC[2] = {32, 32};
    if ((rc = encode_tmap_f32(&tmC, C, 2, dimsC, strC, boxC)) != PCD_OK) return rc;
  }
  tmR = tmC;
  if (resid) {
    uint64_t strR[1] = {(uint64_t)ldr * 4};
    uint32_t boxR[2] = {32, 32};
    if ((rc = encode_tmap_f32(&tmR, residual, 2, dimsC, strR, boxR)) != PCD_OK) return rc;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (epilogue == PCD_EPI_RESIDUAL_STATS || lnfold) {
    LnArgs ln = {};
    ln.tmC2 = tmC;
    if (epilogue == PCD_EPI_RESIDUAL_STATS) {
      uint64_t strC2[1] = {(uint64_t)g.ldc2 * 2};
      uint32_t boxC2[2] = {64, 32};
      if ((rc = encode_tmap_bf16(&ln.tmC2, g.C2, 2, dimsC, strC2, boxC2)) != PCD_OK) return rc;
    }
    ln.colsum = g.colsum;
    ln.stats_in = reinterpret_cast<const float2*>(g.stats_in);
    ln.stats_in_slots = K / 128;
    ln.stats_out = reinterpret_cast<float2*>(g.stats_out);
    ln.ln_eps = g.ln_eps;
    return dispatch_epi2(tmA, tmW, tmC, tmR, g.bias, out_precision, M, N, K, epilogue, st, &ln, dbg);
  }
  if (use_pair) return dispatch_epi2(tmA, tmW, tmC, tmR, g.bias, out_precision, M, N, K, epilogue, st, nullptr, dbg);
  if (BN == 256) return dispatch_epi<256>(tmA, tmW, tmC, tmR, g.bias, out_precision, M, N, K, epilogue, st, dbg);
  return dispatch_epi<128>(tmA, tmW, tmC, tmR, g.bias, out_precision, M, N, K, epilogue, st, dbg);
}

extern "C" int pcd_gemm_bf16_ex(const pcd_gemm_args* args, void* stream) {
  PCD_CHECK_ARG(args != nullptr, "gemm_bf16_ex: null argument block");
  return gemm_bf16_impl(*args, stream);
}

extern "C" int pcd_gemm_bf16(const uint16_t* A, int lda, const uint16_t* W, int ldw, const float* bias,
                             const float* residual, int ldr, void* C, int ldc, int out_precision,
                             int M, int N, int K, int epilogue, void* stream) {
  PCD_CHECK_ARG(epilogue >= PCD_EPI_BIAS && epilogue <= PCD_EPI_BIAS_RESIDUAL,
                "gemm_bf16: epilogue %d needs pcd_gemm_bf16_ex", epilogue);
  pcd_gemm_args g = {};
  g.A = A; g.lda = lda; g.W = W; g.ldw = ldw; g.bias = bias; g.residual = residual; g.ldr = ldr;
  g.C = C; g.ldc = ldc; g.out_precision = out_precision; g.M = M; g.N = N; g.K = K; g.epilogue = epilogue;
  return gemm_bf16_impl(g, stream);
}
