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

extern int g_gemm_debug;

// ---------------------------------------------------------------------------
// CTA-pair kernel (cta_group::2): two CTAs of a cluster compute one 256 x 256 tile.  Each CTA
// stages its own 128 rows of A and HALF of the W tile (128 of the 256 N rows); the leader's MMA
// thread issues M=256 instructions that read both halves from both shared memories, so every CTA
// writes and reads 32 KB per k-block instead of 48 KB.  tools/gemm_probe.py shows the single-CTA
// kernel bound by shared-memory bandwidth (TMA writes + UMMA operand reads + epilogue staging =
// ~1 MB per 128x256 tile at 128 B/clk), which is what this layout relieves.
// ---------------------------------------------------------------------------
template <int EPI>
struct Gemm2Cfg {
  static constexpr int BN = 256;                     // N of the pair tile (each CTA stages 128 rows of W)
  static constexpr int A_BYTES = G_BM * G_BK * 2;    // 16 KB
  static constexpr int B_BYTES = (BN / 2) * G_BK * 2;  // 16 KB
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  // two staging tiles per epilogue warp: a TMA store only has to have finished READING its tile
  // by the time the warp comes back to it two store groups later (the round trip is ~2k cycles)
  static constexpr int NBUF = 2;
  static constexpr int OUT_BYTES = G_EPI_WARPS * NBUF * G_STAGE_TILE;
  static constexpr int MISC_BYTES = 512 + 2 * BN * 4;
  static constexpr int BUDGET = 232448 - OUT_BYTES - MISC_BYTES;
  static constexpr int STAGES = (BUDGET / STAGE_BYTES) > 6 ? 6 : (BUDGET / STAGE_BYTES);
  static constexpr int TMEM_COLS = 2 * BN;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + OUT_BYTES + MISC_BYTES;
};

template <int EPI, bool OUT_BF16>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(G_THREADS, 1)
gemm_bf16_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                     const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR,
                     const float* __restrict__ bias, int M, int N, int K, int dbg) {
  using Cfg = Gemm2Cfg<EPI>;
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
    if (EPI == PCD_EPI_BIAS_RESIDUAL) prefetch_tensormap(&tmR);
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
    static_assert(!(EPI == PCD_EPI_BIAS_RESIDUAL && OUT_BF16), "the residual stream is fp32");
    const int etid = threadIdx.x - 64;
    constexpr int NBUF = Cfg::NBUF;
    unsigned char* my_buf0 = stage_out + ew * NBUF * G_STAGE_TILE;
    int sidx = 0;  // running store-group index of this warp
    uint64_t* my_res_bar = res_bar + 2 * ew;
    const int sw = lane & 7;
    uint32_t res_uses = 0;
    int it = 0;
    for (int t = pair; t < num_tiles; t += num_pairs, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int m_blk = t / num_n, n_blk = t % num_n;
      const int row0 = m_blk * 2 * G_BM + (int)rank * G_BM + quarter * 32;
      const int colbase = n_blk * BN + half * HALF_N;
      float* sb = sbias + as * BN;
      if (etid < BN) {
        const int n = n_blk * BN + etid;
        sb[etid] = (bias != nullptr && n < N) ? __ldg(bias + n) : 0.f;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      mbar_wait(&acc_full[as], aphase);
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
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        uint32_t* r = rbuf[c & 1];
        const bool first_of_store = (c % CH_PER_STORE) == 0;
        const bool last_of_store = (c % CH_PER_STORE) == CH_PER_STORE - 1;
        unsigned char* my_buf = my_buf0 + (sidx % NBUF) * G_STAGE_TILE;
        unsigned char* buf_row = my_buf + lane * 128;
        if (first_of_store) {
          if (elect_one()) {
            tma_store_wait_read<NBUF - 1>();  // the store that used this tile two groups ago has read it
            if (EPI == PCD_EPI_BIAS_RESIDUAL) {
              mbar_expect_tx(&my_res_bar[0], G_STAGE_TILE);
              tma_load_2d(my_buf, &tmR, &my_res_bar[0], colbase + c * 32, row0);
            }
          }
          __syncwarp();
          if (EPI == PCD_EPI_BIAS_RESIDUAL) {
            mbar_wait(&my_res_bar[0], res_uses & 1);
            res_uses++;
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
              const float4 rr = *pp;
              o.x += rr.x; o.y += rr.y; o.z += rr.z; o.w += rr.w;
            }
            *pp = o;
          }
        }
        if (last_of_store) {
          fence_proxy_async_smem();
          __syncwarp();
          if (elect_one()) {
            const int col = colbase + (c / CH_PER_STORE) * (OUT_BF16 ? 64 : 32);
            tma_store_2d(&tmC, my_buf, col, row0);
            tma_store_commit();
          }
          ++sidx;
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(leader_acc_empty);
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

template <int EPI, bool OUT_BF16>
static int launch_gemm2(const CUtensorMap& tmA, const CUtensorMap& tmW, const CUtensorMap& tmC,
                        const CUtensorMap& tmR, const float* bias, int M, int N, int K, cudaStream_t st) {
  using Cfg = Gemm2Cfg<EPI>;
  auto kern = gemm_bf16_tc2_kernel<EPI, OUT_BF16>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) {
      set_error("gemm_bf16(pair): cudaFuncSetAttribute(%d): %s", Cfg::SMEM_BYTES, cudaGetErrorString(e));
      return PCD_ERR_CUDA;
    }
    attr_set = true;
  }
  const int tiles = ceil_div(M, 2 * G_BM) * ceil_div(N, Cfg::BN);
  int pairs = num_sms() / 2;
  if (tiles < pairs) pairs = tiles;
  kern<<<2 * pairs, G_THREADS, Cfg::SMEM_BYTES, st>>>(tmA, tmW, tmC, tmR, bias, M, N, K, g_gemm_debug);
  PCD_CHECK_LAUNCH("gemm_bf16(pair)");
  return PCD_OK;
}

static int dispatch_epi2(const CUtensorMap& a, const CUtensorMap& w, const CUtensorMap& c, const CUtensorMap& r,
                         const float* bias, int out_prec, int M, int N, int K, int epi, cudaStream_t st) {
  const bool ob = out_prec == PCD_BF16;
  switch (epi) {
    case PCD_EPI_BIAS:
      return ob ? launch_gemm2<PCD_EPI_BIAS, true>(a, w, c, r, bias, M, N, K, st)
                : launch_gemm2<PCD_EPI_BIAS, false>(a, w, c, r, bias, M, N, K, st);
    case PCD_EPI_BIAS_GELU:
      return ob ? launch_gemm2<PCD_EPI_BIAS_GELU, true>(a, w, c, r, bias, M, N, K, st)
                : launch_gemm2<PCD_EPI_BIAS_GELU, false>(a, w, c, r, bias, M, N, K, st);
    case PCD_EPI_BIAS_RESIDUAL:
      if (ob) break;
      return launch_gemm2<PCD_EPI_BIAS_RESIDUAL, false>(a, w, c, r, bias, M, N, K, st);
  }
  set_error("gemm_bf16: unsupported epilogue %d / output precision %d", epi, out_prec);
  return PCD_ERR_INVALID;
}

int g_gemm_debug = 0;  // profiling aid, see pcd_set_debug_flags (bit 2: force the single-CTA kernel)

template <int BN, int EPI, bool OUT_BF16>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmW, const CUtensorMap& tmC,
                       const CUtensorMap& tmR, const float* bias, int M, int N, int K, cudaStream_t st) {
  using Cfg = GemmCfg<BN, EPI>;
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
  kern<<<grid, G_THREADS, Cfg::SMEM_BYTES, st>>>(tmA, tmW, tmC, tmR, bias, M, N, K, g_gemm_debug);
  PCD_CHECK_LAUNCH("gemm_bf16");
  return PCD_OK;
}

template <int BN>
static int dispatch_epi(const CUtensorMap& a, const CUtensorMap& w, const CUtensorMap& c, const CUtensorMap& r,
                        const float* bias, int out_prec, int M, int N, int K, int epi, cudaStream_t st) {
  const bool ob = out_prec == PCD_BF16;
  switch (epi) {
    case PCD_EPI_BIAS:
      return ob ? launch_gemm<BN, PCD_EPI_BIAS, true>(a, w, c, r, bias, M, N, K, st)
                : launch_gemm<BN, PCD_EPI_BIAS, false>(a, w, c, r, bias, M, N, K, st);
    case PCD_EPI_BIAS_GELU:
      return ob ? launch_gemm<BN, PCD_EPI_BIAS_GELU, true>(a, w, c, r, bias, M, N, K, st)
                : launch_gemm<BN, PCD_EPI_BIAS_GELU, false>(a, w, c, r, bias, M, N, K, st);
    case PCD_EPI_BIAS_RESIDUAL:
      if (ob) break;
      return launch_gemm<BN, PCD_EPI_BIAS_RESIDUAL, false>(a, w, c, r, bias, M, N, K, st);
  }
  set_error("gemm_bf16: unsupported epilogue %d / output precision %d", epi, out_prec);
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
  const bool ob = out_precision == PCD_BF16;
  PCD_CHECK_ARG(out_precision == PCD_BF16 || out_precision == PCD_F32, "gemm_bf16: bad output precision");
  PCD_CHECK_ARG(ldc % (ob ? 8 : 4) == 0 && (reinterpret_cast<uintptr_t>(C) & 15) == 0, "gemm_bf16: C must be 16-byte aligned with a 16-byte row pitch (ldc=%d)", ldc);
  PCD_CHECK_ARG(epilogue != PCD_EPI_BIAS_RESIDUAL || (residual != nullptr && ldr % 4 == 0 && (reinterpret_cast<uintptr_t>(residual) & 15) == 0 && !ob),
                "gemm_bf16: residual epilogue needs an fp32, 16-byte aligned residual and fp32 output");
  const bool use_pair = (N % 256 == 0) && M >= 512 && !(g_gemm_debug & 4);
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
  if (epilogue == PCD_EPI_BIAS_RESIDUAL) {
    uint64_t strR[1] = {(uint64_t)ldr * 4};
    uint32_t boxR[2] = {32, 32};
    if ((rc = encode_tmap_f32(&tmR, residual, 2, dimsC, strR, boxR)) != PCD_OK) return rc;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (use_pair) return dispatch_epi2(tmA, tmW, tmC, tmR, bias, out_precision, M, N, K, epilogue, st);
  if (BN == 256) return dispatch_epi<256>(tmA, tmW, tmC, tmR, bias, out_precision, M, N, K, epilogue, st);
  return dispatch_epi<128>(tmA, tmW, tmC, tmR, bias, out_precision, M, N, K, epilogue, st);
}
