// bf16 flash attention (head dim 64, non-causal) on tcgen05 tensor cores.
//
// One CTA = one 128-query tile of one (batch, head); 256 threads = 2 warpgroups (warpgroup 0
// hands its registers to the softmax warpgroup with setmaxnreg: 40 vs 216 per thread):
//   warp 0   : TMA producer -- Q once, then a 2-stage ring of K/V tiles (128 keys x 64),
//              read straight out of the packed projection buffers through 4-D tensor maps
//              {64, H, L, B} (rows past L are zero-filled by TMA, so tails need no copies)
//   warp 1   : MMA issuer   -- S = Q K^T (M128 x N128 x K64) and O_j = P_j V_j (M128 x N64 x
//              K128, V consumed as an MN-major operand), accumulators in TMEM
//   warps 4-7: softmax      -- thread = query row: the whole 128-wide S row is pulled into
//              registers with four back-to-back tcgen05.ld (one wait), fp32 softmax in the exp2
//              domain (scale folded into one FFMA), P_j written back as bf16 (to TMEM, or to
//              swizzled shared memory in the SS variant).  O accumulates across KV tiles in TMEM
//              (PV MMAs with accumulate); the running maximum is only refreshed -- and O / l
//              rescaled in TMEM -- when a row maximum grows by more than 2^8 (lazy rescaling),
//              so the common iteration touches O not at all.
// TMEM budget 256 columns (S 128 | P 64 | O 64) so two CTAs are co-resident per SM and
// one CTA's softmax overlaps the other's MMAs.
#include "common.cuh"
#include "tc_sm100.cuh"

namespace pcd {

using namespace tc;

constexpr int T_BQ = 128, T_BKV = 128, T_HD = 64;
constexpr int T_TILE_BYTES = T_BKV * T_HD * 2;  // 16 KB (Q, K and V tiles alike)

template <bool P_TMEM>
struct AttnCfg {
  static constexpr int KV_STAGES = 2;
  static constexpr int P_BYTES = P_TMEM ? 0 : 2 * T_TILE_BYTES;
  static constexpr int TILE_BYTES = T_TILE_BYTES * (1 + 2 * KV_STAGES) + P_BYTES;
  static constexpr int SMEM_BYTES = TILE_BYTES + 1024 + 128;
  static constexpr int TMEM_COLS = 256;
  static constexpr int S_COL = 0, P_COL = 128, O_COL = 192;
};

template <bool P_TMEM>
__global__ void __launch_bounds__(256, P_TMEM ? 2 : 1)
attn_bf16_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, uint16_t* __restrict__ out, int64_t o_bs,
                    int64_t o_ls, int len_q, int len_kv, float scale_log2) {
  using Cfg = AttnCfg<P_TMEM>;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* sQ = smem;
  unsigned char* sK = sQ + T_TILE_BYTES;                        // [KV_STAGES]
  unsigned char* sV = sK + Cfg::KV_STAGES * T_TILE_BYTES;       // [KV_STAGES]
  unsigned char* sP = sV + Cfg::KV_STAGES * T_TILE_BYTES;       // SS variant only (2 slabs of 64 keys)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::TILE_BYTES);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;    // [2]
  uint64_t* kv_empty = bars + 3;   // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* p_ready = bars + 6;
  uint64_t* o_full = bars + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * T_BQ, h = blockIdx.y, b = blockIdx.z;
  const int num_kv = (len_kv + T_BKV - 1) / T_BKV;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmQ);
    prefetch_tensormap(&tmK);
    prefetch_tensormap(&tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_ready, 128);
    mbar_init(o_full, 1);
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

  if (warp < 4) {
  // warpgroup 0 (TMA, MMA, two idle warps) keeps 40 registers per thread; the freed ones go to
  // the softmax warpgroup below (setmaxnreg must sit inside the role branch for ptxas to apply
  // the per-region budget)
  setmaxnreg_dec<40>();
  if (warp == 0) {
    // --------------------------- TMA producer ---------------------------
    if (elect_one()) {
      mbar_expect_tx(q_full, T_TILE_BYTES);
      tma_load_4d(sQ, &tmQ, q_full, 0, h, q0, b);
      for (int j = 0; j < num_kv; ++j) {
        const int st = j & 1;
        mbar_wait(&kv_empty[st], ((j >> 1) & 1) ^ 1);
        mbar_expect_tx(&kv_full[st], 2 * T_TILE_BYTES);
        tma_load_4d(sK + st * T_TILE_BYTES, &tmK, &kv_full[st], 0, h, j * T_BKV, b);
        tma_load_4d(sV + st * T_TILE_BYTES, &tmV, &kv_full[st], 0, h, j * T_BKV, b);
      }
    }
  } else if (warp == 1) {
    // ---------------------------- MMA issuer ----------------------------
    const uint32_t s_tmem = tmem_base + Cfg::S_COL;
    const uint32_t p_tmem = tmem_base + Cfg::P_COL;
    const uint32_t o_tmem = tmem_base + Cfg::O_COL;
    constexpr uint32_t idesc_pv = idesc_bf16_f32(T_BQ, T_HD, /*B MN-major*/ 1);
    auto issue_qk = [&](int j) {
      const int st = j & 1;
      mbar_wait(&kv_full[st], (j >> 1) & 1);
      tcgen05_fence_after();
      if (elect_one()) {
        int nvalid = min(T_BKV, len_kv - j * T_BKV);
        int n = max(16, (nvalid + 15) & ~15);
        const uint32_t idesc_qk = idesc_bf16_f32(T_BQ, n, 0);
        const uint64_t adesc = smem_desc_sw128(smem_u32(sQ));
        const uint64_t bdesc = smem_desc_sw128(smem_u32(sK + st * T_TILE_BYTES));
#pragma unroll
        for (int k = 0; k < T_HD / 16; ++k) umma_bf16_ss(s_tmem, adesc + 2 * k, bdesc + 2 * k, idesc_qk, k != 0);
        umma_commit(s_full);
      }
      __syncwarp();
    };
    mbar_wait(q_full, 0);
    issue_qk(0);
    for (int j = 0; j < num_kv; ++j) {
      mbar_wait(p_ready, j & 1);
      tcgen05_fence_after();
      if (j + 1 < num_kv) issue_qk(j + 1);  // S is free: every softmax thread has read S_j
      if (elect_one()) {
        const int st = j & 1;
        int nvalid = min(T_BKV, len_kv - j * T_BKV);
        int nks = (nvalid + 15) >> 4;  // 16 keys per MMA
        const uint32_t v_addr = smem_u32(sV + st * T_TILE_BYTES);
        for (int kk = 0; kk < nks; ++kk) {
          // V tile: 128 key rows of 128 B; 16 keys = two 8-row swizzle groups (SBO 1024)
          const uint64_t bdesc = smem_desc_sw128(v_addr + kk * 2048);
          if (P_TMEM) {
            umma_bf16_ts(o_tmem, p_tmem + kk * 8, bdesc, idesc_pv, (j | kk) != 0);
          } else {
            const uint64_t adesc = smem_desc_sw128(smem_u32(sP) + (kk >> 2) * T_TILE_BYTES + (kk & 3) * 32);
            umma_bf16_ss(o_tmem, adesc, bdesc, idesc_pv, (j | kk) != 0);
          }
        }
        umma_commit(&kv_empty[st]);
        umma_commit(o_full);
      }
      __syncwarp();
    }
  }
  } else {
    setmaxnreg_inc<216>();
    // ----------------------------- softmax ------------------------------
    const int quarter = warp & 3;
    const int row_in_tile = quarter * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t s_tmem = tmem_base + lane_off + Cfg::S_COL;
    const uint32_t p_tmem = tmem_base + lane_off + Cfg::P_COL;
    const uint32_t o_tmem = tmem_base + lane_off + Cfg::O_COL;
    if (q0 + quarter * 32 >= len_q) {
      // every row of this warp is past the end of the sequence (ragged last query tile):
      // skip the softmax work, only keep the hand-off barrier counts in step
      for (int j = 0; j < num_kv; ++j) {
        mbar_arrive(p_ready);
        mbar_wait(p_ready, j & 1);
      }
    } else {
    float m_used = 0.f, l_run = 0.f;   // m_used: the maximum currently baked into P, l and O
    constexpr float kRescaleThreshold = 8.f;  // log2 units: P <= 2^8, safe in bf16 / fp32

    for (int j = 0; j < num_kv; ++j) {
      const int nvalid = min(T_BKV, len_kv - j * T_BKV);
      mbar_wait(s_full, j & 1);
      tcgen05_fence_after();
      // the whole S row -> registers (columns past nvalid are stale: masked below)
      uint32_t sr[T_BKV];
#pragma unroll
      for (int c = 0; c < T_BKV / 32; ++c)
        if (c * 32 < nvalid) tmem_ld_32x32b_x32(s_tmem + c * 32, sr + c * 32);
      tmem_ld_wait();
      if (nvalid < T_BKV) {
#pragma unroll
        for (int i = 0; i < T_BKV; ++i)
          if (i >= nvalid) sr[i] = 0xff800000u;  // -inf
      }
      float mx = -INFINITY;
#pragma unroll
      for (int i = 0; i < T_BKV; ++i) mx = fmaxf(mx, __uint_as_float(sr[i]));
      mx *= scale_log2;
      if (j == 0) {
        m_used = mx;
      } else {
        const bool grow = mx > m_used + kRescaleThreshold;
        if (__any_sync(0xffffffffu, grow)) {
          // rare path: rescale this warp's rows of O (and l) to the new maximum.  PV_{j-1} has
          // to be complete; PV_j is only issued after this thread arrives on p_ready below.
          mbar_wait(o_full, (j - 1) & 1);
          tcgen05_fence_after();
          const float alpha = grow ? ex2_approx(m_used - mx) : 1.f;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t r[32];
            tmem_ld_32x32b_x32(o_tmem + c * 32, r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
            tmem_st_32x32b_x32(o_tmem + c * 32, r);
          }
          tmem_st_wait();
          l_run *= alpha;
          if (grow) m_used = mx;
        }
      }
      // P = exp2(S*scale - m_used), row sum, bf16 pack, hand-off (32 columns at a time so the
      // S registers retire as the packed P words are produced)
      float s0 = 0.f, s1 = 0.f;
#pragma unroll
      for (int c = 0; c < T_BKV / 32; ++c) {
        if (c * 32 < nvalid) {
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float p0 = ex2_approx(fmaf(__uint_as_float(sr[c * 32 + i]), scale_log2, -m_used));
            const float p1 = ex2_approx(fmaf(__uint_as_float(sr[c * 32 + i + 1]), scale_log2, -m_used));
            s0 += p0;
            s1 += p1;
            pk[i >> 1] = pack_bf16x2(p0, p1);
          }
          if (P_TMEM) {
            tmem_st_32x32b_x16(p_tmem + c * 16, pk);
          } else {
            // K-major bf16 tile, 128B swizzle: 16-byte chunk index ^= (row & 7)
            unsigned char* prow = sP + (c >> 1) * T_TILE_BYTES + row_in_tile * 128;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int chunk = ((c & 1) * 4 + i) ^ (row_in_tile & 7);
              *reinterpret_cast<uint4*>(prow + chunk * 16) = make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
            }
          }
        }
      }
      l_run += s0 + s1;
      if (P_TMEM) {
        tmem_st_wait();
      } else {
        fence_proxy_async_smem();
      }
      tcgen05_fence_before();
      mbar_arrive(p_ready);
    }
    // epilogue: O / l
    mbar_wait(o_full, (num_kv - 1) & 1);
    tcgen05_fence_after();
    const float inv = 1.f / l_run;
    const int row = q0 + row_in_tile;
    uint16_t* orow = out + b * o_bs + (int64_t)row * o_ls + h * T_HD;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t r[32];
      tmem_ld_32x32b_x32(o_tmem + c * 32, r);
      tmem_ld_wait();
      if (row < len_q) {
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          float v[8];
#pragma unroll
          for (int t = 0; t < 8; ++t) v[t] = __uint_as_float(r[i + t]) * inv;
          *reinterpret_cast<uint4*>(orow + c * 32 + i) =
              make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
        }
      }
    }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------
// Ping-pong variant (default): one CTA per SM owns TWO 128-query tiles of one (batch, head).
// They share the K/V ring; each has its own softmax warpgroup and its own S / P / O columns in
// TMEM (S0 S1 | P0 P1 | O0 O1 = 512 columns).  The single MMA thread alternates between the tiles,
// so while one warpgroup exponentiates its S tile (the MUFU-bound part: 16 ex2/clk/SM) the tensor
// core computes the other tile's QK^T / PV -- the two independent CTAs of the first design ran in
// lock-step instead (both in softmax, then both waiting on the tensor core).
// ---------------------------------------------------------------------------
struct Attn2Cfg {
  static constexpr int KV_STAGES = 3;
  static constexpr int TILE_BYTES = T_TILE_BYTES * (2 + 2 * KV_STAGES);
  static constexpr int SMEM_BYTES = TILE_BYTES + 1024 + 256;
  static constexpr int TMEM_COLS = 512;
  static constexpr int S_COL = 0, P_COL = 256, O_COL = 384;  // + tile*128 / tile*64 / tile*64
};

__global__ void __launch_bounds__(384, 1)
attn_bf16_tc2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, uint16_t* __restrict__ out, int64_t o_bs,
                     int64_t o_ls, int len_q, int len_kv, float scale_log2) {
  using Cfg = Attn2Cfg;
  constexpr int KS = Cfg::KV_STAGES;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* sQ = smem;                                   // [2]
  unsigned char* sK = sQ + 2 * T_TILE_BYTES;                  // [KS]
  unsigned char* sV = sK + KS * T_TILE_BYTES;                 // [KS]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::TILE_BYTES);
  uint64_t* q_full = bars;            // [1]
  uint64_t* kv_full = bars + 1;       // [KS]
  uint64_t* kv_empty = kv_full + KS;  // [KS]
  uint64_t* s_full = kv_empty + KS;   // [2]
  uint64_t* p_ready = s_full + 2;     // [2]
  uint64_t* o_full = p_ready + 2;     // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 2 * T_BQ, h = blockIdx.y, b = blockIdx.z;
  const int num_kv = (len_kv + T_BKV - 1) / T_BKV;
  const int ntq = (q0 + T_BQ < len_q) ? 2 : 1;  // second query tile present?

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmQ);
    prefetch_tensormap(&tmK);
    prefetch_tensormap(&tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < KS; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&p_ready[t], 128);
      mbar_init(&o_full[t], 1);
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

  if (warp < 4) {
    setmaxnreg_dec<40>();
    if (warp == 0) {
      // --------------------------- TMA producer ---------------------------
      if (elect_one()) {
        mbar_expect_tx(q_full, ntq * T_TILE_BYTES);
        for (int t = 0; t < ntq; ++t) tma_load_4d(sQ + t * T_TILE_BYTES, &tmQ, q_full, 0, h, q0 + t * T_BQ, b);
        for (int j = 0; j < num_kv; ++j) {
          const int st = j % KS;
          mbar_wait(&kv_empty[st], ((j / KS) & 1) ^ 1);
          mbar_expect_tx(&kv_full[st], 2 * T_TILE_BYTES);
          tma_load_4d(sK + st * T_TILE_BYTES, &tmK, &kv_full[st], 0, h, j * T_BKV, b);
          tma_load_4d(sV + st * T_TILE_BYTES, &tmV, &kv_full[st], 0, h, j * T_BKV, b);
        }
      }
    } else if (warp == 1) {
      // ---------------------------- MMA issuer ----------------------------
      constexpr uint32_t idesc_pv = idesc_bf16_f32(T_BQ, T_HD, /*B MN-major*/ 1);
      auto issue_qk = [&](int t, int j) {
        if (elect_one()) {
          const int st = j % KS;
          const int nvalid = min(T_BKV, len_kv - j * T_BKV);
          const int n = max(16, (nvalid + 15) & ~15);
          const uint32_t idesc_qk = idesc_bf16_f32(T_BQ, n, 0);
          const uint64_t adesc = smem_desc_sw128(smem_u32(sQ + t * T_TILE_BYTES));
          const uint64_t bdesc = smem_desc_sw128(smem_u32(sK + st * T_TILE_BYTES));
          const uint32_t s_tmem = tmem_base + Cfg::S_COL + t * T_BKV;
#pragma unroll
          for (int k = 0; k < T_HD / 16; ++k) umma_bf16_ss(s_tmem, adesc + 2 * k, bdesc + 2 * k, idesc_qk, k != 0);
          umma_commit(&s_full[t]);
        }
        __syncwarp();
      };
      auto issue_pv = [&](int t, int j, bool release_kv) {
        if (elect_one()) {
          const int st = j % KS;
          const int nvalid = min(T_BKV, len_kv - j * T_BKV);
          const int nks = (nvalid + 15) >> 4;  // 16 keys per MMA
          const uint32_t v_addr = smem_u32(sV + st * T_TILE_BYTES);
          const uint32_t p_tmem = tmem_base + Cfg::P_COL + t * (T_BKV / 2);
          const uint32_t o_tmem = tmem_base + Cfg::O_COL + t * T_HD;
          for (int kk = 0; kk < nks; ++kk)
            umma_bf16_ts(o_tmem, p_tmem + kk * 8, smem_desc_sw128(v_addr + kk * 2048), idesc_pv, (j | kk) != 0);
          if (release_kv) umma_commit(&kv_empty[st]);
          umma_commit(&o_full[t]);
        }
        __syncwarp();
      };
      mbar_wait(q_full, 0);
      mbar_wait(&kv_full[0], 0);
      tcgen05_fence_after();
      for (int t = 0; t < ntq; ++t) issue_qk(t, 0);
      for (int j = 0; j < num_kv; ++j) {
        for (int t = 0; t < ntq; ++t) {
          mbar_wait(&p_ready[t], j & 1);  // P_t(j) is in TMEM and S_t has been fully read
          tcgen05_fence_after();
          if (j + 1 < num_kv) {
            if (t == 0) {
              mbar_wait(&kv_full[(j + 1) % KS], ((j + 1) / KS) & 1);
              tcgen05_fence_after();
            }
            issue_qk(t, j + 1);  // next S for this tile first: it is on the softmax critical path
          }
          issue_pv(t, j, t == ntq - 1);
        }
      }
    }
  } else {
    setmaxnreg_inc<232>();
    // ----------------------------- softmax ------------------------------
    const int t = (warp - 4) >> 2;  // query tile of this warpgroup
    const int quarter = warp & 3;
    const int tq0 = q0 + t * T_BQ;
    if (t < ntq) {
      const int row_in_tile = quarter * 32 + lane;
      const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
      const uint32_t s_tmem = tmem_base + lane_off + Cfg::S_COL + t * T_BKV;
      const uint32_t p_tmem = tmem_base + lane_off + Cfg::P_COL + t * (T_BKV / 2);
      const uint32_t o_tmem = tmem_base + lane_off + Cfg::O_COL + t * T_HD;
      uint64_t* my_s_full = &s_full[t];
      uint64_t* my_p_ready = &p_ready[t];
      uint64_t* my_o_full = &o_full[t];
      if (tq0 + quarter * 32 >= len_q) {
        // all rows of this warp lie past the end of the sequence: keep the barrier counts in step
        for (int j = 0; j < num_kv; ++j) {
          mbar_arrive(my_p_ready);
          mbar_wait(my_p_ready, j & 1);
        }
      } else {
        float m_used = 0.f, l_run = 0.f;
        constexpr float kRescaleThreshold = 8.f;
        for (int j = 0; j < num_kv; ++j) {
          const int nvalid = min(T_BKV, len_kv - j * T_BKV);
          mbar_wait(my_s_full, j & 1);
          tcgen05_fence_after();
          uint32_t sr[T_BKV];
#pragma unroll
          for (int c = 0; c < T_BKV / 32; ++c)
            if (c * 32 < nvalid) tmem_ld_32x32b_x32(s_tmem + c * 32, sr + c * 32);
          tmem_ld_wait();
          if (nvalid < T_BKV) {
#pragma unroll
            for (int i = 0; i < T_BKV; ++i)
              if (i >= nvalid) sr[i] = 0xff800000u;  // -inf
          }
          float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
          for (int i = 0; i < T_BKV; i += 4) {
            mx0 = fmaxf(mx0, __uint_as_float(sr[i]));
            mx1 = fmaxf(mx1, __uint_as_float(sr[i + 1]));
            mx2 = fmaxf(mx2, __uint_as_float(sr[i + 2]));
            mx3 = fmaxf(mx3, __uint_as_float(sr[i + 3]));
          }
          const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * scale_log2;
          if (j == 0) {
            m_used = mx;
          } else {
            const bool grow = mx > m_used + kRescaleThreshold;
            if (__any_sync(0xffffffffu, grow)) {
              mbar_wait(my_o_full, (j - 1) & 1);  // PV_t(j-1) complete; PV_t(j) not yet issued
              tcgen05_fence_after();
              const float alpha = grow ? ex2_approx(m_used - mx) : 1.f;
#pragma unroll
              for (int c = 0; c < 2; ++c) {
                uint32_t r[32];
                tmem_ld_32x32b_x32(o_tmem + c * 32, r);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
                tmem_st_32x32b_x32(o_tmem + c * 32, r);
              }
              tmem_st_wait();
              l_run *= alpha;
              if (grow) m_used = mx;
            }
          }
          float s0 = 0.f, s1 = 0.f;
#pragma unroll
          for (int c = 0; c < T_BKV / 32; ++c) {
            if (c * 32 < nvalid) {
              uint32_t pk[16];
#pragma unroll
              for (int i = 0; i < 32; i += 2) {
                const float p0 = ex2_approx(fmaf(__uint_as_float(sr[c * 32 + i]), scale_log2, -m_used));
                const float p1 = ex2_approx(fmaf(__uint_as_float(sr[c * 32 + i + 1]), scale_log2, -m_used));
                s0 += p0;
                s1 += p1;
                pk[i >> 1] = pack_bf16x2(p0, p1);
              }
              tmem_st_32x32b_x16(p_tmem + c * 16, pk);
            }
          }
          l_run += s0 + s1;
          tmem_st_wait();
          tcgen05_fence_before();
          mbar_arrive(my_p_ready);
        }
        // epilogue: O / l
        mbar_wait(my_o_full, (num_kv - 1) & 1);
        tcgen05_fence_after();
        const float inv = 1.f / l_run;
        const int row = tq0 + row_in_tile;
        uint16_t* orow = out + b * o_bs + (int64_t)row * o_ls + h * T_HD;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(o_tmem + c * 32, r);
          tmem_ld_wait();
          if (row < len_q) {
#pragma unroll
            for (int i = 0; i < 32; i += 8) {
              float v[8];
#pragma unroll
              for (int u = 0; u < 8; ++u) v[u] = __uint_as_float(r[i + u]) * inv;
              *reinterpret_cast<uint4*>(orow + c * 32 + i) =
                  make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
            }
          }
        }
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

static int launch_attn_tc2(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, uint16_t* out,
                           int64_t o_bs, int64_t o_ls, int batch, int heads, int len_q, int len_kv,
                           float scale_log2, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_bf16_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Attn2Cfg::SMEM_BYTES);
    if (e != cudaSuccess) {
      set_error("attention_bf16: cudaFuncSetAttribute(%d): %s", Attn2Cfg::SMEM_BYTES, cudaGetErrorString(e));
      return PCD_ERR_CUDA;
    }
    attr_set = true;
  }
  dim3 grid(ceil_div(ceil_div(len_q, T_BQ), 2), heads, batch);
  attn_bf16_tc2_kernel<<<grid, 384, Attn2Cfg::SMEM_BYTES, st>>>(tq, tk, tv, out, o_bs, o_ls, len_q, len_kv, scale_log2);
  PCD_CHECK_LAUNCH("attention_bf16");
  return PCD_OK;
}

// ---------------------------------------------------------------------------
// Double-buffered variant (default).  One 128-query tile per CTA, two CTAs per SM, KV tiles of
// 64 keys.  S and P are DOUBLE-BUFFERED in TMEM (S0 S1 | P0 P1 | O = 64+64+32+32+64 = 256
// columns), so the tensor core computes S_{j+1}/S_{j+2} and O += P_j V_j while the softmax
// warpgroup exponentiates S_j: the softmax warps never wait for an MMA round trip and the kernel
// runs at the MUFU rate (16 ex2/clk/SM, measured with tools/ubench) instead of alternating
// between softmax and tensor phases.
// ---------------------------------------------------------------------------
struct Attn3Cfg {
  static constexpr int BKV = 64;
  static constexpr int KV_TILE = BKV * T_HD * 2;   // 8 KB per K or V tile
  static constexpr int KV_STAGES = 5;
  static constexpr int TILE_BYTES = T_TILE_BYTES + KV_STAGES * 2 * KV_TILE;  // Q + ring
  static constexpr int SMEM_BYTES = TILE_BYTES + 1024 + 256;
  static constexpr int TMEM_COLS = 256;
  static constexpr int S_COL = 0, P_COL = 128, O_COL = 192;  // S: +buf*64, P: +buf*32
};

__global__ void __launch_bounds__(256, 2)
attn_bf16_tc3_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, uint16_t* __restrict__ out, int64_t o_bs,
                     int64_t o_ls, int len_q, int len_kv, float scale_log2) {
  using Cfg = Attn3Cfg;
  constexpr int KS = Cfg::KV_STAGES;
  constexpr int BKV = Cfg::BKV;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* sQ = smem;
  unsigned char* sK = sQ + T_TILE_BYTES;            // [KS] 8 KB each
  unsigned char* sV = sK + KS * Cfg::KV_TILE;       // [KS]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::TILE_BYTES);
  uint64_t* q_full = bars;            // [1]
  uint64_t* kv_full = bars + 1;       // [KS]
  uint64_t* kv_empty = kv_full + KS;  // [KS]
  uint64_t* s_full = kv_empty + KS;   // [2]  QK_j done            (MMA -> softmax)
  uint64_t* p_ready = s_full + 2;     // [2]  P_j written, S_j read (softmax -> MMA)
  uint64_t* pv_done = p_ready + 2;    // [2]  PV_j done: P buffer free, O updated (MMA -> softmax)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * T_BQ, h = blockIdx.y, b = blockIdx.z;
  const int num_kv = (len_kv + BKV - 1) / BKV;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmQ);
    prefetch_tensormap(&tmK);
    prefetch_tensormap(&tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < KS; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&p_ready[s], 4);  // one arrival per softmax warp
      mbar_init(&pv_done[s], 1);
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

  if (warp < 4) {
    setmaxnreg_dec<40>();
    if (warp == 0) {
      // --------------------------- TMA producer ---------------------------
      if (elect_one()) {
        mbar_expect_tx(q_full, T_TILE_BYTES);
        tma_load_4d(sQ, &tmQ, q_full, 0, h, q0, b);
        for (int j = 0; j < num_kv; ++j) {
          const int st = j % KS;
          mbar_wait(&kv_empty[st], ((j / KS) & 1) ^ 1);
          mbar_expect_tx(&kv_full[st], 2 * Cfg::KV_TILE);
          tma_load_4d(sK + st * Cfg::KV_TILE, &tmK, &kv_full[st], 0, h, j * BKV, b);
          tma_load_4d(sV + st * Cfg::KV_TILE, &tmV, &kv_full[st], 0, h, j * BKV, b);
        }
      }
    } else if (warp == 1) {
      // ---------------------------- MMA issuer ----------------------------
      constexpr uint32_t idesc_pv = idesc_bf16_f32(T_BQ, T_HD, /*B MN-major*/ 1);
      auto issue_qk = [&](int j) {
        const int st = j % KS;
        mbar_wait(&kv_full[st], (j / KS) & 1);
        tcgen05_fence_after();
        if (elect_one()) {
          const int nvalid = min(BKV, len_kv - j * BKV);
          const int n = max(16, (nvalid + 15) & ~15);
          const uint32_t idesc_qk = idesc_bf16_f32(T_BQ, n, 0);
          const uint64_t adesc = smem_desc_sw128(smem_u32(sQ));
          const uint64_t bdesc = smem_desc_sw128(smem_u32(sK + st * Cfg::KV_TILE));
          const uint32_t s_tmem = tmem_base + Cfg::S_COL + (j & 1) * BKV;
#pragma unroll
          for (int k = 0; k < T_HD / 16; ++k) umma_bf16_ss(s_tmem, adesc + 2 * k, bdesc + 2 * k, idesc_qk, k != 0);
          umma_commit(&s_full[j & 1]);
        }
        __syncwarp();
      };
      mbar_wait(q_full, 0);
      issue_qk(0);
      if (num_kv > 1) issue_qk(1);
      for (int j = 0; j < num_kv; ++j) {
        const int bf = j & 1;
        mbar_wait(&p_ready[bf], (j >> 1) & 1);  // P_j in TMEM, S buffer bf fully read
        tcgen05_fence_after();
        if (elect_one()) {
          const int st = j % KS;
          const int nvalid = min(BKV, len_kv - j * BKV);
          const int nks = (nvalid + 15) >> 4;  // 16 keys per MMA
          const uint32_t v_addr = smem_u32(sV + st * Cfg::KV_TILE);
          const uint32_t p_tmem = tmem_base + Cfg::P_COL + bf * (BKV / 2);
          const uint32_t o_tmem = tmem_base + Cfg::O_COL;
          for (int kk = 0; kk < nks; ++kk)
            umma_bf16_ts(o_tmem, p_tmem + kk * 8, smem_desc_sw128(v_addr + kk * 2048), idesc_pv, (j | kk) != 0);
          umma_commit(&kv_empty[st]);
          umma_commit(&pv_done[bf]);
        }
        __syncwarp();
        if (j + 2 < num_kv) issue_qk(j + 2);  // refill the S buffer that was just drained
      }
    }
  } else {
    setmaxnreg_inc<216>();
    // ----------------------------- softmax ------------------------------
    const int quarter = warp & 3;
    const int row_in_tile = quarter * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t o_tmem = tmem_base + lane_off + Cfg::O_COL;
    if (q0 + quarter * 32 >= len_q) {
      // every row of this warp is past the end of the sequence: only keep the counts in step
      for (int j = 0; j < num_kv; ++j) {
        if (lane == 0) mbar_arrive(&p_ready[j & 1]);
        mbar_wait(&p_ready[j & 1], (j >> 1) & 1);
      }
    } else {
      float m_used = 0.f, l_run = 0.f;
      constexpr float kRescaleThreshold = 8.f;
      for (int j = 0; j < num_kv; ++j) {
        const int bf = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        const int nvalid = min(BKV, len_kv - j * BKV);
        const uint32_t s_tmem = tmem_base + lane_off + Cfg::S_COL + bf * BKV;
        const uint32_t p_tmem = tmem_base + lane_off + Cfg::P_COL + bf * (BKV / 2);
        mbar_wait(&s_full[bf], ph);
        tcgen05_fence_after();
        uint32_t sr[BKV];
#pragma unroll
        for (int c = 0; c < BKV / 32; ++c)
          if (c * 32 < nvalid) tmem_ld_32x32b_x32(s_tmem + c * 32, sr + c * 32);
        tmem_ld_wait();
        if (nvalid < BKV) {
#pragma unroll
          for (int i = 0; i < BKV; ++i)
            if (i >= nvalid) sr[i] = 0xff800000u;  // -inf
        }
        float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
        for (int i = 0; i < BKV; i += 4) {
          mx0 = fmaxf(mx0, __uint_as_float(sr[i]));
          mx1 = fmaxf(mx1, __uint_as_float(sr[i + 1]));
          mx2 = fmaxf(mx2, __uint_as_float(sr[i + 2]));
          mx3 = fmaxf(mx3, __uint_as_float(sr[i + 3]));
        }
        const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * scale_log2;
        if (j == 0) {
          m_used = mx;
        } else {
          const bool grow = mx > m_used + kRescaleThreshold;
          if (__any_sync(0xffffffffu, grow)) {
            // rare: bring O and l to the new maximum.  PV_{j-1} (and all earlier) must be done;
            // PV_j is not issued before this thread arrives on p_ready below.
            mbar_wait(&pv_done[(j - 1) & 1], ((j - 1) >> 1) & 1);
            tcgen05_fence_after();
            const float alpha = grow ? ex2_approx(m_used - mx) : 1.f;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              uint32_t r[32];
              tmem_ld_32x32b_x32(o_tmem + c * 32, r);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
              tmem_st_32x32b_x32(o_tmem + c * 32, r);
            }
            tmem_st_wait();
            l_run *= alpha;
            if (grow) m_used = mx;
          }
        }
        // P buffer bf was last read by PV_{j-2}
        if (j >= 2) {
          mbar_wait(&pv_done[bf], ((j - 2) >> 1) & 1);
          tcgen05_fence_after();
        }
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int c = 0; c < BKV / 32; ++c) {
          if (c * 32 < nvalid) {
            uint32_t pk[16];
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              const float p0 = ex2_approx(fmaf(__uint_as_float(sr[c * 32 + i]), scale_log2, -m_used));
              const float p1 = ex2_approx(fmaf(__uint_as_float(sr[c * 32 + i + 1]), scale_log2, -m_used));
              s0 += p0;
              s1 += p1;
              pk[i >> 1] = pack_bf16x2(p0, p1);
            }
            tmem_st_32x32b_x16(p_tmem + c * 16, pk);
          }
        }
        l_run += s0 + s1;
        tmem_st_wait();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_ready[bf]);
      }
      // epilogue: O / l
      mbar_wait(&pv_done[(num_kv - 1) & 1], ((num_kv - 1) >> 1) & 1);
      tcgen05_fence_after();
      const float inv = 1.f / l_run;
      const int row = q0 + row_in_tile;
      uint16_t* orow = out + b * o_bs + (int64_t)row * o_ls + h * T_HD;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(o_tmem + c * 32, r);
        tmem_ld_wait();
        if (row < len_q) {
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = __uint_as_float(r[i + u]) * inv;
            *reinterpret_cast<uint4*>(orow + c * 32 + i) =
                make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
          }
        }
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

static int launch_attn_tc3(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, uint16_t* out,
                           int64_t o_bs, int64_t o_ls, int batch, int heads, int len_q, int len_kv,
                           float scale_log2, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_bf16_tc3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Attn3Cfg::SMEM_BYTES);
    if (e != cudaSuccess) {
      set_error("attention_bf16: cudaFuncSetAttribute(%d): %s", Attn3Cfg::SMEM_BYTES, cudaGetErrorString(e));
      return PCD_ERR_CUDA;
    }
    attr_set = true;
  }
  dim3 grid(ceil_div(len_q, T_BQ), heads, batch);
  attn_bf16_tc3_kernel<<<grid, 256, Attn3Cfg::SMEM_BYTES, st>>>(tq, tk, tv, out, o_bs, o_ls, len_q, len_kv, scale_log2);
  PCD_CHECK_LAUNCH("attention_bf16");
  return PCD_OK;
}

// ---------------------------------------------------------------------------
// Split-row variant (default): like the double-buffered kernel above, but every query row is
// shared by TWO softmax threads (warps w and w+4 own the same TMEM lane quarter and split the 64
// keys of a KV tile 32/32).  That doubles the number of softmax warps (16 per SM, 4 per
// scheduler) so the per-iteration latencies of one warp (barrier wake-up, TMEM load/store
// round trips) hide behind the exponentials of the others.  The two threads of a row only
// synchronise through a 64-thread named barrier per iteration (to agree on lazy rescaling) and
// once at the end (row sum).
// ---------------------------------------------------------------------------
// 64-thread named barrier of the warp pair that shares TMEM lane quarter q (literal barrier ids,
// so that ptxas reserves 5 hardware barriers per CTA instead of all 16)
__device__ __forceinline__ void pair_sync(int q) {
  switch (q) {
    case 0: asm volatile("bar.sync 1, 64;" ::: "memory"); break;
    case 1: asm volatile("bar.sync 2, 64;" ::: "memory"); break;
    case 2: asm volatile("bar.sync 3, 64;" ::: "memory"); break;
    default: asm volatile("bar.sync 4, 64;" ::: "memory"); break;
  }
}

struct Attn4Cfg {
  static constexpr int BKV = 64;
  static constexpr int KV_TILE = BKV * T_HD * 2;   // 8 KB per K or V tile
  static constexpr int KV_STAGES = 5;
  static constexpr int TILE_BYTES = T_TILE_BYTES + KV_STAGES * 2 * KV_TILE;  // Q + ring
  static constexpr int XCH_BYTES = 128 * 2 * 4 + 64;  // per-row exchange floats + 16 pair flags
  static constexpr int SMEM_BYTES = TILE_BYTES + 1024 + 256 + XCH_BYTES;
  static constexpr int TMEM_COLS = 256;
  static constexpr int S_COL = 0, P_COL = 128, O_COL = 192;  // S: +buf*64, P: +buf*32
};

__global__ void __launch_bounds__(384, 2)
attn_bf16_tc4_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, uint16_t* __restrict__ out, int64_t o_bs,
                     int64_t o_ls, int len_q, int len_kv, float scale_log2) {
  using Cfg = Attn4Cfg;
  constexpr int KS = Cfg::KV_STAGES;
  constexpr int BKV = Cfg::BKV;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* sQ = smem;
  unsigned char* sK = sQ + T_TILE_BYTES;            // [KS] 8 KB each
  unsigned char* sV = sK + KS * Cfg::KV_TILE;       // [KS]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::TILE_BYTES);
  uint64_t* q_full = bars;            // [1]
  uint64_t* kv_full = bars + 1;       // [KS]
  uint64_t* kv_empty = kv_full + KS;  // [KS]
  uint64_t* s_full = kv_empty + KS;   // [2]  QK_j done            (MMA -> softmax)
  uint64_t* p_ready = s_full + 2;     // [2]  P_j written, S_j read (softmax -> MMA)
  uint64_t* pv_done = p_ready + 2;    // [2]  PV_j done: P buffer free, O updated (MMA -> softmax)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 2);
  float* xch = reinterpret_cast<float*>(smem + Cfg::TILE_BYTES + 256);        // [128 rows][2 halves]
  volatile int* pair_flag = reinterpret_cast<volatile int*>(xch + 256);       // [4 quarters][2 halves]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * T_BQ, h = blockIdx.y, b = blockIdx.z;
  const int num_kv = (len_kv + BKV - 1) / BKV;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmQ);
    prefetch_tensormap(&tmK);
    prefetch_tensormap(&tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < KS; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&p_ready[s], 8);  // one arrival per softmax warp
      mbar_init(&pv_done[s], 1);
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

  if (warp < 4) {
    setmaxnreg_dec<40>();
    if (warp == 0) {
      // --------------------------- TMA producer ---------------------------
      if (elect_one()) {
        mbar_expect_tx(q_full, T_TILE_BYTES);
        tma_load_4d(sQ, &tmQ, q_full, 0, h, q0, b);
        for (int j = 0; j < num_kv; ++j) {
          const int st = j % KS;
          mbar_wait(&kv_empty[st], ((j / KS) & 1) ^ 1);
          mbar_expect_tx(&kv_full[st], 2 * Cfg::KV_TILE);
          tma_load_4d(sK + st * Cfg::KV_TILE, &tmK, &kv_full[st], 0, h, j * BKV, b);
          tma_load_4d(sV + st * Cfg::KV_TILE, &tmV, &kv_full[st], 0, h, j * BKV, b);
        }
      }
    } else if (warp == 1) {
      // ---------------------------- MMA issuer ----------------------------
      constexpr uint32_t idesc_pv = idesc_bf16_f32(T_BQ, T_HD, /*B MN-major*/ 1);
      auto issue_qk = [&](int j) {
        const int st = j % KS;
        mbar_wait(&kv_full[st], (j / KS) & 1);
        tcgen05_fence_after();
        if (elect_one()) {
          const int nvalid = min(BKV, len_kv - j * BKV);
          const int n = max(16, (nvalid + 15) & ~15);
          const uint32_t idesc_qk = idesc_bf16_f32(T_BQ, n, 0);
          const uint64_t adesc = smem_desc_sw128(smem_u32(sQ));
          const uint64_t bdesc = smem_desc_sw128(smem_u32(sK + st * Cfg::KV_TILE));
          const uint32_t s_tmem = tmem_base + Cfg::S_COL + (j & 1) * BKV;
#pragma unroll
          for (int k = 0; k < T_HD / 16; ++k) umma_bf16_ss(s_tmem, adesc + 2 * k, bdesc + 2 * k, idesc_qk, k != 0);
          umma_commit(&s_full[j & 1]);
        }
        __syncwarp();
      };
      mbar_wait(q_full, 0);
      issue_qk(0);
      if (num_kv > 1) issue_qk(1);
      for (int j = 0; j < num_kv; ++j) {
        const int bf = j & 1;
        mbar_wait(&p_ready[bf], (j >> 1) & 1);  // P_j in TMEM, S buffer bf fully read
        tcgen05_fence_after();
        if (elect_one()) {
          const int st = j % KS;
          const int nvalid = min(BKV, len_kv - j * BKV);
          const int nks = (nvalid + 15) >> 4;  // 16 keys per MMA
          const uint32_t v_addr = smem_u32(sV + st * Cfg::KV_TILE);
          const uint32_t p_tmem = tmem_base + Cfg::P_COL + bf * (BKV / 2);
          const uint32_t o_tmem = tmem_base + Cfg::O_COL;
          for (int kk = 0; kk < nks; ++kk)
            umma_bf16_ts(o_tmem, p_tmem + kk * 8, smem_desc_sw128(v_addr + kk * 2048), idesc_pv, (j | kk) != 0);
          umma_commit(&kv_empty[st]);
          umma_commit(&pv_done[bf]);
        }
        __syncwarp();
        if (j + 2 < num_kv) issue_qk(j + 2);  // refill the S buffer that was just drained
      }
    }
  } else {
    // register pool of the CTA is fixed at launch: 384 x 80; warpgroup 0 releases 128 x 40, so the
    // two softmax warpgroups can grow to (384*80 - 128*40) / 256 = 100 -> 96 registers
    setmaxnreg_inc<96>();
    // ----------------------------- softmax ------------------------------
    const int quarter = warp & 3;            // TMEM lane quarter (rows 32*quarter .. +31)
    const int half = (warp - 4) >> 2;        // which 32 of the 64 keys / 32 of the 64 O columns
    const int row_in_tile = quarter * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t o_tmem = tmem_base + lane_off + Cfg::O_COL + half * 32;
    if (q0 + quarter * 32 >= len_q) {
      // every row of this warp is past the end of the sequence: only keep the counts in step
      for (int j = 0; j < num_kv; ++j) {
        if (lane == 0) mbar_arrive(&p_ready[j & 1]);
        mbar_wait(&p_ready[j & 1], (j >> 1) & 1);
      }
    } else {
      float m_used = 0.f, l_run = 0.f;
      constexpr float kRescaleThreshold = 8.f;
      for (int j = 0; j < num_kv; ++j) {
        const int bf = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        const int nvalid = min(BKV, len_kv - j * BKV) - half * 32;  // valid keys in my 32 columns
        const uint32_t s_tmem = tmem_base + lane_off + Cfg::S_COL + bf * BKV + half * 32;
        const uint32_t p_tmem = tmem_base + lane_off + Cfg::P_COL + bf * (BKV / 2) + half * 16;
        mbar_wait(&s_full[bf], ph);
        tcgen05_fence_after();
        uint32_t sr[32];
        if (nvalid > 0) {
          tmem_ld_32x32b_x32(s_tmem, sr);
          tmem_ld_wait();
        }
        if (nvalid < 32) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (i >= nvalid) sr[i] = 0xff800000u;  // -inf
        }
        float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          mx0 = fmaxf(mx0, __uint_as_float(sr[i]));
          mx1 = fmaxf(mx1, __uint_as_float(sr[i + 1]));
          mx2 = fmaxf(mx2, __uint_as_float(sr[i + 2]));
          mx3 = fmaxf(mx3, __uint_as_float(sr[i + 3]));
        }
        const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * scale_log2;
        // agree with the partner warp on whether any row of this quarter needs a new maximum
        const bool want = (j == 0) || (mx > m_used + kRescaleThreshold);
        const int my_any = __any_sync(0xffffffffu, want) ? 1 : 0;
        // (flag slots alternate with the iteration parity: a slot is rewritten two barriers later)
        if (lane == 0) pair_flag[(j & 1) * 8 + quarter * 2 + half] = my_any;
        pair_sync(quarter);
        const int any = my_any | pair_flag[(j & 1) * 8 + quarter * 2 + (half ^ 1)];
        if (any) {
          // rare (always at j == 0): exchange the per-row maxima so both halves of a row use the
          // same reference, then bring O and l to it.  PV_{j-1} must be done; PV_j is not issued
          // before both warps arrive on p_ready below.
          xch[row_in_tile * 2 + half] = mx;
          pair_sync(quarter);
          const float m_row = fmaxf(mx, xch[row_in_tile * 2 + (half ^ 1)]);
          if (j == 0) {
            m_used = m_row;
          } else {
            const bool grow = m_row > m_used + kRescaleThreshold;
            mbar_wait(&pv_done[(j - 1) & 1], ((j - 1) >> 1) & 1);
            tcgen05_fence_after();
            const float alpha = grow ? ex2_approx(m_used - m_row) : 1.f;
            uint32_t r[32];
            tmem_ld_32x32b_x32(o_tmem, r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
            tmem_st_32x32b_x32(o_tmem, r);
            tmem_st_wait();
            l_run *= alpha;
            if (grow) m_used = m_row;
          }
          // (the flag / xch slots are rewritten only after the next bar.sync of this pair)
        }
        // P buffer bf was last read by PV_{j-2}
        if (j >= 2) {
          mbar_wait(&pv_done[bf], ((j - 2) >> 1) & 1);
          tcgen05_fence_after();
        }
        if (nvalid > 0) {
          float s0 = 0.f, s1 = 0.f;
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float p0 = ex2_approx(fmaf(__uint_as_float(sr[i]), scale_log2, -m_used));
            const float p1 = ex2_approx(fmaf(__uint_as_float(sr[i + 1]), scale_log2, -m_used));
            s0 += p0;
            s1 += p1;
            pk[i >> 1] = pack_bf16x2(p0, p1);
          }
          l_run += s0 + s1;
          tmem_st_32x32b_x16(p_tmem, pk);
          tmem_st_wait();
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_ready[bf]);
      }
      // epilogue: O / (l_mine + l_partner), 32 output columns per thread
      xch[row_in_tile * 2 + half] = l_run;
      pair_sync(quarter);
      const float inv = 1.f / (l_run + xch[row_in_tile * 2 + (half ^ 1)]);
      mbar_wait(&pv_done[(num_kv - 1) & 1], ((num_kv - 1) >> 1) & 1);
      tcgen05_fence_after();
      const int row = q0 + row_in_tile;
      uint16_t* orow = out + b * o_bs + (int64_t)row * o_ls + h * T_HD + half * 32;
      uint32_t r[32];
      tmem_ld_32x32b_x32(o_tmem, r);
      tmem_ld_wait();
      if (row < len_q) {
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          float v[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) v[u] = __uint_as_float(r[i + u]) * inv;
          *reinterpret_cast<uint4*>(orow + i) =
              make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
        }
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

static int launch_attn_tc4(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, uint16_t* out,
                           int64_t o_bs, int64_t o_ls, int batch, int heads, int len_q, int len_kv,
                           float scale_log2, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_bf16_tc4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Attn4Cfg::SMEM_BYTES);
    if (e != cudaSuccess) {
      set_error("attention_bf16: cudaFuncSetAttribute(%d): %s", Attn4Cfg::SMEM_BYTES, cudaGetErrorString(e));
      return PCD_ERR_CUDA;
    }
    attr_set = true;
  }
  dim3 grid(ceil_div(len_q, T_BQ), heads, batch);
  attn_bf16_tc4_kernel<<<grid, 384, Attn4Cfg::SMEM_BYTES, st>>>(tq, tk, tv, out, o_bs, o_ls, len_q, len_kv, scale_log2);
  PCD_CHECK_LAUNCH("attention_bf16");
  return PCD_OK;
}

template <bool P_TMEM>
static int launch_attn_tc(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv,
                          uint16_t* out, int64_t o_bs, int64_t o_ls, int batch, int heads, int len_q,
                          int len_kv, float scale_log2, cudaStream_t st) {
  using Cfg = AttnCfg<P_TMEM>;
  auto kern = attn_bf16_tc_kernel<P_TMEM>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) {
      set_error("attention_bf16: cudaFuncSetAttribute(%d): %s", Cfg::SMEM_BYTES, cudaGetErrorString(e));
      return PCD_ERR_CUDA;
    }
    attr_set = true;
  }
  dim3 grid(ceil_div(len_q, T_BQ), heads, batch);
  kern<<<grid, 256, Cfg::SMEM_BYTES, st>>>(tq, tk, tv, out, o_bs, o_ls, len_q, len_kv, scale_log2);
  PCD_CHECK_LAUNCH("attention_bf16");
  return PCD_OK;
}

static int make_operand_map(CUtensorMap* m, const pcd_attn_operand* op, int batch, int heads, int len, int box_rows) {
  uint64_t dims[4] = {T_HD, (uint64_t)heads, (uint64_t)len, (uint64_t)batch};
  uint64_t strides[3] = {(uint64_t)op->head_stride * 2, (uint64_t)op->row_stride * 2, (uint64_t)op->batch_stride * 2};
  uint32_t box[4] = {T_HD, 1, (uint32_t)box_rows, 1};
  return encode_tmap_bf16(m, op->ptr, 4, dims, strides, box);
}

int launch_attn_tc5(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, uint16_t* out,
                    int64_t o_bs, int64_t o_ls, int batch, int heads, int len_q, int len_kv, float scale_log2,
                    int poly, cudaStream_t st);  // attn_tc5.cu

int g_attn_variant = 5;  // 5: persistent + software-pipelined softmax (default, fastest measured); 6, 7: as 5 with 1/4, 1/2 of the
                         // exponentials evaluated by an FMA-pipe polynomial; 3: double-buffered S/P, 64-key tiles, 2 CTAs/SM;
                         // 4: as 3 with rows split over two softmax threads; 2: ping-pong
                         // over two query tiles; 1: one tile per CTA, P in
                         // TMEM, 2 CTAs/SM; 0: one tile per CTA, P in shared memory (SS MMA)

int launch_attention_bf16(const pcd_attn_operand* q, const pcd_attn_operand* k,
                          const pcd_attn_operand* v, uint16_t* out, int64_t o_bs, int64_t o_ls,
                          int batch, int heads, int len_q, int len_kv, float q_scale, float k_scale,
                          cudaStream_t st) {
  CUtensorMap tq, tk, tv;
  int rc;
  const int kv_rows = g_attn_variant >= 3 ? Attn3Cfg::BKV : T_BKV;
  if ((rc = make_operand_map(&tq, q, batch, heads, len_q, T_BQ)) != PCD_OK) return rc;
  if ((rc = make_operand_map(&tk, k, batch, heads, len_kv, kv_rows)) != PCD_OK) return rc;
  if ((rc = make_operand_map(&tv, v, batch, heads, len_kv, kv_rows)) != PCD_OK) return rc;
  const float scale_log2 = q_scale * k_scale * 1.4426950408889634f;
  if (g_attn_variant >= 5)  // 5: all exponentials on the MUFU; 6: 1/4, 7: 1/2 of them on the FMA pipes
    return launch_attn_tc5(tq, tk, tv, out, o_bs, o_ls, batch, heads, len_q, len_kv, scale_log2,
                           g_attn_variant == 6 ? 2 : (g_attn_variant == 7 ? 1 : 0), st);
  if (g_attn_variant == 4)
    return launch_attn_tc4(tq, tk, tv, out, o_bs, o_ls, batch, heads, len_q, len_kv, scale_log2, st);
  if (g_attn_variant == 3)
    return launch_attn_tc3(tq, tk, tv, out, o_bs, o_ls, batch, heads, len_q, len_kv, scale_log2, st);
  if (g_attn_variant == 2)
    return launch_attn_tc2(tq, tk, tv, out, o_bs, o_ls, batch, heads, len_q, len_kv, scale_log2, st);
  if (g_attn_variant == 1)
    return launch_attn_tc<true>(tq, tk, tv, out, o_bs, o_ls, batch, heads, len_q, len_kv, scale_log2, st);
  return launch_attn_tc<false>(tq, tk, tv, out, o_bs, o_ls, batch, heads, len_q, len_kv, scale_log2, st);
}

}  // namespace pcd
