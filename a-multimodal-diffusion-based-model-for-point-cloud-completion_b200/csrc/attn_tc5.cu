// bf16 flash attention (head dim 64, non-causal) on tcgen05 -- persistent, software-pipelined.
//
// Same tiling as variant 3 (128-query tile x 64-key tiles, S and P double-buffered in TMEM, O
// resident in TMEM with lazy rescaling, two CTAs per SM), restructured around what the ncu
// capture of variant 3 showed (profiles/r01/prof_attn_v4.ncu-rep: the four softmax warps of a CTA
// issue only 22 % of the time; a third of every iteration is spent waiting for S_j to arrive and
// for the TMEM load round trip, during which the MUFU pipe -- the binding unit at head dim 64 --
// idles):
//   * PERSISTENT CTAs (2 per SM) walk a static list of (query tile, head, sequence) items; the
//     TMA warps and the MMA thread run ahead across item boundaries (Q double-buffered), so the
//     prologue (barrier init, TMEM allocation, first Q/K round trip) is paid once per CTA and the
//     epilogue of item i overlaps the first QK^T products of item i+1;
//   * the softmax warps are SOFTWARE-PIPELINED over KV tiles: the tcgen05.ld of S_{j+1} is issued
//     before the exponentials of S_j, and the S buffer is handed back to the MMA thread (s_free)
//     as soon as it sits in registers, one exponential phase earlier than P_j -- so QK_{j+2} is
//     never on the softmax critical path;
//   * K and V travel through separate rings fed by two producer warps (K is consumed two tiles
//     ahead of V);
//   * the scale/subtract and the row sum use the packed fma.rn.f32x2 / add.rn.f32x2 forms
//     (half the FMA-pipe issue slots), shared-memory barrier addresses are 32-bit values computed
//     once (variant 3 re-derived them with S2UR every iteration).
// Warp roles (256 threads): 0 = Q/K TMA producer, 1 = QK^T issuer, 2 = V TMA producer, 3 = PV issuer,
// 4-7 = softmax (thread = query row).
#include "common.cuh"
#include "tc_sm100.cuh"
#include <stdlib.h>

namespace pcd {

using namespace tc;

namespace a5 {

constexpr int BQ = 128, BKV = 64, HD = 64;
constexpr int Q_TILE = BQ * HD * 2;     // 16 KB
constexpr int KV_TILE = BKV * HD * 2;   // 8 KB
constexpr int KSK = 5, KSV = 4;
constexpr int TILE_BYTES = 2 * Q_TILE + (KSK + KSV) * KV_TILE;  // 104 KB
constexpr int SMEM_BYTES = TILE_BYTES + 1024 + 512;
constexpr int TMEM_COLS = 256;
constexpr int S_COL = 0, P_COL = 128, O_COL = 192;  // S: +buf*64, P: +buf*32

// barrier slots (8 bytes each, relative to the barrier block)
constexpr int B_Q_FULL = 0, B_Q_EMPTY = 2, B_K_FULL = 4, B_K_EMPTY = B_K_FULL + KSK, B_V_FULL = B_K_EMPTY + KSK,
              B_V_EMPTY = B_V_FULL + KSV, B_S_FULL = B_V_EMPTY + KSV, B_S_FREE = B_S_FULL + 2,
              B_P_READY = B_S_FREE + 2, B_PV_DONE = B_P_READY + 2, B_O_FREE = B_PV_DONE + 2, B_COUNT = B_O_FREE + 1;
static_assert(B_COUNT * 8 + 16 <= 512, "barrier block too small");

// ---- 32-bit shared-address forms of the mbarrier / TMA / commit wrappers ----
__device__ __forceinline__ void bar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void bar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void bar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "LAB_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra LAB_WAIT;\n\t"
      "DONE:\n\t"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
// one non-blocking probe: issued early, its ~90-cycle round trip overlaps whatever is scheduled before the
// result is consumed
__device__ __forceinline__ uint32_t bar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P1;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}
__device__ __forceinline__ void tma4(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// 2^x for a pair, x <= ~8, entirely on the FMA / ALU pipes (no MUFU): round-to-nearest split
// x = n + f, |f| <= 0.5 (magic-number add), degree-3 minimax polynomial for 2^f (max rel. error
// 7.5e-5, far below the bf16 rounding of P), exponent patched in with an integer shift-add.  The
// clamp at -126 turns everything below (incl. the -inf of masked keys) into denormals (n = -127 would
// wrap the exponent field of a polynomial value below 1.0 into the sign bit).
__device__ __forceinline__ void exp2_poly2(uint64_t x2, float& p0, float& p1) {
  float x0, x1;
  unpack2(x2, x0, x1);
  x0 = fmaxf(x0, -126.f);
  x1 = fmaxf(x1, -126.f);
  const uint64_t xc = pack2(x0, x1);
  const uint64_t magic = pack2(12582912.f, 12582912.f);  // 1.5 * 2^23
  const uint64_t r2 = add2(xc, magic);                    // n in the low mantissa bits
  const uint64_t fl = add2(r2, pack2(-12582912.f, -12582912.f));
  const uint64_t f2 = fma2(fl, pack2(-1.f, -1.f), xc);
  uint64_t q = fma2(pack2(0.0551716685f, 0.0551716685f), f2, pack2(0.24261114f, 0.24261114f));
  q = fma2(q, f2, pack2(0.693260968f, 0.693260968f));
  q = fma2(q, f2, pack2(0.999928057f, 0.999928057f));
  float q0, q1, r0, r1;
  unpack2(q, q0, q1);
  unpack2(r2, r0, r1);
  p0 = __uint_as_float(__float_as_uint(q0) + (__float_as_uint(r0) << 23));
  p1 = __uint_as_float(__float_as_uint(q1) + (__float_as_uint(r1) << 23));
}

// keep a value in its register: stops ptxas from re-deriving it (S2R / cvta chains) at every use
__device__ __forceinline__ void pin(uint32_t& v) { asm volatile("" : "+r"(v)); }

// decode a work item: consecutive items share (sequence, head) so that their K/V stay in L2
struct Item {
  int q0, h, b;
};
__device__ __forceinline__ Item decode_item(int item, int nq, int heads) {
  const int qt = item % nq, bh = item / nq;
  return Item{qt * BQ, bh % heads, bh / heads};
}

// ------------------------------ softmax warp state ------------------------------
struct Sm {
  uint32_t bars;     // shared address of the barrier block
  uint32_t tm;       // TMEM base + this warp's lane quarter
  uint32_t u0, u1;   // parity of the number of completed uses of S/P buffer 0 / 1
  uint32_t lane;
  float m_used;      // maximum currently baked into P, l and O (log2 domain)
  float l_run;       // running row sum
  float scale;       // softmax scale * log2(e)
};
constexpr float kRescaleThreshold = 8.f;  // log2 units: P <= 2^8, safe in bf16 / fp32

// P = exp2(S * scale - m_used) for the first `ncols` (32 or 64) columns of cur -> bf16 in TMEM
template <int POLY>  // POLY: every POLY-th group of four columns takes two of its exponentials off the MUFU (0 = none)
__device__ __forceinline__ float exp_tile(const Sm& c, const uint32_t (&cur)[BKV], uint32_t p_tmem, int ncols) {
  const uint64_t sc2 = pack2(c.scale, c.scale);
  const uint64_t nm2 = pack2(-c.m_used, -c.m_used);
  uint64_t sum_a = 0ull, sum_b = 0ull;  // (0.f, 0.f)
#pragma unroll
  for (int h = 0; h < BKV / 32; ++h) {
    if (h * 32 < ncols) {
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        const uint64_t xa = fma2(pack2(__uint_as_float(cur[h * 32 + i]), __uint_as_float(cur[h * 32 + i + 1])), sc2, nm2);
        const uint64_t xb = fma2(pack2(__uint_as_float(cur[h * 32 + i + 2]), __uint_as_float(cur[h * 32 + i + 3])), sc2, nm2);
        float a0, a1, b0, b1;
        unpack2(xa, a0, a1);
        unpack2(xb, b0, b1);
        const float p0 = ex2_approx(a0), p1 = ex2_approx(a1);
        float p2, p3;
        if (POLY > 0 && ((i >> 2) % (POLY > 0 ? POLY : 1)) == POLY - 1) {
          exp2_poly2(xb, p2, p3);
        } else {
          p2 = ex2_approx(b0);
          p3 = ex2_approx(b1);
        }
        sum_a = add2(sum_a, pack2(p0, p1));
        sum_b = add2(sum_b, pack2(p2, p3));
        pk[i >> 1] = pack_bf16x2(p0, p1);
        pk[(i >> 1) + 1] = pack_bf16x2(p2, p3);
      }
      tmem_st_32x32b_x16(p_tmem + h * 16, pk);
    }
  }
  float s0, s1, s2, s3;
  unpack2(sum_a, s0, s1);
  unpack2(sum_b, s2, s3);
  return (s0 + s1) + (s2 + s3);
}

// the two 32-column halves of exp_tile as separate calls (sums carried by the caller), so that barrier probes and
// the next tile's TMEM load can sit between them
template <int POLY>
__device__ __forceinline__ void exp_half32(const Sm& c, const uint32_t (&cur)[BKV], int h, uint32_t (&pk)[16], uint64_t& sum_a,
                                           uint64_t& sum_b) {
  const uint64_t sc2 = pack2(c.scale, c.scale);
  const uint64_t nm2 = pack2(-c.m_used, -c.m_used);
#pragma unroll
  for (int i = 0; i < 32; i += 4) {
    const uint64_t xa = fma2(pack2(__uint_as_float(cur[h * 32 + i]), __uint_as_float(cur[h * 32 + i + 1])), sc2, nm2);
    const uint64_t xb = fma2(pack2(__uint_as_float(cur[h * 32 + i + 2]), __uint_as_float(cur[h * 32 + i + 3])), sc2, nm2);
    float a0, a1, b0, b1;
    unpack2(xa, a0, a1);
    unpack2(xb, b0, b1);
    const float p0 = ex2_approx(a0), p1 = ex2_approx(a1);
    float p2, p3;
    if (POLY > 0 && ((i >> 2) % (POLY > 0 ? POLY : 1)) == POLY - 1) {
      exp2_poly2(xb, p2, p3);
    } else {
      p2 = ex2_approx(b0);
      p3 = ex2_approx(b1);
    }
    sum_a = add2(sum_a, pack2(p0, p1));
    sum_b = add2(sum_b, pack2(p2, p3));
    pk[i >> 1] = pack_bf16x2(p0, p1);
    pk[(i >> 1) + 1] = pack_bf16x2(p2, p3);
  }
}

// row maximum (log2 domain); keys >= nvalid of a ragged tile are masked to -inf first
__device__ __forceinline__ float row_max(const Sm& c, uint32_t (&r)[BKV], int nvalid) {
  if (nvalid < BKV) {
#pragma unroll
    for (int i = 0; i < BKV; ++i)
      if (i >= nvalid) r[i] = 0xff800000u;  // -inf
  }
  float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
  for (int i = 0; i < BKV; i += 4) {
    mx0 = fmaxf(mx0, __uint_as_float(r[i]));
    mx1 = fmaxf(mx1, __uint_as_float(r[i + 1]));
    mx2 = fmaxf(mx2, __uint_as_float(r[i + 2]));
    mx3 = fmaxf(mx3, __uint_as_float(r[i + 3]));
  }
  return fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * c.scale;
}

// pipelined step on the tile held in cur (S/P buffer X): the load of the next tile (buffer X^1)
// is in flight while cur is exponentiated.  nvalid_next < 64 only for a ragged last tile.
// (A variant that skipped the per-tile row maximum and re-referenced only when a tile's row sum
// exceeded 2^8 measured 6-17 % slower: the vote then sits between the exponentials and the P
// hand-off instead of overlapping the next tile's maximum with them.)
template <int X, int POLY>
__device__ __forceinline__ void pstep(Sm& c, uint32_t (&cur)[BKV], uint32_t (&nxt)[BKV], int nvalid_next) {
  constexpr int Y = X ^ 1;
  uint32_t& ux = X ? c.u1 : c.u0;
  uint32_t& uy = X ? c.u0 : c.u1;
  // Both barriers of the step are normally complete already, but a try_wait still takes ~90 cycles to say so:
  // probe them now and consume the answers after the first half of the exponentials.
  const uint32_t ok_s = bar_try(c.bars + 8 * (B_S_FULL + Y), uy);
  const uint32_t ok_pv = bar_try(c.bars + 8 * (B_PV_DONE + X), ux ^ 1);
  uint64_t sum_a = 0ull, sum_b = 0ull;
  uint32_t pk[16];
  exp_half32<POLY>(c, cur, 0, pk, sum_a, sum_b);
  // S of the next tile -> registers (not waited for yet)
  if (!ok_s) bar_wait(c.bars + 8 * (B_S_FULL + Y), uy);
  tcgen05_fence_after();
  tmem_ld_32x32b_x32(c.tm + S_COL + Y * BKV, nxt);
  tmem_ld_32x32b_x32(c.tm + S_COL + Y * BKV + 32, nxt + 32);
  // P buffer X is free once the PV product of its previous use has completed
  if (!ok_pv) bar_wait(c.bars + 8 * (B_PV_DONE + X), ux ^ 1);
  tcgen05_fence_after();
  tmem_st_32x32b_x16(c.tm + P_COL + X * (BKV / 2), pk);
  exp_half32<POLY>(c, cur, 1, pk, sum_a, sum_b);
  tmem_st_32x32b_x16(c.tm + P_COL + X * (BKV / 2) + 16, pk);
  {
    float s0, s1, s2, s3;
    unpack2(sum_a, s0, s1);
    unpack2(sum_b, s2, s3);
    c.l_run += (s0 + s1) + (s2 + s3);
  }
  tmem_st_wait();
  tmem_ld_wait();
  tcgen05_fence_before();
  __syncwarp();
  if (c.lane == 0) {
    bar_arrive(c.bars + 8 * (B_P_READY + X));  // P in TMEM -> PV may be issued
    bar_arrive(c.bars + 8 * (B_S_FREE + Y));   // next S tile sits in registers -> its buffer is reusable
  }
  // row maximum of the next tile; lazy rescale of O and l when it grows by more than 2^8
  const float mx = row_max(c, nxt, nvalid_next);
  const bool grow = mx > c.m_used + kRescaleThreshold;
  if (__any_sync(0xffffffffu, grow)) {
    // rare.  PV of the current tile must be complete; the next PV is not issued before this warp
    // arrives on p_ready in the next step.
    bar_wait(c.bars + 8 * (B_PV_DONE + X), ux);
    tcgen05_fence_after();
    const float alpha = grow ? ex2_approx(c.m_used - mx) : 1.f;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      uint32_t r[32];
      tmem_ld_32x32b_x32(c.tm + O_COL + h * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
      tmem_st_32x32b_x32(c.tm + O_COL + h * 32, r);
    }
    tmem_st_wait();
    c.l_run *= alpha;
    if (grow) c.m_used = mx;
  }
  ux ^= 1;
}

// last tile of an item (buffer X, `nvalid` real keys): exponentiate, then O / l -> global
template <int X, int POLY>
__device__ __forceinline__ void last_step(Sm& c, uint32_t (&cur)[BKV], int nvalid, uint16_t* __restrict__ orow,
                                          bool row_ok) {
  uint32_t& ux = X ? c.u1 : c.u0;
  bar_wait(c.bars + 8 * (B_PV_DONE + X), ux ^ 1);
  tcgen05_fence_after();
  c.l_run += exp_tile<POLY>(c, cur, c.tm + P_COL + X * (BKV / 2), nvalid);
  tmem_st_wait();
  tcgen05_fence_before();
  __syncwarp();
  if (c.lane == 0) bar_arrive(c.bars + 8 * (B_P_READY + X));
  bar_wait(c.bars + 8 * (B_PV_DONE + X), ux);  // every PV product of the item has landed in O
  tcgen05_fence_after();
  tmem_ld_32x32b_x32(c.tm + O_COL, cur);
  tmem_ld_32x32b_x32(c.tm + O_COL + 32, cur + 32);
  tmem_ld_wait();
  tcgen05_fence_before();
  __syncwarp();
  if (c.lane == 0) bar_arrive(c.bars + 8 * B_O_FREE);  // O is in registers: the next item may overwrite it
  const float inv = 1.f / c.l_run;
  if (row_ok) {
#pragma unroll
    for (int i = 0; i < 64; i += 8) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = __uint_as_float(cur[i + u]) * inv;
      *reinterpret_cast<uint4*>(orow + i) =
          make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
    }
  }
  ux ^= 1;
}

template <int POLY>
__global__ void __launch_bounds__(256, 2)
attn_bf16_tc5_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, uint16_t* __restrict__ out, int64_t o_bs,
                     int64_t o_ls, int len_q, int len_kv, float scale_log2, int nq, int heads, int n_items) {
  extern __shared__ unsigned char smem_raw[];
  uint32_t smem = (smem_u32(smem_raw) + 1023u) & ~1023u;
  pin(smem);
  const uint32_t sQ = smem;                       // [2] 16 KB
  const uint32_t sK = sQ + 2 * Q_TILE;            // [KSK] 8 KB
  const uint32_t sV = sK + KSK * KV_TILE;         // [KSV] 8 KB
  const uint32_t bars = smem + TILE_BYTES;
  auto bar = [&](int slot) -> uint32_t { return bars + 8u * slot; };
  const uint32_t tmem_slot = bars + 8 * B_COUNT;

  uint32_t tid = threadIdx.x;
  pin(tid);
  const int warp = tid >> 5, lane = tid & 31;
  const int num_kv = (len_kv + BKV - 1) / BKV;
  const int last_valid = len_kv - (num_kv - 1) * BKV;  // keys in the last KV tile (1..64)
  // items of this CTA: blockIdx.x, blockIdx.x + gridDim.x, ...
  const int my_items = (n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmQ);
    prefetch_tensormap(&tmK);
    prefetch_tensormap(&tmV);
    for (int s = 0; s < 2; ++s) {
      bar_init(bar(B_Q_FULL + s), 1);
      bar_init(bar(B_Q_EMPTY + s), 1);
      bar_init(bar(B_S_FULL + s), 1);
      bar_init(bar(B_S_FREE + s), 4);   // one arrival per softmax warp
      bar_init(bar(B_P_READY + s), 4);
      bar_init(bar(B_PV_DONE + s), 1);
    }
    for (int s = 0; s < KSK; ++s) {
      bar_init(bar(B_K_FULL + s), 1);
      bar_init(bar(B_K_EMPTY + s), 1);
    }
    for (int s = 0; s < KSV; ++s) {
      bar_init(bar(B_V_FULL + s), 1);
      bar_init(bar(B_V_EMPTY + s), 1);
    }
    bar_init(bar(B_O_FREE), 4);
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS)
                 : "memory");
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp < 4) {
    setmaxnreg_dec<40>();
    if (warp == 0) {
      // ----------------------- Q / K producer -----------------------
      if (elect_one()) {
        int st = 0;
        uint32_t ph = 1;  // "empty" barriers: the first pass over the ring does not block
        for (int k = 0; k < my_items; ++k) {
          const Item it = decode_item(blockIdx.x + k * gridDim.x, nq, heads);
          const int qb = k & 1;
          bar_wait(bar(B_Q_EMPTY + qb), ((k >> 1) & 1) ^ 1);
          bar_expect_tx(bar(B_Q_FULL + qb), Q_TILE);
          tma4(sQ + qb * Q_TILE, &tmQ, bar(B_Q_FULL + qb), 0, it.h, it.q0, it.b);
          for (int j = 0; j < num_kv; ++j) {
            bar_wait(bar(B_K_EMPTY + st), ph);
            bar_expect_tx(bar(B_K_FULL + st), KV_TILE);
            tma4(sK + st * KV_TILE, &tmK, bar(B_K_FULL + st), 0, it.h, j * BKV, it.b);
            if (++st == KSK) {
              st = 0;
              ph ^= 1;
            }
          }
        }
      }
    } else if (warp == 2) {
      // ------------------------- V producer -------------------------
      if (elect_one()) {
        int st = 0;
        uint32_t ph = 1;
        for (int k = 0; k < my_items; ++k) {
          const Item it = decode_item(blockIdx.x + k * gridDim.x, nq, heads);
          for (int j = 0; j < num_kv; ++j) {
            bar_wait(bar(B_V_EMPTY + st), ph);
            bar_expect_tx(bar(B_V_FULL + st), KV_TILE);
            tma4(sV + st * KV_TILE, &tmV, bar(B_V_FULL + st), 0, it.h, j * BKV, it.b);
            if (++st == KSV) {
              st = 0;
              ph ^= 1;
            }
          }
        }
      }
    } else if (warp == 1) {
      // ------------------------ QK^T issuer -------------------------
      // Walks the CTA's flat stream of KV tiles.  S/P buffer of a tile = its index within the item
      // & 1; S = Q K^T for a tile is issued as soon as the previous S tile of that buffer sits in
      // the softmax registers (s_free) -- up to three tiles ahead of the exponentials.  The PV
      // products have their own issuer (warp 3) so that neither stream ever waits for the other's
      // barriers (a single in-order issuer either delays QK^T behind P_j or deadlocks on short
      // sequences).
      if (elect_one()) {
        constexpr uint32_t idesc_qk_full = idesc_bf16_f32(BQ, BKV, 0);
        const uint32_t idesc_qk_last = idesc_bf16_f32(BQ, max(16, (last_valid + 15) & ~15), 0);
        int st = 0;
        uint32_t ph = 0, u0 = 0, u1 = 0;  // K ring phase; completed uses of S buffer 0 / 1 (parity)
        for (int k = 0; k < my_items; ++k) {
          const int qb = k & 1;
          const uint64_t adesc = smem_desc_sw128(sQ + qb * Q_TILE);
          bar_wait(bar(B_Q_FULL + qb), (k >> 1) & 1);
          for (int j = 0; j < num_kv; ++j) {
            const int b = j & 1;
            bar_wait(bar(B_K_FULL + st), ph);             // (loaded long ago: off the critical path)
            bar_wait(bar(B_S_FREE + b), (b ? u1 : u0) ^ 1);  // passes at once for the first use of a buffer
            if (b) u1 ^= 1; else u0 ^= 1;
            tcgen05_fence_after();
            const bool last = (j == num_kv - 1);
            const uint32_t idesc = last ? idesc_qk_last : idesc_qk_full;
            const uint64_t bdesc = smem_desc_sw128(sK + st * KV_TILE);
            const uint32_t s_tmem = tmem_base + S_COL + b * BKV;
#pragma unroll
            for (int kk = 0; kk < HD / 16; ++kk) umma_bf16_ss(s_tmem, adesc + 2 * kk, bdesc + 2 * kk, idesc, kk != 0);
            commit(bar(B_S_FULL + b));
            commit(bar(B_K_EMPTY + st));
            if (last) commit(bar(B_Q_EMPTY + qb));
            if (++st == KSK) {
              st = 0;
              ph ^= 1;
            }
          }
        }
      }
    } else {
      // -------------------------- PV issuer -------------------------
      if (elect_one()) {
        constexpr uint32_t idesc_pv = idesc_bf16_f32(BQ, HD, /*B MN-major*/ 1);
        const int nks_last = (last_valid + 15) >> 4;
        const uint32_t o_tmem = tmem_base + O_COL;
        int st = 0;
        uint32_t ph = 0, u0 = 0, u1 = 0;
        for (int k = 0; k < my_items; ++k) {
          for (int j = 0; j < num_kv; ++j) {
            const int b = j & 1;
            bar_wait(bar(B_V_FULL + st), ph);
            if (j == 0 && k > 0) bar_wait(bar(B_O_FREE), (k - 1) & 1);  // previous item's O read out
            bar_wait(bar(B_P_READY + b), b ? u1 : u0);                  // P of this tile is in TMEM
            if (b) u1 ^= 1; else u0 ^= 1;
            tcgen05_fence_after();
            const int nks = (j == num_kv - 1) ? nks_last : BKV / 16;
            const uint32_t v_addr = sV + st * KV_TILE;
            const uint32_t p_tmem = tmem_base + P_COL + b * (BKV / 2);
            for (int kk = 0; kk < nks; ++kk)
              umma_bf16_ts(o_tmem, p_tmem + kk * 8, smem_desc_sw128(v_addr + kk * 2048), idesc_pv, (j | kk) != 0);
            commit(bar(B_PV_DONE + b));
            commit(bar(B_V_EMPTY + st));
            if (++st == KSV) {
              st = 0;
              ph ^= 1;
            }
          }
        }
      }
    }
  } else {
    setmaxnreg_inc<216>();
    // --------------------------- softmax ----------------------------
    const int quarter = warp & 3;
    Sm c;
    c.bars = bars;
    c.tm = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    c.u0 = c.u1 = 0;
    c.lane = lane;
    c.m_used = 0.f;
    c.l_run = 0.f;
    c.scale = scale_log2;
    pin(c.bars);
    pin(c.tm);
    pin(c.lane);
    const bool partial = last_valid < BKV;

    for (int k = 0; k < my_items; ++k) {
      const Item it = decode_item(blockIdx.x + k * gridDim.x, nq, heads);
      if (it.q0 + quarter * 32 >= len_q) {
        // no real query row in this warp (ragged last query tile): keep the barrier counts in step
        for (int j = 0; j < num_kv; ++j) {
          const int b = j & 1;
          uint32_t& u = b ? c.u1 : c.u0;
          if (c.lane == 0) {
            bar_arrive(c.bars + 8 * (B_S_FREE + b));
            bar_arrive(c.bars + 8 * (B_P_READY + b));
          }
          bar_wait(c.bars + 8 * (B_P_READY + b), u);  // all four warps are past this tile
          u ^= 1;
        }
        if (c.lane == 0) bar_arrive(c.bars + 8 * B_O_FREE);
        continue;
      }
      uint32_t ra[BKV], rb[BKV];
      // first tile of the item -> ra
      bar_wait(c.bars + 8 * (B_S_FULL + 0), c.u0);
      tcgen05_fence_after();
      tmem_ld_32x32b_x32(c.tm + S_COL, ra);
      tmem_ld_32x32b_x32(c.tm + S_COL + 32, ra + 32);
      tmem_ld_wait();
      tcgen05_fence_before();
      __syncwarp();
      if (c.lane == 0) bar_arrive(c.bars + 8 * (B_S_FREE + 0));
      c.m_used = row_max(c, ra, (num_kv == 1) ? last_valid : BKV);
      c.l_run = 0.f;
      const int row = it.q0 + quarter * 32 + lane;
      uint16_t* orow = out + it.b * o_bs + (int64_t)row * o_ls + it.h * HD;
      const bool row_ok = row < len_q;
      int j = 0;
      bool done = false;
      for (; j + 1 < num_kv; j += 2) {
        pstep<0, POLY>(c, ra, rb, (partial && j + 2 == num_kv) ? last_valid : BKV);
        if (j + 2 < num_kv) {
          pstep<1, POLY>(c, rb, ra, (partial && j + 3 == num_kv) ? last_valid : BKV);
        } else {
          last_step<1, POLY>(c, rb, last_valid, orow, row_ok);
          done = true;
        }
      }
      if (!done) last_step<0, POLY>(c, ra, last_valid, orow, row_ok);
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace a5

template <int POLY>
static int launch_tc5(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, uint16_t* out,
                      int64_t o_bs, int64_t o_ls, int batch, int heads, int len_q, int len_kv, float scale_log2,
                      cudaStream_t st) {
  auto kern = a5::attn_bf16_tc5_kernel<POLY>;
  {  // per device, not per process
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, a5::SMEM_BYTES);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) {
      set_error("attention_bf16: kernel attribute setup (%d B smem): %s", a5::SMEM_BYTES, cudaGetErrorString(e));
      return PCD_ERR_CUDA;
    }
  }
  const int nq = ceil_div(len_q, a5::BQ);
  const int64_t n_items64 = (int64_t)nq * heads * batch;
  if (n_items64 > 0x7fffffff) {
    set_error("attention_bf16: too many (query tile, head, sequence) items");
    return PCD_ERR_INVALID;
  }
  const int n_items = (int)n_items64;
  // two CTAs per SM by construction: 2 x (104.5 KB smem + 1 KB reserved) <= 228 KB, 2 x 256 threads x 128
  // registers = the whole file, 2 x 256 TMEM columns = all 512 (ncu: launch__occupancy_limit_* = 2)
  const int grid = min(n_items, 2 * num_sms());
  kern<<<grid, 256, a5::SMEM_BYTES, st>>>(tq, tk, tv, out, o_bs, o_ls, len_q, len_kv, scale_log2, nq, heads, n_items);
  PCD_CHECK_LAUNCH("attention_bf16");
  return PCD_OK;
}

// poly: 0 = every exponential on the MUFU, 2 = 1/4 and 1 = 1/2 of them on the FMA pipes
int launch_attn_tc5(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, uint16_t* out,
                    int64_t o_bs, int64_t o_ls, int batch, int heads, int len_q, int len_kv, float scale_log2,
                    int poly, cudaStream_t st) {
  if (poly == 2) return launch_tc5<2>(tq, tk, tv, out, o_bs, o_ls, batch, heads, len_q, len_kv, scale_log2, st);
  if (poly == 1) return launch_tc5<1>(tq, tk, tv, out, o_bs, o_ls, batch, heads, len_q, len_kv, scale_log2, st);
  return launch_tc5<0>(tq, tk, tv, out, o_bs, o_ls, batch, heads, len_q, len_kv, scale_log2, st);
}

}  // namespace pcd
