// bf16 flash attention (head dim 64, non-causal) on tcgen05 -- three query tiles per CTA ("grouped" kernel).
//
// What bounds head-dim-64 attention on this chip is the special-function unit: one ex2 per logit against 256
// tensor FLOP, 16 ex2/clk/SM = 8 cycles of the pipe per warp instruction.  The paired kernel (attn_tc5.cu: two CTAs
// per SM, one 128-query tile each, i.e. two softmax warps per scheduler) left that unit 43 % idle
// (profiles/r01/prof_attn_v5.ncu-rep).  A warp's step is a chain -- S from TMEM, row maximum, 64 exponentials,
// P to TMEM, the barrier round trips to the two MMA issuers -- of ~1600 cycles of which it holds the pipe for 512;
// two such warps cannot keep it busy however they interleave.  This kernel gives every scheduler THREE:
//   * ONE CTA per SM with three query tiles ("slots") of the same (sequence, head): 12 softmax warps, three per
//     scheduler, sharing every K / V tile (one TMA load and one shared-memory copy serve three S = Q K^T products);
//   * no software pipelining inside a warp any more (the paired kernel held two S tiles per thread): one S tile in
//     registers, S, P and O single-buffered per slot in TMEM (3 x (64 + 32 + 64) = 480 of 512 columns); latency is
//     hidden by the other two warps of the scheduler instead;
//   * the two MMA issuers walk (KV tile, slot) in a FIXED order with blocking waits.  That order keeps the three
//     slots ~600 cycles apart, which is what lets their exponential phases interleave: with one independent issuer
//     thread per slot the slots fell into phase (all three exponentiating, then all three idle) and the kernel was
//     5 % slower; polling issuers stole issue slots from the softmax warps they share schedulers with (-25 %);
//   * the ragged last KV tile costs only its real 16-key groups (exponentials and P columns of fully masked groups
//     are skipped, the PV product stops at the last real group);
//   * the normalised output leaves through a per-warp 32 x 128 B staging tile and ONE TMA store per (warp, item):
//     storing from registers (thread = row: 32 distinct lines per instruction) held a warp in the load/store unit
//     for 1100-2250 cycles per item (tools/attn_trace.py), and the (group, head, sequence) of the next item is
//     derived incrementally instead of with four integer divisions;
//   * rotary point encoding (rotaryencoderpcd.py:6-27) inside the kernel: a dedicated warp rotates head dims 0..5 of
//     every Q / K tile in shared memory between the TMA arrival and the first MMA that reads it.
//   * 1..4 keys beyond the last full KV tile (every registered sequence length) do not take a step of their own: see
//     the TAIL template parameter of the kernel.
// Measured at the bench shape (128 sequences x 8 heads, L = 1026): 366 us (400 with the two tail keys as a 17th step)
// against 485 us for the paired kernel (361 / 419 us at L = 1024), XU pipe 72 % busy (57 %).  TOKEN = 1 keeps a negative result runnable: a per-scheduler
// MUFU token (an mbarrier passed round-robin so that exactly one warp exponentiates at a time, handed on 16
// exponentials early) costs more in hand-offs than the convoys it prevents (+12 %); so did issuing the 64
// exponentials of a tile as one uninterrupted MUFU run, dropping the per-tile row maximum in favour of a
// sum-triggered re-referencing, and evaluating 1/8-1/2 of the exponentials with an FMA-pipe polynomial.
// Warp roles (512 threads): 0 = TMA producer (Q, K, V), 1 = QK^T issuer, 2 = rotary warp, 3 = PV issuer,
// 4..15 = softmax: slot = (warp - 4) / 4, TMEM lane quarter = warp % 4, thread = query row.
#include <type_traits>

#include "common.cuh"
#include "tc_sm100.cuh"

namespace pcd {

using namespace tc;

namespace a8 {

constexpr int BQ = 128, HD = 64;
constexpr int Q_TILE = BQ * HD * 2;   // 16 KB
constexpr int O_STAGE = 32 * HD * 2;  // per softmax warp: 32 rows x 128 B of normalised output for one TMA store
constexpr int BAR_BYTES = 1024;
constexpr int TMEM_COLS = 512;
constexpr int MAX_TAIL = 16;   // TAIL template values: 0 (none), 4 or 16 = tail score columns a softmax thread keeps
constexpr int MAX_NT = 3;   // (trace buffers)

// Two geometries of the same kernel:
//   WIDE = 0: three query tiles ("slots") per CTA, 64-key steps  -- 12 softmax warps, three per scheduler;
//   WIDE = 1: two slots, 128-key steps -- 8 softmax warps, two per scheduler, but each step carries twice the
//             exponentials for the same barrier round trips (S 128 + P 64 + O 64 columns per slot = all of TMEM, so
//             no tail-key form: a ragged last step is masked).
template <int WIDE>
struct Geo {
  static constexpr int NT = WIDE ? 2 : 3, BKV = WIDE ? 128 : 64;
  static constexpr int KV_TILE = BKV * HD * 2;   // 8 / 16 KB
  static constexpr int KSK = WIDE ? 3 : 4, KSV = KSK;
  static constexpr int TILE_BYTES = 2 * NT * Q_TILE + (KSK + KSV) * KV_TILE + 4 * NT * O_STAGE;  // 208 / 192 KB
  static constexpr int SMEM_BYTES = TILE_BYTES + 1024 + BAR_BYTES;
  static constexpr int THREADS = 128 + 128 * NT;
  static constexpr int SOFTMAX_REGS = WIDE ? 208 : 152;
  // TMEM columns per slot: S BKV (fp32), P BKV / 2 (bf16 pairs), O 64.  TAIL kernels (WIDE = 0) give P eight more
  // columns for the 16-key tail group: S 0..191, O 192..383, P 384..503.
  __host__ __device__ static constexpr int s_col(int s) { return s * BKV; }
  template <int TAIL> __host__ __device__ static constexpr int p_col(int s) {
    return WIDE ? 256 + s * 64 : (TAIL ? 384 + s * 40 : 192 + s * 32);
  }
  template <int TAIL> __host__ __device__ static constexpr int o_col(int s) {
    return WIDE ? 384 + s * 64 : (TAIL ? 192 + s * 64 : 288 + s * 64);
  }
  // barrier slots (8 bytes each)
  static constexpr int B_Q_FULL = 0, B_Q_EMPTY = B_Q_FULL + 2 * NT, B_Q_ROT = B_Q_EMPTY + 2 * NT,
                       B_K_FULL = B_Q_ROT + 2 * NT, B_K_EMPTY = B_K_FULL + KSK, B_K_ROT = B_K_EMPTY + KSK,
                       B_V_FULL = B_K_ROT + KSK, B_V_EMPTY = B_V_FULL + KSV, B_S_FULL = B_V_EMPTY + KSV,
                       B_S_FREE = B_S_FULL + NT, B_P_READY = B_S_FREE + NT, B_PV_DONE = B_P_READY + NT,
                       B_O_FREE = B_PV_DONE + NT, B_TOK = B_O_FREE + NT, B_T_FULL = B_TOK + 4 * NT,
                       B_T_FREE = B_T_FULL + NT, B_COUNT = B_T_FREE + NT;
  static_assert(B_COUNT * 8 + 16 <= BAR_BYTES, "barrier block too small");
  static_assert(SMEM_BYTES <= 232448, "shared memory budget");
};

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
__device__ __forceinline__ void tma4(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// exponential that stays where the source puts it (between the token acquire and release)
__device__ __forceinline__ float ex2v(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void tmem_st_32x32b_x8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x4(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void pin(uint32_t& v) { asm volatile("" : "+r"(v)); }

// work item = (group of NT (2 or 3) consecutive query tiles, head, sequence); consecutive items share (sequence, head)
struct Item {
  int g, h, b, n_act;
};
// The items of a CTA are blockIdx.x, + gridDim.x, ...: decoded once with divisions, then advanced with additions
// and conditional subtractions (four integer divisions by run-time values per item and warp were ~400 cycles of
// every item boundary).
struct ItemIter {
  int g, h, b;            // current item
  int dg, dh, db;         // gridDim.x decomposed in the (group, head, sequence) mixed radix
  int ngroups, heads, nq, nt;
  __device__ __forceinline__ ItemIter(int first, int stride, int ngroups_, int heads_, int nq_, int nt_)
      : ngroups(ngroups_), heads(heads_), nq(nq_), nt(nt_) {
    g = first % ngroups;
    const int bh = first / ngroups;
    h = bh % heads;
    b = bh / heads;
    dg = stride % ngroups;
    const int dbh = stride / ngroups;
    dh = dbh % heads;
    db = dbh / heads;
  }
  __device__ __forceinline__ Item get() const { return Item{g, h, b, min(nt, nq - g * nt)}; }
  __device__ __forceinline__ void next() {
    g += dg;
    int carry = 0;
    if (g >= ngroups) {
      g -= ngroups;
      carry = 1;
    }
    h += dh + carry;
    carry = 0;
    if (h >= heads) {
      h -= heads;
      carry = 1;
    }
    b += db + carry;
  }
};

// 3-axis rotation of head dims 0..5 of one 128-byte tile row held in 128B-swizzled shared memory
// (apply_rotary_pos_emb, rotaryencoderpcd.py:6-27: out[0:3] = even * cos - odd * sin, out[3:6] = even * sin + odd * cos)
__device__ __forceinline__ void rope_row(uint32_t tile, int r, const float* __restrict__ coord) {
  const uint32_t addr = tile + r * 128 + ((r & 7) << 4);  // 16-byte chunk 0 of row r sits at chunk position r & 7
  uint32_t w0, w1, w2, w3;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3) : "r"(addr));
  float s0, c0, s1, c1, s2, c2;
  sincosf(coord[0] * 3.14159265358979323846f, &s0, &c0);
  sincosf(coord[1] * 3.14159265358979323846f, &s1, &c1);
  sincosf(coord[2] * 3.14159265358979323846f, &s2, &c2);
  const float e0 = __uint_as_float(w0 << 16), o0 = __uint_as_float(w0 & 0xffff0000u);
  const float e1 = __uint_as_float(w1 << 16), o1 = __uint_as_float(w1 & 0xffff0000u);
  const float e2 = __uint_as_float(w2 << 16), o2 = __uint_as_float(w2 & 0xffff0000u);
  w0 = pack_bf16x2(e0 * c0 - o0 * s0, e1 * c1 - o1 * s1);
  w1 = pack_bf16x2(e2 * c2 - o2 * s2, e0 * s0 + o0 * c0);
  w2 = pack_bf16x2(e1 * s1 + o1 * c1, e2 * s2 + o2 * c2);
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(w0), "r"(w1), "r"(w2), "r"(w3) : "memory");
}

constexpr float kRescaleThreshold = 8.f;  // log2 units: P <= 2^8, safe in bf16 / fp32

// Timeline of the softmax warps of CTA 0 (tools build only: -DPCD_ATTN_TRACE, tools/attn_trace.py): clock stamps at
// the points of a step where a warp can be held up, [slot][step][point].
#ifdef PCD_ATTN_TRACE
constexpr int TRACE_STEPS = 96, TRACE_PTS = 12;
__device__ unsigned long long g_trace[MAX_NT * TRACE_STEPS * TRACE_PTS];
#define PCD_TRACE(pt)                                                                                  \
  do {                                                                                                 \
    if (blockIdx.x == 0 && quarter == 0 && lane == 0 && trace_step < TRACE_STEPS)                     \
      g_trace[(slot * TRACE_STEPS + trace_step) * TRACE_PTS + (pt)] = clock64();                       \
  } while (0)
#else
#define PCD_TRACE(pt) do { } while (0)
#endif

// TOKEN = 0 (default): warps exponentiate whenever they are ready; 1: MUFU hand-off ring (A/B of the ring itself)
// TAIL = 4 / 16: len_kv = 64 n + t with 0 <= t <= TAIL (every registered config: L = 1025 / 1026 / 1281 / 4097 / 4353 have
// t = 1 or 2; with t = 0 the tail products are all zeros and what is gained is a step loop without masking code).  The t keys do not get a KV step of their own (a step whose exponentials are already skipped still costs
// its barrier round trips: 18.5 us of 399 at L = 1026).  Instead
//   * S_tail[s] = Q[s] K_tail^T (N = 16) is issued once per item, right after the first QK^T, into the slot's O columns --
//     free until the first PV of the item -- and read (4 columns) by the softmax warps during step 0;
//   * the tail scores join the row maximum, exponentials and row sum of the LAST full step, their P goes to eight extra
//     P columns, and the PV issuer appends one K = 16 product with the V_tail tile (rows beyond len_kv are zero-filled by
//     TMA, the matching P entries are zero).
template <int TOKEN, int TAIL, int WIDE>
__global__ void __launch_bounds__(Geo<WIDE>::THREADS, 1)
attn_bf16_tc8_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO,
                     int len_q, int len_kv, float scale_log2, int nq, int ngroups, int heads, int n_items,
                     const float* __restrict__ rope) {
  using G = Geo<WIDE>;
  static_assert(!(WIDE && (TAIL || TOKEN)), "the wide geometry has neither the tail-key form nor the token ring");
  constexpr int NT = G::NT, BKV = G::BKV, KV_TILE = G::KV_TILE, KSK = G::KSK, KSV = G::KSV, TILE_BYTES = G::TILE_BYTES;
  constexpr int B_Q_FULL = G::B_Q_FULL, B_Q_EMPTY = G::B_Q_EMPTY, B_Q_ROT = G::B_Q_ROT, B_K_FULL = G::B_K_FULL,
                B_K_EMPTY = G::B_K_EMPTY, B_K_ROT = G::B_K_ROT, B_V_FULL = G::B_V_FULL, B_V_EMPTY = G::B_V_EMPTY,
                B_S_FULL = G::B_S_FULL, B_S_FREE = G::B_S_FREE, B_P_READY = G::B_P_READY, B_PV_DONE = G::B_PV_DONE,
                B_O_FREE = G::B_O_FREE, B_TOK = G::B_TOK, B_T_FULL = G::B_T_FULL, B_T_FREE = G::B_T_FREE,
                B_COUNT = G::B_COUNT;
  extern __shared__ unsigned char smem_raw[];
  uint32_t smem = (smem_u32(smem_raw) + 1023u) & ~1023u;
  pin(smem);
  const uint32_t sQ = smem;                      // [NT][2] 16 KB
  const uint32_t sK = sQ + 2 * NT * Q_TILE;      // [KSK] 8 KB
  const uint32_t sV = sK + KSK * KV_TILE;        // [KSV] 8 KB
  const uint32_t sO = sV + KSV * KV_TILE;        // [4 * NT] 4 KB output staging, one per softmax warp
  uint32_t bars = smem + TILE_BYTES;
  pin(bars);
  auto bar = [&](int slot) -> uint32_t { return bars + 8u * slot; };
  const uint32_t tmem_slot = bars + 8 * B_COUNT;

  uint32_t tid = threadIdx.x;
  pin(tid);
  const int warp = tid >> 5, lane = tid & 31;
  const int num_kv = TAIL ? len_kv / BKV : (len_kv + BKV - 1) / BKV;   // KV steps (TAIL: full tiles only)
  const int last_valid = TAIL ? BKV : len_kv - (num_kv - 1) * BKV;      // keys in the last KV tile (1..64)
  const int tail = TAIL ? len_kv - num_kv * BKV : 0;                    // keys handled outside the steps (TAIL)
  const int my_items = (n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  ItemIter items((int)blockIdx.x, (int)gridDim.x, ngroups, heads, nq, NT);  // every role walks its own copy

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmQ);
    prefetch_tensormap(&tmK);
    prefetch_tensormap(&tmV);
    prefetch_tensormap(&tmO);
    for (int i = 0; i < 2 * NT; ++i) {
      bar_init(bar(B_Q_FULL + i), 1);
      bar_init(bar(B_Q_EMPTY + i), 1);
      bar_init(bar(B_Q_ROT + i), 1);
    }
    for (int i = 0; i < KSK; ++i) {
      bar_init(bar(B_K_FULL + i), 1);
      bar_init(bar(B_K_EMPTY + i), 1);
      bar_init(bar(B_K_ROT + i), 1);
    }
    for (int i = 0; i < KSV; ++i) {
      bar_init(bar(B_V_FULL + i), 1);
      bar_init(bar(B_V_EMPTY + i), 1);
    }
    for (int s = 0; s < NT; ++s) {
      bar_init(bar(B_S_FULL + s), 1);
      bar_init(bar(B_S_FREE + s), 4);  // one arrival per softmax warp of the slot
      bar_init(bar(B_P_READY + s), 4);
      bar_init(bar(B_PV_DONE + s), 1);
      bar_init(bar(B_O_FREE + s), 4);
      bar_init(bar(B_T_FULL + s), 1);
      bar_init(bar(B_T_FREE + s), 4);
    }
    for (int i = 0; i < 4 * NT; ++i) bar_init(bar(B_TOK + i), 1);
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
    setmaxnreg_dec<56>();
    if (warp == 0) {
      // ----------------------- TMA producer: Q tiles of the item, then K_j, V_j -----------------------
      if (elect_one()) {
        int kst = 0, vst = 0;
        uint32_t kph = 1, vph = 1;  // "empty" barriers: the first pass over a ring does not block
        uint32_t qcnt[NT] = {};
        for (int k = 0; k < my_items; ++k, items.next()) {
          const Item it = items.get();
#pragma unroll
          for (int s = 0; s < NT; ++s) {
            if (s < it.n_act) {
              const int qi = s * 2 + (qcnt[s] & 1);
              bar_wait(bar(B_Q_EMPTY + qi), ((qcnt[s] >> 1) & 1) ^ 1);
              bar_expect_tx(bar(B_Q_FULL + qi), Q_TILE);
              tma4(sQ + qi * Q_TILE, &tmQ, bar(B_Q_FULL + qi), 0, it.h, (it.g * NT + s) * BQ, it.b);
              ++qcnt[s];
            }
          }
          // ring order (the consumers walk it the same way): K_0, [K_tail], V_0, K_1, V_1, ..., [V_tail]
          for (int j = 0; j < num_kv + (TAIL ? 1 : 0); ++j) {
            if (j < num_kv) {
              bar_wait(bar(B_K_EMPTY + kst), kph);
              bar_expect_tx(bar(B_K_FULL + kst), KV_TILE);
              tma4(sK + kst * KV_TILE, &tmK, bar(B_K_FULL + kst), 0, it.h, j * BKV, it.b);
              if (++kst == KSK) {
                kst = 0;
                kph ^= 1;
              }
            }
            if (TAIL && j == 0) {
              bar_wait(bar(B_K_EMPTY + kst), kph);
              bar_expect_tx(bar(B_K_FULL + kst), KV_TILE);
              tma4(sK + kst * KV_TILE, &tmK, bar(B_K_FULL + kst), 0, it.h, num_kv * BKV, it.b);
              if (++kst == KSK) {
                kst = 0;
                kph ^= 1;
              }
            }
            bar_wait(bar(B_V_EMPTY + vst), vph);
            bar_expect_tx(bar(B_V_FULL + vst), KV_TILE);
            tma4(sV + vst * KV_TILE, &tmV, bar(B_V_FULL + vst), 0, it.h, j * BKV, it.b);   // j == num_kv: V_tail
            if (++vst == KSV) {
              vst = 0;
              vph ^= 1;
            }
          }
        }
      }
    } else if (warp == 2) {
      // ------------- rotary warp: rotate head dims 0..5 of the landed Q / K tiles in place -------------
      if (rope != nullptr) {
        int kst = 0;
        uint32_t kph = 0;
        uint32_t qcnt[NT] = {};
        for (int k = 0; k < my_items; ++k, items.next()) {
          const Item it = items.get();
#pragma unroll
          for (int s = 0; s < NT; ++s) {
            if (s < it.n_act) {
              const int qi = s * 2 + (qcnt[s] & 1);
              bar_wait(bar(B_Q_FULL + qi), (qcnt[s] >> 1) & 1);
              const int q0 = (it.g * NT + s) * BQ;
              for (int r = lane; r < BQ; r += 32)
                if (q0 + r < len_q) rope_row(sQ + qi * Q_TILE, r, rope + ((int64_t)it.b * len_q + q0 + r) * 3);
              fence_proxy_async_smem();
              __syncwarp();
              if (lane == 0) bar_arrive(bar(B_Q_ROT + qi));
              ++qcnt[s];
            }
          }
          for (int j = 0; j < num_kv + (TAIL ? 1 : 0); ++j) {
            // ring order K_0, [K_tail], K_1, ...: tile index of the j-th arrival
            const int tile = !TAIL ? j : (j == 0 ? 0 : (j == 1 ? num_kv : j - 1));
            bar_wait(bar(B_K_FULL + kst), kph);
            for (int r = lane; r < BKV; r += 32)
              if (tile * BKV + r < len_kv) rope_row(sK + kst * KV_TILE, r, rope + ((int64_t)it.b * len_kv + tile * BKV + r) * 3);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) bar_arrive(bar(B_K_ROT + kst));
            if (++kst == KSK) {
              kst = 0;
              kph ^= 1;
            }
          }
        }
      }
    } else if (warp == 1) {
      // ------------------------ QK^T issuer: S[s] = Q[s] K_j^T for every active slot ------------------------
      if (elect_one()) {
        constexpr uint32_t idesc_full = idesc_bf16_f32(BQ, BKV, 0);
        const uint32_t idesc_last = idesc_bf16_f32(BQ, max(16, (last_valid + 15) & ~15), 0);
        const int qbar = rope != nullptr ? B_Q_ROT : B_Q_FULL;
        const int kbar = rope != nullptr ? B_K_ROT : B_K_FULL;
        int kst = 0;
        uint32_t kph = 0;
        uint32_t qcnt[NT] = {};
        uint32_t su[NT] = {};  // S products issued per slot
        uint32_t nq_items[NT] = {};  // items the slot took part in (TAIL: parity of O_FREE)
        constexpr uint32_t idesc_tail = idesc_bf16_f32(BQ, 16, 0);
        for (int k = 0; k < my_items; ++k, items.next()) {
          const Item it = items.get();
          uint64_t adesc[NT];
          int qi[NT];
#pragma unroll
          for (int s = 0; s < NT; ++s) {
            qi[s] = s * 2 + (qcnt[s] & 1);
            adesc[s] = smem_desc_sw128(sQ + qi[s] * Q_TILE);
            if (s < it.n_act) {
              bar_wait(bar(qbar + qi[s]), (qcnt[s] >> 1) & 1);
              ++qcnt[s];
            }
          }
          for (int j = 0; j < num_kv; ++j) {
            bar_wait(bar(kbar + kst), kph);
            const bool last = (j == num_kv - 1);
            const uint32_t idesc = last ? idesc_last : idesc_full;
            const uint64_t bdesc = smem_desc_sw128(sK + kst * KV_TILE);
#pragma unroll
            for (int s = 0; s < NT; ++s) {
              if (s < it.n_act) {
                bar_wait(bar(B_S_FREE + s), (su[s] & 1) ^ 1);  // previous S of the slot sits in registers
                ++su[s];
                tcgen05_fence_after();
                const uint32_t s_tmem = tmem_base + G::s_col(s);
#pragma unroll
                for (int kk = 0; kk < HD / 16; ++kk) umma_bf16_ss(s_tmem, adesc[s] + 2 * kk, bdesc + 2 * kk, idesc, kk != 0);
                commit(bar(B_S_FULL + s));
              }
            }
            commit(bar(B_K_EMPTY + kst));
            if (++kst == KSK) {
              kst = 0;
              kph ^= 1;
            }
            if (TAIL && j == 0) {
              // S_tail[s] -> the slot's O columns (free once the previous item's epilogue has read O)
              bar_wait(bar(kbar + kst), kph);
              const uint64_t tdesc = smem_desc_sw128(sK + kst * KV_TILE);
#pragma unroll
              for (int s = 0; s < NT; ++s) {
                if (s < it.n_act) {
                  if (nq_items[s] > 0) bar_wait(bar(B_O_FREE + s), (nq_items[s] - 1) & 1);
                  ++nq_items[s];
                  tcgen05_fence_after();
                  const uint32_t t_tmem = tmem_base + G::template o_col<TAIL>(s);
#pragma unroll
                  for (int kk = 0; kk < HD / 16; ++kk) umma_bf16_ss(t_tmem, adesc[s] + 2 * kk, tdesc + 2 * kk, idesc_tail, kk != 0);
                  commit(bar(B_T_FULL + s));
                }
              }
              commit(bar(B_K_EMPTY + kst));
              if (++kst == KSK) {
                kst = 0;
                kph ^= 1;
              }
            }
            if (last) {
#pragma unroll
              for (int s = 0; s < NT; ++s)
                if (s < it.n_act) commit(bar(B_Q_EMPTY + qi[s]));
            }
          }
        }
      }
    } else {
      // -------------------------- PV issuer: O[s] += P[s] V_j --------------------------
      if (elect_one()) {
        constexpr uint32_t idesc_pv = idesc_bf16_f32(BQ, HD, /*B MN-major*/ 1);
        const int nks_last = (last_valid + 15) >> 4;
        int vst = 0;
        uint32_t vph = 0;
        uint32_t pu[NT] = {};   // PV products issued per slot
        uint32_t ni[NT] = {};   // items the slot took part in
        for (int k = 0; k < my_items; ++k, items.next()) {
          const Item it = items.get();
          for (int j = 0; j < num_kv; ++j) {
            bar_wait(bar(B_V_FULL + vst), vph);
            const bool lastj = (j == num_kv - 1);
            const int nks = lastj ? nks_last : BKV / 16;
            const uint32_t v_addr = sV + vst * KV_TILE;
            // TAIL: the last step also consumes the V_tail tile (next ring stage)
            int tst = vst + 1;
            uint32_t tph = vph;
            if (tst == KSV) {
              tst = 0;
              tph ^= 1;
            }
            if (TAIL && lastj) bar_wait(bar(B_V_FULL + tst), tph);
            const uint64_t tdesc = smem_desc_sw128(sV + tst * KV_TILE);
#pragma unroll
            for (int s = 0; s < NT; ++s) {
              if (s < it.n_act) {
                if (j == 0 && ni[s] > 0) bar_wait(bar(B_O_FREE + s), (ni[s] - 1) & 1);  // previous O read out
                bar_wait(bar(B_P_READY + s), pu[s] & 1);
                ++pu[s];
                if (TAIL && j == 0) bar_wait(bar(B_T_FREE + s), ni[s] & 1);  // S_tail has left the O columns
                tcgen05_fence_after();
                const uint32_t o_tmem = tmem_base + G::template o_col<TAIL>(s), p_tmem = tmem_base + G::template p_col<TAIL>(s);
                for (int kk = 0; kk < nks; ++kk)
                  umma_bf16_ts(o_tmem, p_tmem + kk * 8, smem_desc_sw128(v_addr + kk * 2048), idesc_pv, (j | kk) != 0);
                if (TAIL && lastj) umma_bf16_ts(o_tmem, p_tmem + 32, tdesc, idesc_pv, 1);
                commit(bar(B_PV_DONE + s));
              }
            }
            commit(bar(B_V_EMPTY + vst));
            if (++vst == KSV) {
              vst = 0;
              vph ^= 1;
            }
            if (TAIL && lastj) {
              commit(bar(B_V_EMPTY + vst));
              if (++vst == KSV) {
                vst = 0;
                vph ^= 1;
              }
            }
          }
#pragma unroll
          for (int s = 0; s < NT; ++s)
            if (s < it.n_act) ++ni[s];
        }
      }
    }
  } else {
    setmaxnreg_inc<G::SOFTMAX_REGS>();
    // --------------------------- softmax: thread = query row ----------------------------
    const int slot = (warp - 4) >> 2, quarter = warp & 3;
    uint32_t tm = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    pin(tm);
    const uint32_t b_sfull = bars + 8 * (B_S_FULL + slot), b_sfree = bars + 8 * (B_S_FREE + slot),
                   b_pready = bars + 8 * (B_P_READY + slot), b_pvdone = bars + 8 * (B_PV_DONE + slot),
                   b_ofree = bars + 8 * (B_O_FREE + slot), b_tok = bars + 8 * (B_TOK + quarter * NT),
                   b_tfull = bars + 8 * (B_T_FULL + slot), b_tfree = bars + 8 * (B_T_FREE + slot);
    uint32_t nis = 0;  // items this slot has taken part in (TAIL: parity of T_FULL)
    uint32_t u = 0;    // tiles this slot has processed (parity of S_FULL / P_READY / PV_DONE uses)
    uint32_t tk = 0;   // token acquisitions of this warp
    bool stores_pending = false;
    const uint64_t sc2 = pack2(scale_log2, scale_log2);

    for (int k = 0; k < my_items; ++k, items.next()) {
      const Item it = items.get();
      if (slot >= it.n_act) continue;
      const int q0 = (it.g * NT + slot) * BQ;
      if (q0 + quarter * 32 >= len_q) {
        // no real query row in this warp (ragged last query tile): keep the slot's barrier counts in step
        if (TAIL && lane == 0) bar_arrive(b_tfree);
        ++nis;
        for (int j = 0; j < num_kv; ++j) {
          if (lane == 0) {
            bar_arrive(b_sfree);
            bar_arrive(b_pready);
          }
          bar_wait(b_pready, u & 1);  // all four warps of the slot are past this tile
          ++u;
        }
        if (lane == 0) bar_arrive(b_ofree);
        continue;
      }
      // slots of this item whose quarter holds real rows (a prefix): the members of this scheduler's token ring
      int ring = 0;
#pragma unroll
      for (int s = 0; s < NT; ++s) ring += (s < it.n_act && (it.g * NT + s) * BQ + quarter * 32 < len_q) ? 1 : 0;
      const uint32_t tok_next = b_tok + 8 * ((slot + 1 < ring) ? slot + 1 : 0);

      float m_used = 0.f, l_run = 0.f;
      float ts[TAIL ? TAIL : 1];   // TAIL: raw scores of the tail keys (-inf beyond the real ones)
      // one KV step; FIRST / LAST are compile-time so that the steady-state instance carries neither the item set-up
      // nor the masking / tail code of the last step (a step loop without per-chunk predicates measured 4.5 % faster)
      auto kv_step = [&](const int j, auto first_c, auto last_c) {
        constexpr bool FIRST = decltype(first_c)::value, LAST = decltype(last_c)::value;
        uint32_t r[BKV];
#ifdef PCD_ATTN_TRACE
        const int trace_step = k * num_kv + j;
#endif
        PCD_TRACE(0);  // step begins
        bar_wait(b_sfull, u & 1);
        PCD_TRACE(1);  // S_j has arrived
        tcgen05_fence_after();
#pragma unroll
        for (int c = 0; c < BKV / 32; ++c) tmem_ld_32x32b_x32(tm + G::s_col(slot) + c * 32, r + c * 32);
        tmem_ld_wait();
        PCD_TRACE(2);  // S_j in registers
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) bar_arrive(b_sfree);  // S sits in registers: the next QK^T of the slot may overwrite it
        if (TAIL && FIRST) {
          // the tail scores wait in the O columns: fetch them before the first PV of the item overwrites O
          bar_wait(b_tfull, nis & 1);
          ++nis;
          tcgen05_fence_after();
          uint32_t tr[TAIL ? TAIL : 1];
          if (TAIL == 4) tmem_ld_32x32b_x4(tm + G::template o_col<TAIL>(slot), tr);
          if (TAIL == 16) tmem_ld_32x32b_x16(tm + G::template o_col<TAIL>(slot), tr);
          tmem_ld_wait();
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) bar_arrive(b_tfree);
#pragma unroll
          for (int i = 0; i < TAIL; ++i) ts[i] = i < tail ? __uint_as_float(tr[i]) : -INFINITY;
        }
        constexpr bool with_tail = TAIL && LAST;
        const int nvalid = LAST ? last_valid : BKV;   // (compile-time BKV in every step but the last)
        if (nvalid < BKV) {
#pragma unroll
          for (int i = 0; i < BKV; ++i)
            if (i >= nvalid) r[i] = 0xff800000u;  // -inf
        }
        float mx;
        {
          float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
          for (int i = 0; i < BKV; i += 4) {
            mx0 = fmaxf(mx0, __uint_as_float(r[i]));
            mx1 = fmaxf(mx1, __uint_as_float(r[i + 1]));
            mx2 = fmaxf(mx2, __uint_as_float(r[i + 2]));
            mx3 = fmaxf(mx3, __uint_as_float(r[i + 3]));
          }
          if (with_tail) {
#pragma unroll
            for (int i = 0; i < TAIL; i += 4) {
              mx0 = fmaxf(mx0, ts[i]);
              mx1 = fmaxf(mx1, ts[i + 1]);
              mx2 = fmaxf(mx2, ts[i + 2]);
              mx3 = fmaxf(mx3, ts[i + 3]);
            }
          }
          mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * scale_log2;
        }
        // P buffer / O of the slot are free once the previous PV product has completed
        PCD_TRACE(3);  // row maximum done
        bar_wait(b_pvdone, (u & 1) ^ 1);
        PCD_TRACE(4);  // PV_{j-1} complete
        if (FIRST) {
          m_used = mx;
        } else {
          // lazy rescale of O and l when the maximum grows by more than 2^8 (rare)
          const bool grow = mx > m_used + kRescaleThreshold;
          if (__any_sync(0xffffffffu, grow)) {
            tcgen05_fence_after();
            const float alpha = grow ? ex2_approx(m_used - mx) : 1.f;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              uint32_t o[32];
              tmem_ld_32x32b_x32(tm + G::template o_col<TAIL>(slot) + h * 32, o);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
              tmem_st_32x32b_x32(tm + G::template o_col<TAIL>(slot) + h * 32, o);
            }
            tmem_st_wait();
            l_run *= alpha;
            if (grow) m_used = mx;
          }
        }
        const uint64_t nm2 = pack2(-m_used, -m_used);
        // ---- exponentials, under this scheduler's MUFU token ----
        if (TOKEN) bar_wait(b_tok + 8 * slot, slot == 0 ? ((tk & 1) ^ 1) : (tk & 1));
        ++tk;
        PCD_TRACE(5);  // exponentials begin
        tcgen05_fence_after();
        uint64_t sum_a = 0ull, sum_b = 0ull;
#pragma unroll
        for (int c = 0; c < BKV / 16; ++c) {
          if (c * 16 < nvalid) {
            uint32_t pk[8];
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
              const int e = c * 16 + i;
              const uint64_t xa = fma2(pack2(__uint_as_float(r[e]), __uint_as_float(r[e + 1])), sc2, nm2);
              const uint64_t xb = fma2(pack2(__uint_as_float(r[e + 2]), __uint_as_float(r[e + 3])), sc2, nm2);
              float a0, a1, b0, b1;
              unpack2(xa, a0, a1);
              unpack2(xb, b0, b1);
              const float p0 = ex2v(a0), p1 = ex2v(a1), p2 = ex2v(b0), p3 = ex2v(b1);
              sum_a = add2(sum_a, pack2(p0, p1));
              sum_b = add2(sum_b, pack2(p2, p3));
              pk[i >> 1] = pack_bf16x2(p0, p1);
              pk[(i >> 1) + 1] = pack_bf16x2(p2, p3);
            }
            tmem_st_32x32b_x8(tm + G::template p_col<TAIL>(slot) + c * 8, pk);
          }
          if (TOKEN && c == 2) {
            // pass the token on one group early: the next warp's wake-up overlaps the last 16 exponentials
            __syncwarp();
            if (lane == 0) bar_arrive(tok_next);
          }
        }
        {
          float s0, s1, s2, s3;
          unpack2(sum_a, s0, s1);
          unpack2(sum_b, s2, s3);
          l_run += (s0 + s1) + (s2 + s3);
        }
        if (with_tail) {
          // the tail keys: same reference maximum, P into the eight extra columns (keys beyond the tail: exp2(-inf) = 0)
          uint32_t pk[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
#pragma unroll
          for (int i = 0; i < TAIL; i += 4) {
            const float p0 = ex2v(fmaf(ts[i], scale_log2, -m_used)), p1 = ex2v(fmaf(ts[i + 1], scale_log2, -m_used));
            const float p2 = ex2v(fmaf(ts[i + 2], scale_log2, -m_used)), p3 = ex2v(fmaf(ts[i + 3], scale_log2, -m_used));
            l_run += (p0 + p1) + (p2 + p3);
            pk[i >> 1] = pack_bf16x2(p0, p1);
            pk[(i >> 1) + 1] = pack_bf16x2(p2, p3);
          }
          tmem_st_32x32b_x8(tm + G::template p_col<TAIL>(slot) + 32, pk);
        }
        PCD_TRACE(6);  // exponentials issued
        tmem_st_wait();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) bar_arrive(b_pready);  // P in TMEM -> PV may be issued
        PCD_TRACE(7);  // P handed over
        ++u;
      };
      using T_ = std::integral_constant<bool, true>;
      using F_ = std::integral_constant<bool, false>;
      if (num_kv == 1) {
        kv_step(0, T_{}, T_{});
      } else {
        kv_step(0, T_{}, F_{});
        for (int j = 1; j < num_kv - 1; ++j) kv_step(j, F_{}, F_{});
        kv_step(num_kv - 1, F_{}, T_{});
      }
      // ---- epilogue: O / l -> global ----
#ifdef PCD_ATTN_TRACE
      const int trace_step = k * num_kv + num_kv - 1;
#endif
      PCD_TRACE(8);   // epilogue begins
      bar_wait(b_pvdone, (u & 1) ^ 1);  // every PV product of the item has landed in O
      PCD_TRACE(9);   // last PV complete
      tcgen05_fence_after();
      uint32_t o[HD];
      tmem_ld_32x32b_x32(tm + G::template o_col<TAIL>(slot), o);
      tmem_ld_32x32b_x32(tm + G::template o_col<TAIL>(slot) + 32, o + 32);
      tmem_ld_wait();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) bar_arrive(b_ofree);  // O is in registers: the slot's next item may overwrite it
      PCD_TRACE(10);  // O in registers
      // O / l -> bf16 -> this warp's 32 x 128 B staging tile (128B-swizzled like the tensor map) -> ONE TMA store.
      // Storing from registers (each thread its own 128-byte row: 32 distinct lines per instruction) kept a warp
      // in the load/store unit for 1100-2250 cycles per item (tools/attn_trace.py); rows >= len_q are clipped by
      // the tensor map.
      const uint32_t stage = sO + (warp - 4) * O_STAGE;
      if (stores_pending) {
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // previous item's store has read the tile
        __syncwarp();
      }
      {
        const float inv = 1.f / l_run;
        const uint32_t row_addr = stage + lane * 128;
#pragma unroll
        for (int i = 0; i < HD; i += 8) {
          float v[8];
#pragma unroll
          for (int t = 0; t < 8; ++t) v[t] = __uint_as_float(o[i + t]) * inv;
          const uint32_t a = row_addr + ((((uint32_t)i >> 3) ^ ((uint32_t)lane & 7u)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(pack_bf16x2(v[0], v[1])),
                       "r"(pack_bf16x2(v[2], v[3])), "r"(pack_bf16x2(v[4], v[5])), "r"(pack_bf16x2(v[6], v[7]))
                       : "memory");
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                         reinterpret_cast<uint64_t>(&tmO)),
                     "r"(stage), "r"(0), "r"(it.h), "r"(q0 + quarter * 32), "r"(it.b)
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
      stores_pending = true;
      PCD_TRACE(11);  // output handed to the TMA engine
    }
    if (stores_pending && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // outputs globally visible
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace a8

#ifdef PCD_ATTN_TRACE
extern "C" __attribute__((visibility("default"))) int pcd_attn_trace_read(unsigned long long* dst, int n) {
  const int total = a8::MAX_NT * a8::TRACE_STEPS * a8::TRACE_PTS;
  return cudaMemcpyFromSymbol(dst, a8::g_trace, sizeof(unsigned long long) * (n < total ? n : total)) == cudaSuccess ? total : -1;
}
#endif

template <int TOKEN, int TAIL, int WIDE>
static int launch_tc8(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& to, int batch, int heads, int len_q, int len_kv, float scale_log2, const float* rope,
                      cudaStream_t st) {
  using G = a8::Geo<WIDE>;
  auto kern = a8::attn_bf16_tc8_kernel<TOKEN, TAIL, WIDE>;
  // the attribute is per device: set it on every launch (cheap) rather than caching a per-process flag
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM_BYTES);
  if (e != cudaSuccess) {
    set_error("attention_bf16: kernel attribute setup (%d B smem): %s", G::SMEM_BYTES, cudaGetErrorString(e));
    return PCD_ERR_CUDA;
  }
  const int nq = ceil_div(len_q, a8::BQ);
  const int ngroups = ceil_div(nq, G::NT);
  const int64_t n_items64 = (int64_t)ngroups * heads * batch;
  if (n_items64 > 0x7fffffff) {
    set_error("attention_bf16: too many (query group, head, sequence) items");
    return PCD_ERR_INVALID;
  }
  const int n_items = (int)n_items64;
  const int grid = min(n_items, num_sms());  // one CTA per SM: ~200 KB shared memory, all 512 TMEM columns
  kern<<<grid, G::THREADS, G::SMEM_BYTES, st>>>(tq, tk, tv, to, len_q, len_kv, scale_log2, nq, ngroups, heads,
                                                n_items, rope);
  PCD_CHECK_LAUNCH("attention_bf16");
  return PCD_OK;
}

// rows of a K / V tensor-map box (= keys per step) of a mode
int attn_tc8_kv_rows(int mode) { return mode == 3 ? a8::Geo<1>::BKV : a8::Geo<0>::BKV; }

// mode: 0 = free-running softmax warps, short tails ride on the last step (default); 1 = MUFU hand-off ring;
// 2 = like 0 but a short tail stays an ordinary masked KV step (A/B measurements); 3 = wide geometry (two slots,
// 128-key steps; tk / tv must be built with 128-row boxes, attn_tc8_kv_rows)
int launch_attn_tc8(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& to, int batch, int heads, int len_q, int len_kv, float scale_log2, const float* rope,
                    int mode, cudaStream_t st) {
  constexpr int BKV = a8::Geo<0>::BKV;
  const bool no_tail = mode == 2;
  if (mode == 3) return launch_tc8<0, 0, 1>(tq, tk, tv, to, batch, heads, len_q, len_kv, scale_log2, rope, st);
  if (mode == 1) return launch_tc8<1, 0, 0>(tq, tk, tv, to, batch, heads, len_q, len_kv, scale_log2, rope, st);
  // a short tail (0..16 keys beyond the last full KV tile: every registered sequence length has 1 or 2, the 77 text
  // tokens of the perceiver 13) rides on the last step; tail == 0 takes the same kernel because its step loop has no
  // masking code at all (344 vs 360 us at L = 1024)
  const int tail = len_kv % BKV;
  if (len_kv >= BKV && tail <= 4 && !no_tail)
    return launch_tc8<0, 4, 0>(tq, tk, tv, to, batch, heads, len_q, len_kv, scale_log2, rope, st);
  if (len_kv >= BKV && tail <= a8::MAX_TAIL && !no_tail)
    return launch_tc8<0, 16, 0>(tq, tk, tv, to, batch, heads, len_q, len_kv, scale_log2, rope, st);
  return launch_tc8<0, 0, 0>(tq, tk, tv, to, batch, heads, len_q, len_kv, scale_log2, rope, st);
}

}  // namespace pcd
