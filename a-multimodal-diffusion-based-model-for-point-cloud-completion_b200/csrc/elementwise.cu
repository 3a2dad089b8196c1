// LayerNorm, timestep embedding, token assembly and output projection kernels.
// All are HBM-bound row kernels: one warp per token row, 128-bit accesses.
#include "common.cuh"

namespace pcd {

// ---------------------------------------------------------------------------
// timestep embedding (reference models/util.py:72-89)
// ---------------------------------------------------------------------------
__global__ void timestep_embed_kernel(const float* __restrict__ t, const float* __restrict__ freqs,
                                      int batch, int dim, float* __restrict__ out, int ld) {
  int half = dim / 2;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= batch * half) return;
  int b = i / half, k = i % half;
  float a = __fmul_rn(t[b], freqs[k]);
  // accurate cosf/sinf (arguments reach ~1e3 rad; no fast-math here)
  out[(size_t)b * ld + k] = cosf(a);
  out[(size_t)b * ld + half + k] = sinf(a);
  if ((dim & 1) && k == 0) out[(size_t)b * ld + dim - 1] = 0.f;
}

// ---------------------------------------------------------------------------
// Row LayerNorm in registers.  MAXV = ceil(dim/128) float4 per lane.
// mean = sum/d; var = sum((x-mean)^2)/d (two-pass, like ATen); rstd = rsqrt(var+eps).
// ---------------------------------------------------------------------------
template <int MAXV>
struct RowRegs {
  float4 v[MAXV];
};

template <int MAXV>
__device__ __forceinline__ void row_stats(const RowRegs<MAXV>& r, int dim, int lane, float& mean,
                                          float& rstd, float eps) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    int c = (i * 32 + lane) * 4;
    if (c < dim) s += (r.v[i].x + r.v[i].y) + (r.v[i].z + r.v[i].w);
  }
  mean = warp_sum(s) / (float)dim;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    int c = (i * 32 + lane) * 4;
    if (c < dim) {
      float a = r.v[i].x - mean, b = r.v[i].y - mean, cc = r.v[i].z - mean, d = r.v[i].w - mean;
      q += (a * a + b * b) + (cc * cc + d * d);
    }
  }
  rstd = rsqrtf(warp_sum(q) / (float)dim + eps);
}

// Adds y (bf16 or fp32, the previous GEMM's output) to the row held in registers.
template <int MAXV>
__device__ __forceinline__ void add_rows(RowRegs<MAXV>& r, const void* y, int y_bf16, size_t row_off,
                                         int dim, int lane) {
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    int c = (i * 32 + lane) * 4;
    if (c < dim) {
      if (y_bf16) {
        uint2 p = *reinterpret_cast<const uint2*>(reinterpret_cast<const uint16_t*>(y) + row_off + c);
        __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&p.x);
        __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&p.y);
        float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
        r.v[i].x += fa.x; r.v[i].y += fa.y; r.v[i].z += fb.x; r.v[i].w += fb.y;
      } else {
        float4 t = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(y) + row_off + c);
        r.v[i].x += t.x; r.v[i].y += t.y; r.v[i].z += t.z; r.v[i].w += t.w;
      }
    }
  }
}

// out = LayerNorm(x) -- or, with y != nullptr, the fused residual update of the block
//   x <- x + y ; out = LayerNorm(x)                     (reference transformer.py:113-114)
// where y is the output of the preceding c_proj GEMM.  One pass: 4+2 B read, 4+2 B written
// per element in bf16 mode.
template <int MAXV, bool OUT_BF16>
__global__ void __launch_bounds__(256) layernorm_kernel(float* __restrict__ x, int ldx,
                                                        const void* __restrict__ y, int ldy, int y_bf16,
                                                        const float* __restrict__ gamma,
                                                        const float* __restrict__ beta,
                                                        void* __restrict__ out, int ldo, int rows,
                                                        int dim, float eps) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  float* xr = x + (size_t)warp * ldx;
  RowRegs<MAXV> r;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    int c = (i * 32 + lane) * 4;
    if (c < dim) r.v[i] = *reinterpret_cast<const float4*>(xr + c);
  }
  if (y != nullptr) {
    add_rows<MAXV>(r, y, y_bf16, (size_t)warp * ldy, dim, lane);
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      int c = (i * 32 + lane) * 4;
      if (c < dim) *reinterpret_cast<float4*>(xr + c) = r.v[i];
    }
  }
  float mean, rstd;
  row_stats<MAXV>(r, dim, lane, mean, rstd, eps);
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    int c = (i * 32 + lane) * 4;
    if (c < dim) {
      float4 g = *reinterpret_cast<const float4*>(gamma + c);
      float4 b = *reinterpret_cast<const float4*>(beta + c);
      float4 o;
      o.x = (r.v[i].x - mean) * rstd * g.x + b.x;
      o.y = (r.v[i].y - mean) * rstd * g.y + b.y;
      o.z = (r.v[i].z - mean) * rstd * g.z + b.z;
      o.w = (r.v[i].w - mean) * rstd * g.w + b.w;
      if (OUT_BF16) {
        uint2 p = make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
        *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(out) + (size_t)warp * ldo + c) = p;
      } else {
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + (size_t)warp * ldo + c) = o;
      }
    }
  }
}

// ---------------------------------------------------------------------------
// bf16 copy of the residual stream + per-128-column (mean, M2) of the rounded values: the inputs
// of the LayerNorm-folded projections (gemm_tc.cu, PCD_EPI_LN_*).  Lane l holds columns
// i*128 + 4l .. +3 of slot i, so one warp reduction per slot.
// ---------------------------------------------------------------------------
template <int MAXV>
__global__ void __launch_bounds__(256) cast_rowstats_kernel(const float* __restrict__ x, int ldx,
                                                            uint16_t* __restrict__ out, int ldo,
                                                            float2* __restrict__ stats, int rows, int dim) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const float* xr = x + (size_t)warp * ldx;
  const int slots = dim / 128;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    if (i < slots) {
      const int c = (i * 32 + lane) * 4;
      const float4 v = *reinterpret_cast<const float4*>(xr + c);
      const uint32_t p0 = pack_bf16x2(v.x, v.y), p1 = pack_bf16x2(v.z, v.w);
      *reinterpret_cast<uint2*>(out + (size_t)warp * ldo + c) = make_uint2(p0, p1);
      const float q0 = __uint_as_float(p0 << 16), q1 = __uint_as_float(p0 & 0xffff0000u);
      const float q2 = __uint_as_float(p1 << 16), q3 = __uint_as_float(p1 & 0xffff0000u);
      const float mean = warp_sum((q0 + q1) + (q2 + q3)) * (1.f / 128.f);
      const float a = q0 - mean, b = q1 - mean, cc = q2 - mean, d = q3 - mean;
      const float m2 = warp_sum((a * a + b * b) + (cc * cc + d * d));
      if (lane == 0) stats[(size_t)warp * slots + i] = make_float2(mean, m2);
    }
  }
}

// The bf16 copy of one stream row held in registers + per-128-column (mean, M2) of the ROUNDED values, exactly as
// cast_rowstats_kernel computes them (dim % 128 == 0): the first LayerNorm-folded projection consumes them, so the
// forward saves one pass over the stream (read 4 + write 2 bytes per element).  The MAXV slices reduce side by side
// (independent shuffle chains).
template <int MAXV>
__device__ __forceinline__ void emit_bf16_stats(const RowRegs<MAXV>& r, int64_t row, int dim, int lane,
                                                uint16_t* __restrict__ hb, float2* __restrict__ stats) {
  float q[MAXV][4], sm[MAXV];
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = (i * 32 + lane) * 4;
    sm[i] = 0.f;
    if (c < dim) {
      const uint32_t p0 = pack_bf16x2(r.v[i].x, r.v[i].y), p1 = pack_bf16x2(r.v[i].z, r.v[i].w);
      *reinterpret_cast<uint2*>(hb + (size_t)row * dim + c) = make_uint2(p0, p1);
      q[i][0] = __uint_as_float(p0 << 16); q[i][1] = __uint_as_float(p0 & 0xffff0000u);
      q[i][2] = __uint_as_float(p1 << 16); q[i][3] = __uint_as_float(p1 & 0xffff0000u);
      sm[i] = (q[i][0] + q[i][1]) + (q[i][2] + q[i][3]);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int i = 0; i < MAXV; ++i) sm[i] += __shfl_xor_sync(0xffffffffu, sm[i], o);
  }
  float m2[MAXV];
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = (i * 32 + lane) * 4;
    sm[i] *= (1.f / 128.f);
    m2[i] = 0.f;
    if (c < dim) {
      const float a = q[i][0] - sm[i], b = q[i][1] - sm[i], cc = q[i][2] - sm[i], d = q[i][3] - sm[i];
      m2[i] = (a * a + b * b) + (cc * cc + d * d);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int i = 0; i < MAXV; ++i) m2[i] += __shfl_xor_sync(0xffffffffu, m2[i], o);
  }
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < MAXV; ++i)
      if (i * 128 < dim) stats[(size_t)row * (dim / 128) + i] = make_float2(sm[i], m2[i]);
  }
}

// out = a + b (fp32, out may alias a): the stream additions of the TwoStream denoiser that are not the
// tail of a projection (z + ln_latent(..), token-type embeddings; models/modules.py:228-229, model.py:536)
__global__ void add_f32_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out,
                               int64_t n4, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n4) {
    const float4 x = reinterpret_cast<const float4*>(a)[i], y = reinterpret_cast<const float4*>(b)[i];
    reinterpret_cast<float4*>(out)[i] = make_float4(x.x + y.x, x.y + y.y, x.z + y.z, x.w + y.w);
  } else if (i == n4) {
    for (int64_t k = n4 * 4; k < n; ++k) out[k] = a[k] + b[k];
  }
}

// ---------------------------------------------------------------------------
// token assembly + ln_pre (reference models/transformer.py:205-220)
// ---------------------------------------------------------------------------
constexpr int EMB_ROWS_PER_WARP = 16;

template <int MAXV>
__global__ void __launch_bounds__(256) embed_tokens_kernel(
    const float* __restrict__ x, int x_seqs, int c_in, int n_points,
    const float* __restrict__ w_in, const float* __restrict__ b_in,
    const float* __restrict__ prefix, int n_prefix, const float* __restrict__ add_cond,
    const float* __restrict__ ln_g, const float* __restrict__ ln_b, float eps,
    float* __restrict__ h, int seqs, int dim, uint16_t* __restrict__ hb, float2* __restrict__ stats) {
  // Each warp owns EMB_ROWS_PER_WARP consecutive token rows; the lane's slice of the input
  // projection (4*MAXV output features x c_in) is loaded once into shared memory per block and
  // re-used for every row, so the loop is bound by the 4*dim bytes written per token.
  extern __shared__ __align__(16) float sw[];  // [c_in + 1][dim]: transposed w_in, then the bias
  const int L = n_prefix + n_points;
  for (int i = threadIdx.x; i < dim * (c_in + 1); i += blockDim.x) {
    int k = i / dim, f = i % dim;
    sw[i] = (k < c_in) ? w_in[(size_t)f * c_in + k] : b_in[f];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t total = (int64_t)seqs * L;
  for (int rr = 0; rr < EMB_ROWS_PER_WARP; ++rr) {
    const int64_t row = warp_global * EMB_ROWS_PER_WARP + rr;
    if (row >= total) break;
    const int s = (int)(row / L), l = (int)(row % L);
    RowRegs<MAXV> r;
    if (l < n_prefix) {
      const float* p = prefix + ((size_t)s * n_prefix + l) * dim;
#pragma unroll
      for (int i = 0; i < MAXV; ++i) {
        int c = (i * 32 + lane) * 4;
        if (c < dim) r.v[i] = *reinterpret_cast<const float4*>(p + c);
      }
    } else {
      const int n = l - n_prefix;
      const float* xs = x + (size_t)(s % x_seqs) * c_in * n_points + n;
      float xv[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) xv[k] = (k < c_in) ? xs[(size_t)k * n_points] : 0.f;
#pragma unroll
      for (int i = 0; i < MAXV; ++i) {
        int c = (i * 32 + lane) * 4;
        if (c < dim) {
          float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            if (k < c_in) {
              const float4 w4 = *reinterpret_cast<const float4*>(sw + k * dim + c);  // conflict-free
              acc[0] = fmaf(xv[k], w4.x, acc[0]);
              acc[1] = fmaf(xv[k], w4.y, acc[1]);
              acc[2] = fmaf(xv[k], w4.z, acc[2]);
              acc[3] = fmaf(xv[k], w4.w, acc[3]);
            }
          }
          {
            const float4 b4 = *reinterpret_cast<const float4*>(sw + c_in * dim + c);
            acc[0] += b4.x; acc[1] += b4.y; acc[2] += b4.z; acc[3] += b4.w;
          }
          if (add_cond != nullptr) {
            float4 e = *reinterpret_cast<const float4*>(add_cond + (size_t)s * dim + c);
            acc[0] += e.x; acc[1] += e.y; acc[2] += e.z; acc[3] += e.w;
          }
          r.v[i] = make_float4(acc[0], acc[1], acc[2], acc[3]);
        }
      }
    }
    float mean, rstd;
    row_stats<MAXV>(r, dim, lane, mean, rstd, eps);
    float* hr = h + (size_t)row * dim;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      int c = (i * 32 + lane) * 4;
      if (c < dim) {
        float4 g = *reinterpret_cast<const float4*>(ln_g + c);
        float4 b = *reinterpret_cast<const float4*>(ln_b + c);
        float4 y;
        y.x = (r.v[i].x - mean) * rstd * g.x + b.x;
        y.y = (r.v[i].y - mean) * rstd * g.y + b.y;
        y.z = (r.v[i].z - mean) * rstd * g.z + b.z;
        y.w = (r.v[i].w - mean) * rstd * g.w + b.w;
        *reinterpret_cast<float4*>(hr + c) = y;
        r.v[i] = y;
      }
    }
    if (hb != nullptr) emit_bf16_stats<MAXV>(r, row, dim, lane, hb, stats);
  }
}

// The same operation with the lane's slice of the input projection, its bias and the ln_pre parameters held in REGISTERS
// (widths <= 512, CIN channels): the shared-memory form above re-reads 7 x 16 B of weights and 2 x 16 B of LayerNorm
// parameters per 16 B it writes and is bound by the shared-memory / L1 datapath (170 us at the bench shape against a
// 41 us write floor); here a row costs only FMAs, shuffles and its stores.  A warp takes chunks of EMBR_CHUNK
// consecutive rows: 16 lanes fetch the rows' point channels with coalesced loads up front, the row loop broadcasts
// them by shuffle.  Arithmetic (FMA order, statistics, LayerNorm expression) is the shared-memory kernel's, bit for bit.
constexpr int EMBR_CHUNK = 16;

template <int MAXV, int CIN>
__global__ void __launch_bounds__(128, 2) embed_tokens_reg_kernel(
    const float* __restrict__ x, int x_seqs, int n_points, const float* __restrict__ w_in,
    const float* __restrict__ b_in, const float* __restrict__ prefix, int n_prefix,
    const float* __restrict__ add_cond, const float* __restrict__ ln_g, const float* __restrict__ ln_b, float eps,
    float* __restrict__ h, int seqs, int dim, uint16_t* __restrict__ hb, float2* __restrict__ stats, int64_t n_chunks) {
  const int lane = threadIdx.x & 31;
  const int L = n_prefix + n_points;
  const int64_t total = (int64_t)seqs * L;
  float4 w[CIN][MAXV], bs[MAXV], g[MAXV], bt[MAXV];
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = (i * 32 + lane) * 4;
    const bool ok = c < dim;
#pragma unroll
    for (int k = 0; k < CIN; ++k)
      w[k][i] = ok ? make_float4(w_in[(size_t)c * CIN + k], w_in[(size_t)(c + 1) * CIN + k], w_in[(size_t)(c + 2) * CIN + k],
                                 w_in[(size_t)(c + 3) * CIN + k])
                   : make_float4(0.f, 0.f, 0.f, 0.f);
    bs[i] = ok ? *reinterpret_cast<const float4*>(b_in + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    g[i] = ok ? *reinterpret_cast<const float4*>(ln_g + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    bt[i] = ok ? *reinterpret_cast<const float4*>(ln_b + c) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t ch = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; ch < n_chunks; ch += nwarps) {
    const int64_t base = ch * EMBR_CHUNK;
    const int nrows = (int)((total - base < EMBR_CHUNK) ? (total - base) : EMBR_CHUNK);
    float xm[CIN];
    {
      int64_t mine = base + (lane & (EMBR_CHUNK - 1));
      if (mine >= total) mine = total - 1;
      const int ms = (int)(mine / L), ml = (int)(mine % L);
      const float* xs = x + (size_t)(ms % x_seqs) * CIN * n_points + (ml >= n_prefix ? ml - n_prefix : 0);
#pragma unroll
      for (int k = 0; k < CIN; ++k) xm[k] = xs[(size_t)k * n_points];
    }
    int s = (int)(base / L), l = (int)(base % L);
    for (int rr = 0; rr < nrows; ++rr) {
      const int64_t row = base + rr;
      RowRegs<MAXV> r;
      if (l < n_prefix) {
        const float* p = prefix + ((size_t)s * n_prefix + l) * dim;
#pragma unroll
        for (int i = 0; i < MAXV; ++i) {
          const int c = (i * 32 + lane) * 4;
          if (c < dim) r.v[i] = *reinterpret_cast<const float4*>(p + c);
        }
      } else {
        float xv[CIN];
#pragma unroll
        for (int k = 0; k < CIN; ++k) xv[k] = __shfl_sync(0xffffffffu, xm[k], rr);
#pragma unroll
        for (int i = 0; i < MAXV; ++i) {
          const int c = (i * 32 + lane) * 4;
          if (c < dim) {
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int k = 0; k < CIN; ++k) {
              acc[0] = fmaf(xv[k], w[k][i].x, acc[0]);
              acc[1] = fmaf(xv[k], w[k][i].y, acc[1]);
              acc[2] = fmaf(xv[k], w[k][i].z, acc[2]);
              acc[3] = fmaf(xv[k], w[k][i].w, acc[3]);
            }
            acc[0] += bs[i].x; acc[1] += bs[i].y; acc[2] += bs[i].z; acc[3] += bs[i].w;
            if (add_cond != nullptr) {
              const float4 e = *reinterpret_cast<const float4*>(add_cond + (size_t)s * dim + c);
              acc[0] += e.x; acc[1] += e.y; acc[2] += e.z; acc[3] += e.w;
            }
            r.v[i] = make_float4(acc[0], acc[1], acc[2], acc[3]);
          }
        }
      }
      float mean, rstd;
      row_stats<MAXV>(r, dim, lane, mean, rstd, eps);
      float* hr = h + (size_t)row * dim;
#pragma unroll
      for (int i = 0; i < MAXV; ++i) {
        const int c = (i * 32 + lane) * 4;
        if (c < dim) {
          float4 y;
          y.x = (r.v[i].x - mean) * rstd * g[i].x + bt[i].x;
          y.y = (r.v[i].y - mean) * rstd * g[i].y + bt[i].y;
          y.z = (r.v[i].z - mean) * rstd * g[i].z + bt[i].z;
          y.w = (r.v[i].w - mean) * rstd * g[i].w + bt[i].w;
          *reinterpret_cast<float4*>(hr + c) = y;
          r.v[i] = y;
        }
      }
      if (hb != nullptr) emit_bf16_stats<MAXV>(r, row, dim, lane, hb, stats);
      if (++l == L) { l = 0; ++s; }
    }
  }
}

// ---------------------------------------------------------------------------
// ln_post + slice + output_proj + permute (reference models/transformer.py:222-226)
// One warp per point token; the c_out dot products are reduced with shuffles and
// staged through shared memory so that the NCL store is coalesced along n.
// ---------------------------------------------------------------------------
template <int MAXV>
__global__ void __launch_bounds__(256) output_proj_kernel(
    const float* __restrict__ h, const void* __restrict__ y, int y_bf16, int seqs, int n_prefix,
    int n_points, int dim, const float* __restrict__ ln_g, const float* __restrict__ ln_b, float eps,
    const float* __restrict__ w_out, const float* __restrict__ b_out, int c_out,
    float* __restrict__ out) {
  // block = 8 warps = 8 consecutive point tokens of one sequence
  __shared__ float stage[8][33];
  int L = n_prefix + n_points;
  int blocks_per_seq = (n_points + 7) / 8;
  int s = blockIdx.x / blocks_per_seq;
  int n0 = (blockIdx.x % blocks_per_seq) * 8;
  int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int n = n0 + w;
  if (n < n_points) {
    const float* hr = h + ((size_t)s * L + n_prefix + n) * dim;
    RowRegs<MAXV> r;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      int c = (i * 32 + lane) * 4;
      if (c < dim) r.v[i] = *reinterpret_cast<const float4*>(hr + c);
    }
    // last block's MLP output (residual add folded in; h itself is not needed afterwards)
    if (y != nullptr) add_rows<MAXV>(r, y, y_bf16, ((size_t)s * L + n_prefix + n) * dim, dim, lane);
    float mean, rstd;
    row_stats<MAXV>(r, dim, lane, mean, rstd, eps);
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      int c = (i * 32 + lane) * 4;
      if (c < dim) {
        float4 g = *reinterpret_cast<const float4*>(ln_g + c);
        float4 b = *reinterpret_cast<const float4*>(ln_b + c);
        r.v[i].x = (r.v[i].x - mean) * rstd * g.x + b.x;
        r.v[i].y = (r.v[i].y - mean) * rstd * g.y + b.y;
        r.v[i].z = (r.v[i].z - mean) * rstd * g.z + b.z;
        r.v[i].w = (r.v[i].w - mean) * rstd * g.w + b.w;
      }
    }
    for (int o = 0; o < c_out; ++o) {
      const float* wr = w_out + (size_t)o * dim;
      float a = 0.f;
#pragma unroll
      for (int i = 0; i < MAXV; ++i) {
        int c = (i * 32 + lane) * 4;
        if (c < dim) {
          float4 ww = *reinterpret_cast<const float4*>(wr + c);
          a = fmaf(r.v[i].x, ww.x, a);
          a = fmaf(r.v[i].y, ww.y, a);
          a = fmaf(r.v[i].z, ww.z, a);
          a = fmaf(r.v[i].w, ww.w, a);
        }
      }
      a = warp_sum(a);
      if (lane == 0) stage[w][o] = a + b_out[o];
    }
  }
  __syncthreads();
  // coalesced-ish store: thread (o, j) writes out[s, o, n0 + j]
  for (int idx = threadIdx.x; idx < c_out * 8; idx += blockDim.x) {
    int o = idx / 8, j = idx % 8;
    if (n0 + j < n_points) out[((size_t)s * c_out + o) * n_points + n0 + j] = stage[j][o];
  }
}

// The same operation for the LayerNorm-folded forward (no pending MLP output, widths <= 512, COUT channels) with
// gamma o w_out held in registers and beta . w_out + b_out precomputed per warp:
//   out_o = rstd * sum_c (v_c - mean) (gamma_c w_oc) + (sum_c beta_c w_oc + b_o)
// The kernel above issues 36 16-byte loads per 2 KB row (h, gamma, beta, 6 weight rows) and is bound by the L1
// datapath (107 us at the bench shape against a 41 us read floor); here a row costs its own 4 loads.  A warp takes chunks
// of OPR_CHUNK consecutive point tokens, prefetches the next row while it reduces the current one, and lane j keeps the
// results of row j so that the NCL store is one coalesced segment per channel.
constexpr int OPR_CHUNK = 16;

template <int MAXV, int COUT>
__global__ void __launch_bounds__(128, 2) output_proj_reg_kernel(
    const float* __restrict__ h, int seqs, int n_prefix, int n_points, int dim, const float* __restrict__ ln_g,
    const float* __restrict__ ln_b, float eps, const float* __restrict__ w_out, const float* __restrict__ b_out,
    float* __restrict__ out, int64_t n_chunks) {
  const int lane = threadIdx.x & 31;
  const int L = n_prefix + n_points;
  const int64_t total = (int64_t)seqs * n_points;
  float4 gw[COUT][MAXV];
  float cst[COUT];
#pragma unroll
  for (int o = 0; o < COUT; ++o) {
    float a = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int c = (i * 32 + lane) * 4;
      gw[o][i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < dim) {
        const float4 g = *reinterpret_cast<const float4*>(ln_g + c), b = *reinterpret_cast<const float4*>(ln_b + c);
        const float4 w = *reinterpret_cast<const float4*>(w_out + (size_t)o * dim + c);
        gw[o][i] = make_float4(g.x * w.x, g.y * w.y, g.z * w.z, g.w * w.w);
        a += (b.x * w.x + b.y * w.y) + (b.z * w.z + b.w * w.w);
      }
    }
    cst[o] = warp_sum(a) + b_out[o];
  }
  auto load_row = [&](RowRegs<MAXV>& r, int s, int n) {
    const float* hr = h + ((size_t)s * L + n_prefix + n) * dim;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int c = (i * 32 + lane) * 4;
      r.v[i] = (c < dim) ? *reinterpret_cast<const float4*>(hr + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t ch = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; ch < n_chunks; ch += nwarps) {
    const int64_t base = ch * OPR_CHUNK;
    const int nrows = (int)((total - base < OPR_CHUNK) ? (total - base) : OPR_CHUNK);
    int s = (int)(base / n_points), n = (int)(base % n_points);
    RowRegs<MAXV> cur, nxt;
    load_row(cur, s, n);
    float res[COUT];
#pragma unroll
    for (int o = 0; o < COUT; ++o) res[o] = 0.f;
    for (int rr = 0; rr < nrows; ++rr) {
      if (++n == n_points) { n = 0; ++s; }
      if (rr + 1 < nrows) load_row(nxt, s, n);
      float mean, rstd;
      row_stats<MAXV>(cur, dim, lane, mean, rstd, eps);
      float a[COUT];
#pragma unroll
      for (int o = 0; o < COUT; ++o) a[o] = 0.f;
#pragma unroll
      for (int i = 0; i < MAXV; ++i) {
        const float d0 = cur.v[i].x - mean, d1 = cur.v[i].y - mean, d2 = cur.v[i].z - mean, d3 = cur.v[i].w - mean;
#pragma unroll
        for (int o = 0; o < COUT; ++o) {   // gw is zero beyond dim
          a[o] = fmaf(d0, gw[o][i].x, a[o]);
          a[o] = fmaf(d1, gw[o][i].y, a[o]);
          a[o] = fmaf(d2, gw[o][i].z, a[o]);
          a[o] = fmaf(d3, gw[o][i].w, a[o]);
        }
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
        for (int o = 0; o < COUT; ++o) a[o] += __shfl_xor_sync(0xffffffffu, a[o], off);
      }
#pragma unroll
      for (int o = 0; o < COUT; ++o) {
        const float val = fmaf(a[o], rstd, cst[o]);
        if (lane == rr) res[o] = val;
      }
#pragma unroll
      for (int i = 0; i < MAXV; ++i) cur.v[i] = nxt.v[i];
    }
    if (lane < nrows) {
      const int64_t mine = base + lane;
      const int ms = (int)(mine / n_points), mn = (int)(mine % n_points);
#pragma unroll
      for (int o = 0; o < COUT; ++o) out[((size_t)ms * COUT + o) * n_points + mn] = res[o];
    }
  }
}

}  // namespace pcd

using namespace pcd;

#define DISPATCH_MAXV(dim, ...)                              \
  do {                                                       \
    int nv__ = ((dim) + 127) / 128;                          \
    if (nv__ <= 1) { constexpr int MAXV = 1; __VA_ARGS__; }  \
    else if (nv__ <= 2) { constexpr int MAXV = 2; __VA_ARGS__; } \
    else if (nv__ <= 4) { constexpr int MAXV = 4; __VA_ARGS__; } \
    else if (nv__ <= 8) { constexpr int MAXV = 8; __VA_ARGS__; } \
    else { constexpr int MAXV = 16; __VA_ARGS__; }           \
  } while (0)

extern "C" int pcd_timestep_embed(const float* t, const float* freqs, int batch, int dim,
                                  float* out, int ld_out, void* stream) {
  PCD_CHECK_ARG(batch > 0 && dim >= 2 && ld_out >= dim, "timestep_embed: bad shape");
  int n = batch * (dim / 2);
  timestep_embed_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(t, freqs, batch, dim, out, ld_out);
  PCD_CHECK_LAUNCH("timestep_embed");
  return PCD_OK;
}

static int launch_layernorm(float* x, int ldx, const void* y, int ldy, int y_precision, const float* gamma,
                            const float* beta, void* out, int ld_out, int out_precision, int rows, int dim,
                            float eps, void* stream, const char* what) {
  PCD_CHECK_ARG(rows > 0 && dim > 0 && dim % 4 == 0 && dim <= 2048, "%s: dim must be a multiple of 4 and <= 2048 (got %d)", what, dim);
  PCD_CHECK_ARG(ldx % 4 == 0 && ld_out % 4 == 0 && ldy % 4 == 0, "%s: leading dims must be multiples of 4", what);
  dim3 grid(ceil_div(rows, 8)), block(256);
  cudaStream_t st = (cudaStream_t)stream;
  const int yb = y_precision == PCD_BF16;
  if (out_precision == PCD_BF16) {
    DISPATCH_MAXV(dim, (layernorm_kernel<MAXV, true><<<grid, block, 0, st>>>(x, ldx, y, ldy, yb, gamma, beta, out, ld_out, rows, dim, eps)));
  } else {
    DISPATCH_MAXV(dim, (layernorm_kernel<MAXV, false><<<grid, block, 0, st>>>(x, ldx, y, ldy, yb, gamma, beta, out, ld_out, rows, dim, eps)));
  }
  PCD_CHECK_LAUNCH(what);
  return PCD_OK;
}

extern "C" int pcd_layernorm(const float* x, int ldx, const float* gamma, const float* beta,
                             void* out, int ld_out, int out_precision, int rows, int dim, float eps,
                             void* stream) {
  return launch_layernorm(const_cast<float*>(x), ldx, nullptr, 0, PCD_F32, gamma, beta, out, ld_out, out_precision,
                          rows, dim, eps, stream, "layernorm");
}

extern "C" int pcd_add_layernorm(float* h, int ldh, const void* y, int ldy, int y_precision,
                                 const float* gamma, const float* beta, void* out, int ld_out,
                                 int out_precision, int rows, int dim, float eps, void* stream) {
  PCD_CHECK_ARG(y != nullptr, "add_layernorm: y missing");
  return launch_layernorm(h, ldh, y, ldy, y_precision, gamma, beta, out, ld_out, out_precision, rows, dim, eps,
                          stream, "add_layernorm");
}

extern "C" int pcd_add_f32(const float* a, const float* b, float* out, int64_t n, void* stream) {
  PCD_CHECK_ARG(a != nullptr && b != nullptr && out != nullptr && n > 0, "add_f32: bad arguments");
  PCD_CHECK_ARG(((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(out)) & 15) == 0,
                "add_f32: operands must be 16-byte aligned");
  const int64_t n4 = n / 4;
  add_f32_kernel<<<(unsigned)ceil_div64(n4 + 1, 256), 256, 0, (cudaStream_t)stream>>>(a, b, out, n4, n);
  PCD_CHECK_LAUNCH("add_f32");
  return PCD_OK;
}

extern "C" int pcd_cast_rowstats(const float* x, int ldx, uint16_t* out, int ld_out, float* stats, int rows,
                                 int dim, void* stream) {
  PCD_CHECK_ARG(x != nullptr && out != nullptr && stats != nullptr, "cast_rowstats: null argument");
  PCD_CHECK_ARG(rows > 0 && dim > 0 && dim % 128 == 0 && dim <= 2048, "cast_rowstats: dim must be a multiple of 128, <= 2048 (got %d)", dim);
  PCD_CHECK_ARG(ldx % 4 == 0 && ld_out % 4 == 0, "cast_rowstats: leading dims must be multiples of 4");
  dim3 grid(ceil_div(rows, 8)), block(256);
  cudaStream_t st = (cudaStream_t)stream;
  DISPATCH_MAXV(dim, (cast_rowstats_kernel<MAXV><<<grid, block, 0, st>>>(x, ldx, out, ld_out, reinterpret_cast<float2*>(stats), rows, dim)));
  PCD_CHECK_LAUNCH("cast_rowstats");
  return PCD_OK;
}

extern "C" int pcd_embed_tokens(const float* x, int x_seqs, int c_in, int n_points,
                                const float* w_in, const float* b_in, const float* prefix,
                                int n_prefix, const float* add_cond, const float* ln_g,
                                const float* ln_b, float eps, float* h, int seqs, int dim,
                                uint16_t* h_bf16, float* stats, void* stream) {
  PCD_CHECK_ARG((h_bf16 == nullptr) == (stats == nullptr), "embed_tokens: h_bf16 and stats come together");
  PCD_CHECK_ARG(h_bf16 == nullptr || dim % 128 == 0, "embed_tokens: the bf16 copy + statistics need dim %% 128 == 0");
  PCD_CHECK_ARG(seqs > 0 && x_seqs > 0 && seqs % x_seqs == 0, "embed_tokens: seqs must be a multiple of x_seqs");
  PCD_CHECK_ARG(c_in >= 1 && c_in <= 8, "embed_tokens: c_in must be in [1,8] (got %d)", c_in);
  PCD_CHECK_ARG(dim % 4 == 0 && dim <= 2048, "embed_tokens: bad width %d", dim);
  PCD_CHECK_ARG(n_prefix >= 0 && (n_prefix == 0 || prefix != nullptr), "embed_tokens: prefix missing");
  int64_t rows = (int64_t)seqs * (n_prefix + n_points);
  cudaStream_t st = (cudaStream_t)stream;
  if (dim <= 512 && (c_in == 3 || c_in == 6)) {
    // weights in registers; persistent grid, two 128-thread blocks per SM
    const int64_t n_chunks = ceil_div64(rows, EMBR_CHUNK);
    int64_t blocks = ceil_div64(n_chunks, 4);
    if (blocks > 2 * (int64_t)num_sms()) blocks = 2 * num_sms();
#define PCD_EMBR(CIN) DISPATCH_MAXV(dim, (embed_tokens_reg_kernel<(MAXV <= 4 ? MAXV : 4), CIN><<<(unsigned)blocks, 128, 0, st>>>( \
      x, x_seqs, n_points, w_in, b_in, prefix, n_prefix, add_cond, ln_g, ln_b, eps, h, seqs, dim, h_bf16, reinterpret_cast<float2*>(stats), n_chunks)))
    if (c_in == 3) { PCD_EMBR(3); } else { PCD_EMBR(6); }
#undef PCD_EMBR
    PCD_CHECK_LAUNCH("embed_tokens");
    return PCD_OK;
  }
  dim3 grid((unsigned)ceil_div64(rows, 8 * EMB_ROWS_PER_WARP)), block(256);
  const size_t smem = (size_t)dim * (c_in + 1) * sizeof(float);  // <= 2048 * 9 * 4 = 72 KB
  if (smem > 48 * 1024) {
    // width 2048 (base1B) with >= 5 channels: opt in to more than 48 KB of dynamic shared memory
    cudaError_t e = cudaSuccess;
    DISPATCH_MAXV(dim, (e = cudaFuncSetAttribute(embed_tokens_kernel<MAXV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)));
    if (e != cudaSuccess) {
      set_error("embed_tokens: width %d with %d channels needs %zu B of shared memory: %s", dim, c_in, smem, cudaGetErrorString(e));
      return PCD_ERR_CUDA;
    }
  }
  DISPATCH_MAXV(dim, (embed_tokens_kernel<MAXV><<<grid, block, smem, st>>>(x, x_seqs, c_in, n_points, w_in, b_in, prefix, n_prefix, add_cond, ln_g, ln_b, eps, h, seqs, dim, h_bf16, reinterpret_cast<float2*>(stats))));
  PCD_CHECK_LAUNCH("embed_tokens");
  return PCD_OK;
}

extern "C" int pcd_output_proj(const float* h, const void* y, int y_precision, int seqs, int n_prefix,
                               int n_points, int dim, const float* ln_g, const float* ln_b, float eps,
                               const float* w_out, const float* b_out, int c_out, float* out,
                               void* stream) {
  PCD_CHECK_ARG(seqs > 0 && n_points > 0 && dim % 4 == 0 && dim <= 2048, "output_proj: bad shape");
  PCD_CHECK_ARG(c_out >= 1 && c_out <= 32, "output_proj: c_out must be in [1,32] (got %d)", c_out);
  cudaStream_t st = (cudaStream_t)stream;
  if (y == nullptr && dim <= 512 && (c_out == 3 || c_out == 6)) {
    const int64_t n_chunks = ceil_div64((int64_t)seqs * n_points, OPR_CHUNK);
    int64_t blocks = ceil_div64(n_chunks, 4);
    if (blocks > 2 * (int64_t)num_sms()) blocks = 2 * num_sms();
#define PCD_OPR(COUT) DISPATCH_MAXV(dim, (output_proj_reg_kernel<(MAXV <= 4 ? MAXV : 4), COUT><<<(unsigned)blocks, 128, 0, st>>>( \
      h, seqs, n_prefix, n_points, dim, ln_g, ln_b, eps, w_out, b_out, out, n_chunks)))
    if (c_out == 3) { PCD_OPR(3); } else { PCD_OPR(6); }
#undef PCD_OPR
    PCD_CHECK_LAUNCH("output_proj");
    return PCD_OK;
  }
  dim3 grid(seqs * ceil_div(n_points, 8)), block(256);
  DISPATCH_MAXV(dim, (output_proj_kernel<MAXV><<<grid, block, 0, st>>>(h, y, y_precision == PCD_BF16, seqs, n_prefix, n_points, dim, ln_g, ln_b, eps, w_out, b_out, c_out, out)));
  PCD_CHECK_LAUNCH("output_proj");
  return PCD_OK;
}
