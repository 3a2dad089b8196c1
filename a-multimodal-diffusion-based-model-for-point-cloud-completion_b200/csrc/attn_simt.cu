// fp32 CUDA-core flash attention (head dim 64, or 32 for the TwoStream denoiser) for the 1e-4 parity mode.
// Online softmax in fp32 with exact expf; q and k are scaled separately on load
// exactly like the reference (q*hd^-1/4, k*hd^-1/4, transformer.py:76-80), optional
// 3-axis rotary on head dims 0..5 (rotaryencoderpcd.py:6-27).
#include "common.cuh"

namespace pcd {

constexpr int AQ = 64, AKV = 64, APAD = 4;

template <int HD>
struct AttnSmem {
  float Qt[HD][AQ + APAD];    // [k][query]
  float Kt[HD][AKV + APAD];   // [k][key]
  float Vs[AKV][HD + APAD];   // [key][c]
  float Pt[AKV][AQ + APAD];   // [key][query]
};

__device__ __forceinline__ void rope6(float* x /*6 values: dims 0..5*/, const float* coord) {
  float s0, c0, s1, c1, s2, c2;
  sincosf(coord[0] * 3.14159265358979323846f, &s0, &c0);
  sincosf(coord[1] * 3.14159265358979323846f, &s1, &c1);
  sincosf(coord[2] * 3.14159265358979323846f, &s2, &c2);
  float e0 = x[0], o0 = x[1], e1 = x[2], o1 = x[3], e2 = x[4], o2 = x[5];
  x[0] = e0 * c0 - o0 * s0;
  x[1] = e1 * c1 - o1 * s1;
  x[2] = e2 * c2 - o2 * s2;
  x[3] = e0 * s0 + o0 * c0;
  x[4] = e1 * s1 + o1 * c1;
  x[5] = e2 * s2 + o2 * c2;
}

template <int HD>
__global__ void __launch_bounds__(256) attn_f32_kernel(
    const float* __restrict__ q, int64_t q_bs, int64_t q_ls, int64_t q_hs,
    const float* __restrict__ k, int64_t k_bs, int64_t k_ls, int64_t k_hs,
    const float* __restrict__ v, int64_t v_bs, int64_t v_ls, int64_t v_hs,
    float* __restrict__ out, int64_t o_bs, int64_t o_ls, int len_q, int len_kv, float q_scale,
    float k_scale, const float* __restrict__ rope) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  AttnSmem<HD>& sm = *reinterpret_cast<AttnSmem<HD>*>(smem_raw);
  constexpr int OC = HD / 16;  // output columns per thread (16 threads span a head)
  int tid = threadIdx.x;
  int q0 = blockIdx.x * AQ, h = blockIdx.y, b = blockIdx.z;
  const float* qb = q + b * q_bs + h * q_hs;
  const float* kb = k + b * k_bs + h * k_hs;
  const float* vb = v + b * v_bs + h * v_hs;

  // ---- load Q tile (scaled, transposed) ----
  for (int i = tid; i < AQ * (HD / 4); i += 256) {
    int r = i / (HD / 4), c4 = (i % (HD / 4)) * 4;
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q0 + r < len_q) t = *reinterpret_cast<const float4*>(qb + (int64_t)(q0 + r) * q_ls + c4);
    sm.Qt[c4 + 0][r] = t.x; sm.Qt[c4 + 1][r] = t.y; sm.Qt[c4 + 2][r] = t.z; sm.Qt[c4 + 3][r] = t.w;
  }
  __syncthreads();
  if (tid < AQ) {
    float x[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) x[j] = sm.Qt[j][tid];
    if (rope != nullptr && q0 + tid < len_q) rope6(x, rope + ((int64_t)b * len_q + q0 + tid) * 3);
#pragma unroll
    for (int j = 0; j < 6; ++j) sm.Qt[j][tid] = x[j];
  }
  __syncthreads();
  for (int i = tid; i < HD * AQ; i += 256) sm.Qt[i / AQ][i % AQ] *= q_scale;

  int tx = tid & 15, ty = tid >> 4;
  float m_run[4], l_run[4], acc[4][OC];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    m_run[i] = -INFINITY;
    l_run[i] = 0.f;
#pragma unroll
    for (int j = 0; j < OC; ++j) acc[i][j] = 0.f;
  }

  for (int kv0 = 0; kv0 < len_kv; kv0 += AKV) {
    __syncthreads();  // previous tile fully consumed (also orders the Q scaling above)
    for (int i = tid; i < AKV * (HD / 4); i += 256) {
      int r = i / (HD / 4), c4 = (i % (HD / 4)) * 4;
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f), u = t;
      if (kv0 + r < len_kv) {
        t = *reinterpret_cast<const float4*>(kb + (int64_t)(kv0 + r) * k_ls + c4);
        u = *reinterpret_cast<const float4*>(vb + (int64_t)(kv0 + r) * v_ls + c4);
      }
      sm.Kt[c4 + 0][r] = t.x; sm.Kt[c4 + 1][r] = t.y; sm.Kt[c4 + 2][r] = t.z; sm.Kt[c4 + 3][r] = t.w;
      *reinterpret_cast<float4*>(&sm.Vs[r][c4]) = u;
    }
    __syncthreads();
    if (tid < AKV) {
      float x[6];
#pragma unroll
      for (int j = 0; j < 6; ++j) x[j] = sm.Kt[j][tid];
      if (rope != nullptr && kv0 + tid < len_kv) rope6(x, rope + ((int64_t)b * len_kv + kv0 + tid) * 3);
#pragma unroll
      for (int j = 0; j < 6; ++j) sm.Kt[j][tid] = x[j] * k_scale;
    }
    for (int i = tid; i < (HD - 6) * AKV; i += 256) sm.Kt[6 + i / AKV][i % AKV] *= k_scale;
    __syncthreads();

    // ---- S = Q K^T (4x4 per thread) ----
    float s[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
#pragma unroll 8
    for (int kk = 0; kk < HD; ++kk) {
      float4 a = *reinterpret_cast<const float4*>(&sm.Qt[kk][ty * 4]);
      float4 bb = *reinterpret_cast<const float4*>(&sm.Kt[kk][tx * 4]);
      float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s[i][j] = fmaf(av[i], bv[j], s[i][j]);
    }
    // ---- online softmax (row = 16 consecutive lanes) ----
    float alpha[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (kv0 + tx * 4 + j >= len_kv) s[i][j] = -INFINITY;
        mx = fmaxf(mx, s[i][j]);
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      float m_new = fmaxf(m_run[i], mx);
      alpha[i] = expf(m_run[i] - m_new);
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        s[i][j] = expf(s[i][j] - m_new);
        sum += s[i][j];
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      l_run[i] = l_run[i] * alpha[i] + sum;
      m_run[i] = m_new;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
      *reinterpret_cast<float4*>(&sm.Pt[tx * 4 + j][ty * 4]) = make_float4(s[0][j], s[1][j], s[2][j], s[3][j]);
    __syncthreads();
    // ---- O = O*alpha + P V ----
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < OC; ++j) acc[i][j] *= alpha[i];
#pragma unroll 8
    for (int kk = 0; kk < AKV; ++kk) {
      float4 a = *reinterpret_cast<const float4*>(&sm.Pt[kk][ty * 4]);
      float av[4] = {a.x, a.y, a.z, a.w}, bv[OC];
#pragma unroll
      for (int j = 0; j < OC; ++j) bv[j] = sm.Vs[kk][tx * OC + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < OC; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int r = q0 + ty * 4 + i;
    if (r < len_q) {
      float inv = 1.f / l_run[i];
      float* o = out + b * o_bs + (int64_t)r * o_ls + h * HD + tx * OC;
#pragma unroll
      for (int j = 0; j < OC; ++j) o[j] = acc[i][j] * inv;
    }
  }
}

// In-place 3-axis RoPE on head dims 0..5 of a bf16 q or k operand (reference rotaryencoderpcd.py:6-27):
// the tensor-core attention kernel consumes its tiles straight from TMA, so in bf16 mode the rotation
// is applied to the projection output before it (12 bytes per token and head; fp32 arithmetic).
__global__ void rope_bf16_kernel(uint16_t* __restrict__ x, int64_t bs, int64_t ls, int64_t hs,
                                 const float* __restrict__ coords, int batch, int heads, int len) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)batch * len * heads;
  if (idx >= total) return;
  const int h = (int)(idx % heads);
  const int64_t bl = idx / heads;
  const int l = (int)(bl % len);
  const int b = (int)(bl / len);
  uint16_t* p = x + b * bs + (int64_t)l * ls + (int64_t)h * hs;
  float v[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) v[i] = __uint_as_float((uint32_t)p[i] << 16);
  rope6(v, coords + ((int64_t)b * len + l) * 3);
#pragma unroll
  for (int i = 0; i < 6; i += 2) *reinterpret_cast<uint32_t*>(p + i) = pack_bf16x2(v[i], v[i + 1]);
}

int launch_rope_bf16(uint16_t* x, int64_t bs, int64_t ls, int64_t hs, const float* coords, int batch, int heads,
                     int len, cudaStream_t st) {
  const int64_t total = (int64_t)batch * len * heads;
  rope_bf16_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, st>>>(x, bs, ls, hs, coords, batch, heads, len);
  PCD_CHECK_LAUNCH("rope_bf16");
  return PCD_OK;
}

template <int HD>
static int launch_attention_f32_hd(const pcd_attn_operand* q, const pcd_attn_operand* k, const pcd_attn_operand* v,
                                   float* out, int64_t o_bs, int64_t o_ls, int batch, int heads, int len_q, int len_kv,
                                   float q_scale, float k_scale, const float* rope, cudaStream_t st) {
  {  // per device, not per process
    cudaError_t e = cudaFuncSetAttribute(attn_f32_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)sizeof(AttnSmem<HD>));
    if (e != cudaSuccess) {
      set_error("attention_f32: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return PCD_ERR_CUDA;
    }
  }
  dim3 grid(ceil_div(len_q, AQ), heads, batch);
  attn_f32_kernel<HD><<<grid, 256, sizeof(AttnSmem<HD>), st>>>(
      (const float*)q->ptr, q->batch_stride, q->row_stride, q->head_stride,
      (const float*)k->ptr, k->batch_stride, k->row_stride, k->head_stride,
      (const float*)v->ptr, v->batch_stride, v->row_stride, v->head_stride, out, o_bs, o_ls, len_q,
      len_kv, q_scale, k_scale, rope);
  PCD_CHECK_LAUNCH("attention_f32");
  return PCD_OK;
}

int launch_attention_f32(const pcd_attn_operand* q, const pcd_attn_operand* k,
                         const pcd_attn_operand* v, float* out, int64_t o_bs, int64_t o_ls,
                         int batch, int heads, int len_q, int len_kv, float q_scale, float k_scale,
                         const float* rope, int head_dim, cudaStream_t st) {
  if (head_dim == 32)
    return launch_attention_f32_hd<32>(q, k, v, out, o_bs, o_ls, batch, heads, len_q, len_kv, q_scale, k_scale, rope, st);
  return launch_attention_f32_hd<64>(q, k, v, out, o_bs, o_ls, batch, heads, len_q, len_kv, q_scale, k_scale, rope, st);
}

}  // namespace pcd
