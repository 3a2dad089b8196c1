// Point-cloud utilities either side of the sampler: farthest-point
// sampling, nearest-point search and the F-score of the evaluation loop, without the reference's
// [B, N, M, 3] difference tensor (models/util.py:213-214: 8192 x 8192 x 3 floats = 805 MB per cloud).
#include "common.cuh"

namespace pcd {

// ---------------------------------------------------------------------------
// Farthest-point sampling (reference util/point_cloud.py:82-118; evaluation.py:40-48 uses the same
// greedy scheme on batches).  One CTA per cloud; every thread keeps the running minimum distance of its
// points in registers; each round: block arg-max (first maximum wins, like np.argmax), broadcast the
// winner's coordinates, update.  Distances use the reference's |a|^2 + |b|^2 - 2 a.b form with
// un-fused fp32 operations.
// ---------------------------------------------------------------------------
constexpr int FPS_THREADS = 512, FPS_MAX_PER_THREAD = 16;  // up to 8192 points per cloud (x, y, z, |p|^2, dist in registers)

__global__ void __launch_bounds__(FPS_THREADS) fps_kernel(const float* __restrict__ pts, int n, int n_samples,
                                                          const int* __restrict__ init_idx,
                                                          long long* __restrict__ out) {
  __shared__ float red_v[FPS_THREADS / 32];
  __shared__ int red_i[FPS_THREADS / 32];
  __shared__ int s_win;
  const float* p = pts + (size_t)blockIdx.x * n * 3;
  long long* o = out + (size_t)blockIdx.x * n_samples;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  float x[FPS_MAX_PER_THREAD], y[FPS_MAX_PER_THREAD], z[FPS_MAX_PER_THREAD], sq[FPS_MAX_PER_THREAD],
      dist[FPS_MAX_PER_THREAD];
#pragma unroll
  for (int k = 0; k < FPS_MAX_PER_THREAD; ++k) {
    const int i = k * FPS_THREADS + tid;
    if (i < n) {
      x[k] = p[3 * i]; y[k] = p[3 * i + 1]; z[k] = p[3 * i + 2];
      sq[k] = __fadd_rn(__fadd_rn(__fmul_rn(x[k], x[k]), __fmul_rn(y[k], y[k])), __fmul_rn(z[k], z[k]));
    }
    dist[k] = 3.4e38f;
  }
  int win = min(max(init_idx[blockIdx.x], 0), n - 1);  // a caller-supplied index never reads outside the cloud
  for (int s = 0; s < n_samples; ++s) {
    if (tid == 0) o[s] = win;
    if (s + 1 == n_samples) break;
    const float wx = p[3 * win], wy = p[3 * win + 1], wz = p[3 * win + 2];
    const float wsq = __fadd_rn(__fadd_rn(__fmul_rn(wx, wx), __fmul_rn(wy, wy)), __fmul_rn(wz, wz));
    float best = -3.4e38f;
    int best_i = 0x7fffffff;
#pragma unroll
    for (int k = 0; k < FPS_MAX_PER_THREAD; ++k) {
      const int i = k * FPS_THREADS + tid;
      if (i < n) {
        const float dot = __fadd_rn(__fadd_rn(__fmul_rn(x[k], wx), __fmul_rn(y[k], wy)), __fmul_rn(z[k], wz));
        const float d = __fsub_rn(__fadd_rn(sq[k], wsq), __fmul_rn(2.f, dot));
        dist[k] = fminf(dist[k], d);
        if (dist[k] > best) {  // ascending i within the thread: strict > keeps the first maximum
          best = dist[k];
          best_i = i;
        }
      }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, off);
      const int oi = __shfl_xor_sync(0xffffffffu, best_i, off);
      if (ov > best || (ov == best && oi < best_i)) { best = ov; best_i = oi; }
    }
    if (lane == 0) { red_v[wid] = best; red_i[wid] = best_i; }
    __syncthreads();
    if (wid == 0) {
      best = lane < FPS_THREADS / 32 ? red_v[lane] : -3.4e38f;
      best_i = lane < FPS_THREADS / 32 ? red_i[lane] : 0x7fffffff;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, off);
        const int oi = __shfl_xor_sync(0xffffffffu, best_i, off);
        if (ov > best || (ov == best && oi < best_i)) { best = ov; best_i = oi; }
      }
      if (lane == 0) s_win = best_i;
    }
    __syncthreads();
    win = s_win;
  }
}

// Clouds of more than 8192 points (e.g. the 16384-point complete clouds of the MVP dataset): the same scheme with the
// running minimum distances in a caller-supplied global workspace [batch, n] and the coordinates re-read every round.
__global__ void __launch_bounds__(FPS_THREADS) fps_large_kernel(const float* __restrict__ pts, int n, int n_samples,
                                                                const int* __restrict__ init_idx, float* __restrict__ ws,
                                                                long long* __restrict__ out) {
  __shared__ float red_v[FPS_THREADS / 32];
  __shared__ int red_i[FPS_THREADS / 32];
  __shared__ int s_win;
  const float* p = pts + (size_t)blockIdx.x * n * 3;
  float* dist = ws + (size_t)blockIdx.x * n;
  long long* o = out + (size_t)blockIdx.x * n_samples;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  for (int i = tid; i < n; i += FPS_THREADS) dist[i] = 3.4e38f;
  int win = min(max(init_idx[blockIdx.x], 0), n - 1);
  for (int s = 0; s < n_samples; ++s) {
    if (tid == 0) o[s] = win;
    if (s + 1 == n_samples) break;
    const float wx = p[3 * win], wy = p[3 * win + 1], wz = p[3 * win + 2];
    const float wsq = __fadd_rn(__fadd_rn(__fmul_rn(wx, wx), __fmul_rn(wy, wy)), __fmul_rn(wz, wz));
    float best = -3.4e38f;
    int best_i = 0x7fffffff;
    for (int i = tid; i < n; i += FPS_THREADS) {
      const float x = p[3 * i], y = p[3 * i + 1], z = p[3 * i + 2];
      const float sq = __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
      const float dot = __fadd_rn(__fadd_rn(__fmul_rn(x, wx), __fmul_rn(y, wy)), __fmul_rn(z, wz));
      const float d = fminf(dist[i], __fsub_rn(__fadd_rn(sq, wsq), __fmul_rn(2.f, dot)));
      dist[i] = d;
      if (d > best) {  // ascending i within the thread: strict > keeps the first maximum
        best = d;
        best_i = i;
      }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, off);
      const int oi = __shfl_xor_sync(0xffffffffu, best_i, off);
      if (ov > best || (ov == best && oi < best_i)) { best = ov; best_i = oi; }
    }
    if (lane == 0) { red_v[wid] = best; red_i[wid] = best_i; }
    __syncthreads();
    if (wid == 0) {
      best = lane < FPS_THREADS / 32 ? red_v[lane] : -3.4e38f;
      best_i = lane < FPS_THREADS / 32 ? red_i[lane] : 0x7fffffff;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, off);
        const int oi = __shfl_xor_sync(0xffffffffu, best_i, off);
        if (ov > best || (ov == best && oi < best_i)) { best = ov; best_i = oi; }
      }
      if (lane == 0) s_win = best_i;
    }
    __syncthreads();
    win = s_win;
  }
}

// ---------------------------------------------------------------------------
// Nearest point of cloud B for every point of A ([batch, n, 3] row-major): arg-min index (first minimum,
// like np.argmin) and / or squared distance.  form == 0: (a-b)^2 summed (models/util.py:213-214);
// form == 1: |a|^2 + |b|^2 - 2 a.b (util/point_cloud.py:159-163).  B is tiled through shared memory.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) nn_rowmajor_kernel(const float* __restrict__ a, int na,
                                                          const float* __restrict__ b, int nb, int form,
                                                          float* __restrict__ out_d2, long long* __restrict__ out_idx) {
  __shared__ float sb[4][256];
  const int bi = blockIdx.y;
  const float* pa = a + (size_t)bi * na * 3;
  const float* pb = b + (size_t)bi * nb * 3;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float ax = 0.f, ay = 0.f, az = 0.f;
  if (i < na) { ax = pa[3 * i]; ay = pa[3 * i + 1]; az = pa[3 * i + 2]; }
  const float asq = __fadd_rn(__fadd_rn(__fmul_rn(ax, ax), __fmul_rn(ay, ay)), __fmul_rn(az, az));
  float best = 3.4e38f;
  int best_j = 0;
  for (int j0 = 0; j0 < nb; j0 += 256) {
    const int j = j0 + threadIdx.x;
    __syncthreads();
    if (j < nb) {
      const float bx = pb[3 * j], by = pb[3 * j + 1], bz = pb[3 * j + 2];
      sb[0][threadIdx.x] = bx; sb[1][threadIdx.x] = by; sb[2][threadIdx.x] = bz;
      sb[3][threadIdx.x] = __fadd_rn(__fadd_rn(__fmul_rn(bx, bx), __fmul_rn(by, by)), __fmul_rn(bz, bz));
    }
    __syncthreads();
    const int m = min(256, nb - j0);
    for (int k = 0; k < m; ++k) {
      float d;
      if (form == 0) {
        const float dx = __fsub_rn(ax, sb[0][k]), dy = __fsub_rn(ay, sb[1][k]), dz = __fsub_rn(az, sb[2][k]);
        d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
      } else {
        const float dot = __fadd_rn(__fadd_rn(__fmul_rn(ax, sb[0][k]), __fmul_rn(ay, sb[1][k])), __fmul_rn(az, sb[2][k]));
        d = __fsub_rn(__fadd_rn(sb[3][k], asq), __fmul_rn(2.f, dot));
      }
      if (d < best) { best = d; best_j = j0 + k; }
    }
  }
  if (i < na) {
    if (out_d2) out_d2[(size_t)bi * na + i] = best;
    if (out_idx) out_idx[(size_t)bi * na + i] = best_j;
  }
}

// precision / recall / F-score from the two nearest-neighbour distance arrays (models/util.py:216-229)
__global__ void fscore_reduce_kernel(const float* __restrict__ d1, int n1, const float* __restrict__ d2, int n2,
                                     float threshold, int squared, float* __restrict__ out /*[3, batch]*/, int batch) {
  __shared__ float red[2][32];
  const int b = blockIdx.x;
  float c1 = 0.f, c2 = 0.f;
  for (int i = threadIdx.x; i < n1; i += blockDim.x) {
    const float v = d1[(size_t)b * n1 + i];
    c1 += ((squared ? v : sqrtf(v)) < threshold) ? 1.f : 0.f;
  }
  for (int i = threadIdx.x; i < n2; i += blockDim.x) {
    const float v = d2[(size_t)b * n2 + i];
    c2 += ((squared ? v : sqrtf(v)) < threshold) ? 1.f : 0.f;
  }
  c1 = warp_sum(c1); c2 = warp_sum(c2);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = c1; red[1][threadIdx.x >> 5] = c2; }
  __syncthreads();
  if (threadIdx.x < 32) {
    const int nw = blockDim.x >> 5;
    c1 = threadIdx.x < nw ? red[0][threadIdx.x] : 0.f;
    c2 = threadIdx.x < nw ? red[1][threadIdx.x] : 0.f;
    c1 = warp_sum(c1); c2 = warp_sum(c2);
    if (threadIdx.x == 0) {
      const float p = c1 / (float)n1, r = c2 / (float)n2;
      out[b] = 2.f * p * r / (p + r + 1e-8f);
      out[batch + b] = p;
      out[2 * batch + b] = r;
    }
  }
}

}  // namespace pcd

using namespace pcd;

extern "C" int pcd_farthest_point_sample(const float* points, int batch, int n, int n_samples, const int* init_idx,
                                         long long* out_idx, float* workspace, void* stream) {
  PCD_CHECK_ARG(points != nullptr && init_idx != nullptr && out_idx != nullptr, "farthest_point_sample: null argument");
  PCD_CHECK_ARG(batch > 0 && n > 0 && n_samples > 0 && n_samples <= n, "farthest_point_sample: need 0 < n_samples <= n");
  if (n > FPS_THREADS * FPS_MAX_PER_THREAD) {
    PCD_CHECK_ARG(workspace != nullptr, "farthest_point_sample: clouds of more than %d points need a [batch, n] float workspace",
                  FPS_THREADS * FPS_MAX_PER_THREAD);
    fps_large_kernel<<<batch, FPS_THREADS, 0, (cudaStream_t)stream>>>(points, n, n_samples, init_idx, workspace, out_idx);
  } else
    fps_kernel<<<batch, FPS_THREADS, 0, (cudaStream_t)stream>>>(points, n, n_samples, init_idx, out_idx);
  PCD_CHECK_LAUNCH("farthest_point_sample");
  return PCD_OK;
}

extern "C" int pcd_nearest_points(const float* a, int na, const float* b, int nb, int batch, int form, float* out_d2,
                                  long long* out_idx, void* stream) {
  PCD_CHECK_ARG(a != nullptr && b != nullptr && (out_d2 != nullptr || out_idx != nullptr), "nearest_points: null argument");
  PCD_CHECK_ARG(batch > 0 && batch <= 65535 && na > 0 && nb > 0 && (form == 0 || form == 1), "nearest_points: bad shape");
  nn_rowmajor_kernel<<<dim3(ceil_div(na, 256), batch), 256, 0, (cudaStream_t)stream>>>(a, na, b, nb, form, out_d2, out_idx);
  PCD_CHECK_LAUNCH("nearest_points");
  return PCD_OK;
}

extern "C" int pcd_fscore(const float* pred, int n, const float* gt, int m, int batch, float threshold, int squared,
                          float* out, float* workspace, void* stream) {
  PCD_CHECK_ARG(pred != nullptr && gt != nullptr && out != nullptr && workspace != nullptr, "fscore: null argument");
  PCD_CHECK_ARG(batch > 0 && batch <= 65535 && n > 0 && m > 0, "fscore: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  float* d1 = workspace;
  float* d2 = workspace + (size_t)batch * n;
  nn_rowmajor_kernel<<<dim3(ceil_div(n, 256), batch), 256, 0, st>>>(pred, n, gt, m, 0, d1, nullptr);
  nn_rowmajor_kernel<<<dim3(ceil_div(m, 256), batch), 256, 0, st>>>(gt, m, pred, n, 0, d2, nullptr);
  fscore_reduce_kernel<<<batch, 256, 0, st>>>(d1, n, d2, m, threshold, squared, out, batch);
  g_launch_count += 2;
  PCD_CHECK_LAUNCH("fscore");
  return PCD_OK;
}
