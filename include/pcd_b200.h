/*
 * pcd_b200.h -- C ABI of the B200-native denoising-sampler hot path.
 *
 * The reference (entheeb/A-Multimodal-Diffusion-Based-Model-for-Point-Cloud-Completion)
 * is 100 % Python and has no FFI: its boundary for this path is two Python call
 * signatures, PointCloudSampler(...) (diffusion/sampler.py:26-171) and
 * model(x, t, **kwargs) (models/transformer.py:195-226).  The Python shim in this
 * repo keeps those signatures and calls the functions below through ctypes; each
 * entry point cites the reference code it replaces.
 *
 * Conventions
 *  - every function returns 0 on success, a negative pcd_status otherwise;
 *    pcd_last_error() returns a thread-local message for the last failure.
 *  - all data pointers are CALLER-OWNED DEVICE memory (the shim allocates with
 *    PyTorch); nothing here allocates, synchronises or calls back into the host,
 *    so every entry point is CUDA-graph capturable.
 *  - `stream` is a cudaStream_t passed as void*; all work is enqueued on it.
 *  - matrices are row-major; "ld*" are leading dimensions in ELEMENTS.
 *  - bf16 values are raw uint16 storage (torch.bfloat16).
 */
#ifndef PCD_B200_H_
#define PCD_B200_H_

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define PCD_API __attribute__((visibility("default")))
#else
#define PCD_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  PCD_OK = 0,
  PCD_ERR_INVALID = -1,   /* bad argument / unsupported shape */
  PCD_ERR_CUDA = -2,      /* CUDA runtime or driver error (see pcd_last_error) */
  PCD_ERR_UNSUPPORTED = -3
} pcd_status;

typedef enum { PCD_F32 = 0, PCD_BF16 = 1 } pcd_precision;

/* GEMM epilogues: C = epi(A W^T + bias) */
typedef enum {
  PCD_EPI_BIAS = 0,          /* nn.Linear                         (transformer.py:45-47,55-56) */
  PCD_EPI_BIAS_GELU = 1,     /* Linear + exact-erf GELU           (transformer.py:57,62)       */
  PCD_EPI_BIAS_RESIDUAL = 2, /* residual + Linear                 (transformer.py:113-114)     */
  /* pcd_gemm_bf16_ex only -- the LayerNorms of a block folded into its projections: */
  PCD_EPI_RESIDUAL_STATS = 3,/* h <- h + A W^T + b (fp32, in place), C2 = bf16(h), per-row (mean, M2)
                                of C2 per 128 columns -> stats_out        (transformer.py:113-114) */
  PCD_EPI_LN_BIAS = 4,       /* LayerNorm(x) W^T + b evaluated as rstd (x (gamma o W)^T - mu s) + c
                                from the bf16 copy x and its row statistics (transformer.py:108-114) */
  PCD_EPI_LN_BIAS_GELU = 5   /* same + exact-erf GELU                  (transformer.py:110,57,62) */
} pcd_epilogue;

PCD_API int pcd_abi_version(void);
PCD_API const char* pcd_last_error(void);
/* Kernels launched through this library since it was loaded (host-side counter). */
PCD_API unsigned long long pcd_launch_count(void);
/* Device capability probe: 0 if the current device is sm_100 (B200). */
PCD_API int pcd_check_device(void);

/* There is no process-global state behind this ABI: kernel variants and profiling switches are per-call
 * arguments (pcd_attention `variant`, pcd_gemm_args.debug) or per-handle options (pcd_model_desc.flags). */

/* ------------------------------------------------------------------ */
/* Elementwise / normalisation kernels                                 */
/* ------------------------------------------------------------------ */

/* timestep_embedding (models/util.py:72-89): out[b, :] = [cos(t_b f) || sin(t_b f)],
 * f = host-computed frequency table [dim/2] (same fp32 values as the reference's CPU
 * torch.exp).  t is float (int64 timesteps are exact in fp32 below 2^24). */
PCD_API int pcd_timestep_embed(const float* t, const float* freqs, int batch, int dim,
                       float* out, int ld_out, void* stream);

/* nn.LayerNorm(dim, eps) with affine (transformer.py:108,110,176,186): x fp32
 * [rows, dim] -> out fp32 or bf16 [rows, dim]. */
PCD_API int pcd_layernorm(const float* x, int ldx, const float* gamma, const float* beta,
                  void* out, int ld_out, int out_precision, int rows, int dim, float eps,
                  void* stream);

/* Fused residual update + LayerNorm (transformer.py:113-114): h <- h + y, out = LayerNorm(h).
 * y is the bias-added output of the preceding c_proj GEMM ([rows, dim], bf16 or fp32); the
 * residual stream h stays fp32.  One HBM pass instead of a GEMM residual epilogue + a LayerNorm. */
PCD_API int pcd_add_layernorm(float* h, int ldh, const void* y, int ldy, int y_precision,
                              const float* gamma, const float* beta, void* out, int ld_out,
                              int out_precision, int rows, int dim, float eps, void* stream);

/* Token assembly + ln_pre (transformer.py:205-220).  For sequence s in [0, seqs)
 * and position l in [0, n_prefix + n_points):
 *   l <  n_prefix : row = prefix[s, l, :]                     (conditioning tokens)
 *   l >= n_prefix : row = W_in x[s % x_seqs, :, l-n_prefix] + b_in (+ add_cond[s, :])
 * then h[s, l, :] = LayerNorm(row).  x is [x_seqs, c_in, n_points] (NCL, as the
 * sampler holds it); w_in is [dim, c_in].  Optional (both or neither, dim % 128 == 0): h_bf16 = bf16 copy of h and
 * stats float2 [rows, dim/128] = (mean, M2) of the rounded values per 128 columns -- what pcd_cast_rowstats would
 * compute from h, emitted here so that the LayerNorm-folded forward needs no separate pass over the stream. */
PCD_API int pcd_embed_tokens(const float* x, int x_seqs, int c_in, int n_points,
                     const float* w_in, const float* b_in,
                     const float* prefix, int n_prefix, const float* add_cond,
                     const float* ln_g, const float* ln_b, float eps,
                     float* h, int seqs, int dim, uint16_t* h_bf16, float* stats, void* stream);

/* ln_post + token slice + output_proj + NLC->NCL permute (transformer.py:222-226):
 * out[s, c, n] = W_out[c, :] . LayerNorm(h[s, n_prefix + n, :] (+ y[...])) + b_out[c];
 * y (nullable, [seqs*L, dim] bf16/fp32) is the last block's pending MLP output. */
PCD_API int pcd_output_proj(const float* h, const void* y, int y_precision, int seqs, int n_prefix,
                            int n_points, int dim,
                    const float* ln_g, const float* ln_b, float eps,
                    const float* w_out, const float* b_out, int c_out,
                    float* out, void* stream);

/* ------------------------------------------------------------------ */
/* Projections: C[M,N] = epi(A[M,K] W[N,K]^T + bias)                    */
/* ------------------------------------------------------------------ */

/* fp32 mode: CUDA-core FFMA GEMM (fp32 accumulate), used for the 1e-4 parity mode
 * and for the tiny per-sequence layers (time MLP, clip_embed). residual/out fp32. */
PCD_API int pcd_gemm_f32(const float* A, int lda, const float* W, int ldw, const float* bias,
                 const float* residual, int ldr, float* C, int ldc,
                 int M, int N, int K, int epilogue, void* stream);

/* bf16 mode: tcgen05.mma (kind::f16, bf16 x bf16 -> fp32 in TMEM) fed by TMA with
 * 128B swizzle, persistent over 128 x BN tiles.  A, W bf16; bias/residual fp32;
 * C is bf16 (out_precision = PCD_BF16) or fp32.  Requires K % 8 == 0, lda % 8 == 0,
 * ldw % 8 == 0 (16-byte TMA strides). */
PCD_API int pcd_gemm_bf16(const uint16_t* A, int lda, const uint16_t* W, int ldw, const float* bias,
                  const float* residual, int ldr, void* C, int ldc, int out_precision,
                  int M, int N, int K, int epilogue, void* stream);

/* Extended form (argument block; zero-initialise, then fill in what the epilogue needs).
 *  PCD_EPI_RESIDUAL_STATS: residual (fp32, may alias C), C fp32, C2 = bf16 copy of the updated C,
 *    stats_out float2 [M, N/128] = (mean, M2 = sum (x - mean)^2) of C2 per 128 columns.
 *  PCD_EPI_LN_BIAS(_GELU): A = that bf16 copy [M, K], W = bf16(gamma o W_lin), colsum[n] = sum_k W[n, k]
 *    (of the rounded folded weights), bias[n] = beta . W_lin[n, :] + b_lin[n], stats_in float2 [M, K/128],
 *    ln_eps; C bf16.  Both need N % 256 == 0 and M >= 512 (CTA-pair kernel). */
typedef struct {
  const void* A; int lda;          /* bf16 [M, K] */
  const void* W; int ldw;          /* bf16 [N, K] */
  const float* bias;               /* fp32 [N] or NULL */
  const float* residual; int ldr;  /* fp32 [M, N] (residual epilogues) */
  void* C; int ldc; int out_precision;
  void* C2; int ldc2;              /* bf16 [M, N] second output (RESIDUAL_STATS) */
  float* stats_out;                /* float2 [M, N/128] (RESIDUAL_STATS) */
  const float* stats_in;           /* float2 [M, K/128] (LN_*) */
  const float* colsum;             /* fp32 [N] (LN_*) */
  float ln_eps;
  int M, N, K, epilogue;
  int debug;                       /* profiling aid (tools/gemm_probe.py): bit 0 skips the epilogue, bit 1 the TMA
                                      operand loads, bit 3 the MMAs -- results INVALID; bit 2 (results valid)
                                      forces the single-CTA kernel where the CTA-pair kernel would run */
} pcd_gemm_args;
PCD_API int pcd_gemm_bf16_ex(const pcd_gemm_args* args, void* stream);

/* out = a + b, fp32, n elements (out may alias a): stream additions of the TwoStream denoiser that are not the tail of
 * a projection (models/modules.py:228-229, models/model.py:536). */
PCD_API int pcd_add_f32(const float* a, const float* b, float* out, int64_t n, void* stream);

/* bf16 copy + row statistics of an fp32 matrix x [rows, dim] (dim % 128 == 0): out = bf16(x),
 * stats float2 [rows, dim/128] = (mean, M2) of the ROUNDED values per 128 columns -- the form the
 * LN-folded projections consume (used once per forward, after ln_pre). */
PCD_API int pcd_cast_rowstats(const float* x, int ldx, uint16_t* out, int ld_out, float* stats,
                              int rows, int dim, void* stream);

/* ------------------------------------------------------------------ */
/* Attention (head dim 64), softmax in fp32                             */
/* ------------------------------------------------------------------ */

/* Strided description of one operand: element (b, l, h, c) lives at
 * ptr[b*batch_stride + l*row_stride + h*head_stride + c], c in [0, 64). */
typedef struct {
  const void* ptr;
  int64_t batch_stride;
  int64_t row_stride;
  int64_t head_stride;
} pcd_attn_operand;

/* Tensor-core attention kernels selectable per call (bf16 only; the fp32 kernel ignores `variant`):
 *  GROUPED        one CTA per SM over three query tiles of one (sequence, head) that share every K / V tile: three
 *                 softmax warps per scheduler keep the special-function pipe fed (attn_tc8.cu) -- the default;
 *                 0..16 keys beyond the last full 64-key tile (every registered sequence length) ride on the last full
 *                 step instead of taking a step of their own;
 *  GROUPED_STEPTAIL the same kernel with such a tail as an ordinary (masked) KV step: A/B of the tail path;
 *  GROUPED_WIDE   the same kernel in its second geometry: two query tiles per CTA, 128-key steps;
 *  GROUPED_TOKEN  the same kernel with a per-scheduler MUFU token (one of the three warps exponentiates at a time);
 *                 measured slower than GROUPED, kept for A/B runs;
 *  PAIRED         round-1 kernel: two CTAs per SM, one query tile each, softmax software-pipelined over KV tiles
 *                 (attn_tc5.cu); _POLY4 / _POLY2 evaluate 1/4 / 1/2 of the exponentials with an FMA-pipe polynomial. */
typedef enum {
  PCD_ATTN_DEFAULT = 0,
  PCD_ATTN_GROUPED = 8,
  PCD_ATTN_GROUPED_TOKEN = 9,
  PCD_ATTN_GROUPED_STEPTAIL = 10,
  PCD_ATTN_GROUPED_WIDE = 11,
  PCD_ATTN_PAIRED = 5,
  PCD_ATTN_PAIRED_POLY4 = 6,
  PCD_ATTN_PAIRED_POLY2 = 7
} pcd_attn_variant;

/* out[b, l, h*64 + c] = sum_s softmax_s((q_scale q[b,l,h]) . (k_scale k[b,s,h])) v[b,s,h,c]
 *  - self-attention   (transformer.py:65-84):   q,k,v views of qkv [B,L,H,3,64], scales 64^-1/4
 *  - cross-attention  (perceiver.py:46-67):     q [B,Lq,H,64], k,v views of kv [B,Lkv,H,2,64]
 *  - rotary attention (rotaryencoderpcd.py:6-27,68-84): rope_coords [B,L,3] != NULL applies
 *    theta = pi*coords to head dims 0..5 of q and k (requires Lq == Lkv);
 *    q_scale = width^-1/2, k_scale = 1.  Both kernels rotate inside the attention kernel: the fp32 one in
 *    registers while it loads the tiles, the tcgen05 one in shared memory between the TMA arrival of a Q / K
 *    tile and the first MMA that reads it -- no rotated copy of q / k is ever written to HBM.
 * `precision` selects the operand/out dtype: PCD_F32 (CUDA-core kernel) or PCD_BF16
 * (tcgen05 flash kernel: QK^T and PV on tensor cores, accumulators in TMEM). */
PCD_API int pcd_attention(const pcd_attn_operand* q, const pcd_attn_operand* k, const pcd_attn_operand* v,
                  void* out, int64_t out_batch_stride, int64_t out_row_stride,
                  int batch, int heads, int len_q, int len_kv,
                  float q_scale, float k_scale, const float* rope_coords,
                  int precision, int variant, void* stream);

/* fp32 attention with head dim 32 (c in [0, 32) in the operand description): the cross-attention of the TwoStream
 * denoiser's read / compute / write blocks (models/modules.py:17-63: z_dim = x_dim = 256, 8 heads); CUDA-core kernel,
 * softmax((q_scale q) . (k_scale k)) v like pcd_attention. */
PCD_API int pcd_attention_hd32(const pcd_attn_operand* q, const pcd_attn_operand* k, const pcd_attn_operand* v,
                               float* out, int64_t out_batch_stride, int64_t out_row_stride, int batch, int heads,
                               int len_q, int len_kv, float q_scale, float k_scale, void* stream);

/* The same product on tensor cores (bf16 operands [.., H, 32], bf16 out [.., H, 32]): the grouped 64-wide kernel with
 * tensor maps whose 64-column boxes hang over the 32-column heads -- TMA zero-fills the missing columns on load and
 * drops them on store, so no zero-padded copy of q / k / v exists in memory.  `variant`: a grouped pcd_attn_variant or
 * PCD_ATTN_DEFAULT. */
PCD_API int pcd_attention_hd32_bf16(const pcd_attn_operand* q, const pcd_attn_operand* k, const pcd_attn_operand* v,
                            void* out, int64_t out_batch_stride, int64_t out_row_stride,
                            int batch, int heads, int len_q, int len_kv,
                            float q_scale, float k_scale, int variant, void* stream);

/* Stand-alone form of the rotation (apply_rotary_pos_emb, rotaryencoderpcd.py:6-27; fp32 arithmetic): rotate head
 * dims 0..5 of a bf16 q or k operand IN PLACE with theta = pi * coords[b, l, :].  pcd_attention does this inside the
 * kernel; this entry exists for callers that want rotated projections for something else (and as the test yardstick
 * of the fused path). */
PCD_API int pcd_rope_bf16(const pcd_attn_operand* x, const float* coords, int batch, int heads, int len,
                          void* stream);

/* ------------------------------------------------------------------ */
/* Fused Karras/Heun sampler updates (k_diffusion.py:270-310,79-108,182-207;  */
/* gaussian_diffusion.py:320-357,949-958).  State fp32 [B, C, N].              */
/* ------------------------------------------------------------------ */

typedef struct {
  float c_in;         /* 1/sqrt(sigma_eval^2+1) of the evaluation that produced `model_out` */
  float coef_x;       /* sqrt(1/abar_t)     at the truncated integer t of that evaluation   */
  float coef_eps;     /* sqrt(1/abar_t - 1)                                                 */
  float sigma;        /* sigma of that evaluation (sigma_hat_i or sigma_{i+1})              */
  float dt;           /* sigma_{i+1} - sigma_hat_i                                          */
  float guidance;     /* classifier-free guidance scale (used iff uncond rows present)      */
  float clip;         /* 1: clamp x0 to [-1,1] (clip_denoised), 0: no clamp                  */
  float next_c_in;    /* c_in of the NEXT evaluation (prescale of the next model input)     */
  float next_noise;   /* sqrt(sigma_hat^2 - sigma^2) of the next step's churn (0: none)     */
  float dt2;          /* DPM-2 only: sigma_{i+1} - sigma_hat_i (dt is then sigma_mid - sigma_hat_i) */
  float mode;         /* corrector: 0 = Heun average (k_diffusion.py:305-309), 1 = DPM-2 midpoint
                         update x <- x + d_2*dt2 (k_diffusion.py:343-350)                    */
} pcd_step_scalars;

/* Start of the loop: x <- x + noise*s.next_noise (if != 0); model_in <- x*s.next_c_in. */
PCD_API int pcd_sampler_begin(float* x, const float* noise, float* model_in,
                      const pcd_step_scalars* s, int64_t numel, void* stream);

/* First (Euler/predictor) evaluation of Heun step i.  model_out is [seqs_out, c_out, N]
 * with the conditional rows first and, if `guided`, the unconditional rows after them;
 * only channels [0, C) (epsilon) are read.
 *   x0 = clamp(coef_x*(x*c_in) - coef_eps*eps)      per branch, then CFG combine
 *   d  = (x - x0)/sigma;  pred_unscaled = (x0 - ch_bias)/ch_scale
 *   last == 0: model_in <- (x + d*dt)*next_c_in, d stored for the corrector
 *   last != 0: x <- x + d*dt                                        (sigma_{i+1} == 0) */
PCD_API int pcd_sampler_predictor(float* x, const float* model_out, int c_out, int guided,
                          float* d, float* model_in, float* pred_unscaled,
                          const float* ch_scale, const float* ch_bias,
                          const pcd_step_scalars* s, int batch, int channels, int n_points,
                          int last, void* stream);

/* Second (corrector) evaluation: x2 = x + d*dt; d2 = (x2 - x0_2)/sigma;
 * x <- x + (d+d2)/2*dt; then the next step's churn + prescale:
 * x <- x + noise*next_noise; model_in <- x*next_c_in. */
PCD_API int pcd_sampler_corrector(float* x, const float* model_out, int c_out, int guided,
                          const float* d, const float* noise, float* model_in,
                          const pcd_step_scalars* s, int batch, int channels, int n_points,
                          void* stream);

/* Ancestral (DDPM) sampling step: GaussianDiffusion.p_mean_variance + p_sample
 * (reference diffusion/gaussian_diffusion.py:257-350, 407-449; the loop of p_sample_loop_progressive, :499-548, calls
 * it once per step) fused into one pass over the state.  `table` is [num_timesteps, PCD_DDPM_COLS] float32 (float64
 * schedule arrays rounded like _extract_into_tensor(...).float()), `t` the per-sample int64 step indices [batch].
 * model_out is [batch, out_channels, n_points]: channels [0, C) epsilon, [C, 2C) the variance channel of
 * learned / learned_range models.  x is read, x_next written (may alias x); every other output pointer may be NULL.
 * noise == NULL gives the mean (p_mean_variance only). */
enum { PCD_DDPM_RECIP = 0,    /* sqrt(1 / alphas_cumprod)          gaussian_diffusion.py:180 */
       PCD_DDPM_RECIPM1 = 1,  /* sqrt(1 / alphas_cumprod - 1)      :181 */
       PCD_DDPM_MEAN_X0 = 2,  /* posterior_mean_coef1              :191-193 */
       PCD_DDPM_MEAN_XT = 3,  /* posterior_mean_coef2              :194-196 */
       PCD_DDPM_MIN_LOG = 4,  /* posterior_log_variance_clipped    :188-190 */
       PCD_DDPM_MAX_LOG = 5,  /* log(betas)                        :299 */
       PCD_DDPM_FIXED_LOG = 6,/* log-variance of fixed_small / fixed_large models  :305-318 */
       PCD_DDPM_COLS = 8 };
enum { PCD_VAR_FIXED = 0, PCD_VAR_LEARNED_RANGE = 1, PCD_VAR_LEARNED = 2 };
typedef struct pcd_ddpm_args {
  const float* x;           /* [batch, channels, n_points] state x_t */
  const float* model_out;   /* [batch, out_channels, n_points] */
  const float* noise;       /* [batch, channels, n_points] or NULL */
  const int64_t* t;         /* [batch] */
  const float* table;       /* [num_timesteps, PCD_DDPM_COLS] */
  const float* ch_scale;    /* [channels] or NULL (GaussianDiffusion.channel_scales) */
  const float* ch_bias;     /* [channels] or NULL */
  float* x_next;            /* sample x_{t-1} (scaled units) */
  float* pred_xstart;       /* x0 prediction; unscaled when `unscale` != 0 */
  float* sample_unscaled;   /* unscale_channels(x_next) */
  float* mean;              /* posterior mean */
  float* log_variance;
  int batch, channels, n_points, out_channels;
  int var_mode;             /* PCD_VAR_* */
  int clip_denoised;
  int unscale;
  int num_timesteps;        /* rows of `table`: step indices are clamped to [0, num_timesteps) on the device */
} pcd_ddpm_args;
PCD_API int pcd_ddpm_step(const pcd_ddpm_args* args, void* stream);


/* Squared-L2 Chamfer distance on xyz (models/util.py:265-295): p1 [B,C1,N1],
 * p2 [B,C2,N2] -> out [B].  Tiled nearest-neighbour search, no [B,N1,N2] matrix. */
PCD_API int pcd_chamfer(const float* p1, int c1, int n1, const float* p2, int c2, int n2,
                int batch, float* out, float* workspace /* >= batch*(n1+n2) floats */,
                void* stream);

/* ------------------------------------------------------------------ */
/* Point-cloud utilities either side of the sampler (evaluation / IO layer) */
/* ------------------------------------------------------------------ */

/* Greedy farthest-point sampling (util/point_cloud.py:82-118; evaluation.py:40-48): points fp32
 * [batch, n, 3] row-major, init_idx int32 [batch] -> out_idx int64 [batch, n_samples]; distances in the
 * reference's |a|^2 + |b|^2 - 2 a.b form, first arg-max.  Up to 8192 points per cloud the running distances live in
 * registers (workspace may be NULL); larger clouds need workspace = [batch, n] floats.  init_idx is clamped to the cloud. */
PCD_API int pcd_farthest_point_sample(const float* points, int batch, int n, int n_samples, const int* init_idx,
                                      long long* out_idx, float* workspace, void* stream);
/* For every point of a [batch, na, 3]: squared distance to / index of its nearest point of b [batch, nb, 3]
 * (either output may be NULL).  form 0: sum (a-b)^2 (models/util.py:213-214); form 1: |a|^2 + |b|^2 - 2 a.b
 * (PointCloud.nearest_points, util/point_cloud.py:148-165).  First minimum wins. */
PCD_API int pcd_nearest_points(const float* a, int na, const float* b, int nb, int batch, int form, float* out_d2,
                               long long* out_idx, void* stream);
/* fscore_point_cloud_batch (squared == 0: threshold on the distance) / _squared (threshold on the squared
 * distance), models/util.py:195-262: pred [batch, n, 3], gt [batch, m, 3] -> out fp32 [3, batch] =
 * (fscore, precision, recall); workspace >= batch * (n + m) floats.  No [n, m, 3] tensor. */
PCD_API int pcd_fscore(const float* pred, int n, const float* gt, int m, int batch, float threshold, int squared,
                       float* out, float* workspace, void* stream);

/* ------------------------------------------------------------------ */
/* Whole-denoiser forward (transformer.py:118-152,195-226)             */
/* ------------------------------------------------------------------ */

typedef struct {
  const float *ln1_g, *ln1_b, *ln2_g, *ln2_b;          /* ln_1, ln_2                     */
  const void *w_qkv, *w_proj, *w_fc, *w_fc2;           /* [3d,d] [d,d] [4d,d] [d,4d]; fp32 or bf16 per precision */
  const float *b_qkv, *b_proj, *b_fc, *b_fc2;
  /* Optional (bf16 mode; all six or none): LayerNorm-folded forms of c_qkv and mlp.c_fc, prepared by
   * the host once per weight load -- w_*_ln = bf16(gamma o W), *_colsum[n] = sum_k w_*_ln[n, k],
   * *_const[n] = beta . W[n, :] + b[n].  When present (and width % 256 == 0, >= 512 tokens) the
   * forward runs without LayerNorm kernels. */
  const void *w_qkv_ln, *w_fc_ln;
  const float *qkv_colsum, *qkv_const, *fc_colsum, *fc_const;
} pcd_block_weights;

typedef struct {
  int precision;            /* pcd_precision of the backbone                        */
  int width, heads, layers; /* width = heads*64                                     */
  int c_in, c_out;          /* input_proj / output_proj channels                    */
  int n_points;             /* n_ctx of the point tokens                            */
  int n_prefix;             /* conditioning tokens in front of the points           */
  int time_slot;            /* index of the time token in the prefix, or -1 when the
                               time embedding is ADDED to the point tokens instead  */
  float ln_eps;
  int flags;                /* PCD_MODEL_* options of this handle                     */
  /* always fp32: */
  const float *time_fc_w, *time_fc_b, *time_proj_w, *time_proj_b;   /* time_embed MLP      */
  const float *freqs;                                                /* [width/2]          */
  const float *ln_pre_g, *ln_pre_b, *ln_post_g, *ln_post_b;
  const float *in_w, *in_b, *out_w, *out_b;                          /* input/output_proj  */
  const pcd_block_weights* blocks;                                   /* HOST array [layers] */
} pcd_model_desc;

enum { PCD_MODEL_SEPARATE_LAYERNORM = 1,  /* run LayerNorm kernels even when the folded weights are present */
       PCD_MODEL_ATTN_VARIANT_SHIFT = 8 };   /* bits 8..15: pcd_attn_variant of the forward (0 = default)      */

typedef struct pcd_model pcd_model;

PCD_API int pcd_model_create(const pcd_model_desc* desc, pcd_model** out);
PCD_API int pcd_model_destroy(pcd_model* m);
/* Bytes of device workspace pcd_model_forward needs for `seqs` sequences. */
PCD_API size_t pcd_model_workspace_bytes(const pcd_model* m, int seqs);

/* out[s] = model(x[s % x_seqs], t[s], prefix[s]) for s in [0, seqs).
 *  x       fp32 [x_seqs, c_in, n_points]  (x_seqs == seqs, or seqs/2 when the cond and
 *          uncond halves of classifier-free guidance share the same x)
 *  t       fp32 [seqs], or NULL when the caller has already written the time token into the
 *          prefix slot / folded it into add_cond (one time-MLP evaluation per distinct t)
 *  prefix  fp32 [seqs, n_prefix, width]; slot `time_slot` is (over)written with the
 *          time-MLP output; the other slots hold step-invariant conditioning tokens
 *  add_cond fp32 [seqs, width] or NULL: non-token conditioning added to point tokens
 *  out     fp32 [seqs, c_out_eff, n_points] where c_out_eff = out_channels (<= c_out:
 *          the sampler only needs the epsilon half, transformer.py:225) */
PCD_API int pcd_model_forward(pcd_model* m, const float* x, int x_seqs, const float* t,
                      float* prefix, const float* add_cond, float* out, int out_channels,
                      void* workspace, size_t workspace_bytes, int seqs, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PCD_B200_H_ */
