// C-ABI glue: error reporting, device probe, tensor-map encoding, attention
// dispatch and the whole-denoiser forward (see include/pcd_b200.h).
#include <stdarg.h>

#include <vector>

#include "common.cuh"
#include "tc_sm100.cuh"

namespace pcd {

static thread_local char g_err[512] = "";
unsigned long long g_launch_count = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int num_sms() {  // of the CURRENT device (a process may drive more than one)
  static int cache[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (cache[dev] == 0) {
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    cache[dev] = n > 0 ? n : 148;
  }
  return cache[dev];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || p == nullptr) {
      set_error("cuTensorMapEncodeTiled unavailable: %s", cudaGetErrorString(e));
      return nullptr;
    }
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

static int encode_tmap(CUtensorMap* map, CUtensorMapDataType dt, const void* base, int rank,
                       const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box);

int encode_tmap_bf16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                     const uint64_t* strides_bytes, const uint32_t* box) {
  return encode_tmap(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, base, rank, dims, strides_bytes, box);
}
int encode_tmap_f32(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box) {
  return encode_tmap(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, base, rank, dims, strides_bytes, box);
}

static int encode_tmap(CUtensorMap* map, CUtensorMapDataType dt, const void* base, int rank,
                       const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) return PCD_ERR_CUDA;
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t bdim[5], estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
  }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  CUresult r = fn(map, dt, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr,
                  bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (CUresult %d; rank %d dims %llu,%llu stride0 %llu)", (int)r, rank,
              (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)strides_bytes[0]);
    return PCD_ERR_CUDA;
  }
  return PCD_OK;
}

int launch_attention_f32(const pcd_attn_operand* q, const pcd_attn_operand* k, const pcd_attn_operand* v,
                         float* out, int64_t o_bs, int64_t o_ls, int batch, int heads, int len_q,
                         int len_kv, float q_scale, float k_scale, const float* rope, int head_dim, cudaStream_t st);
int launch_rope_bf16(uint16_t* x, int64_t bs, int64_t ls, int64_t hs, const float* coords, int batch, int heads,
                     int len, cudaStream_t st);
int launch_attention_bf16(const pcd_attn_operand* q, const pcd_attn_operand* k, const pcd_attn_operand* v,
                          uint16_t* out, int64_t o_bs, int64_t o_ls, int batch, int heads, int len_q,
                          int len_kv, float q_scale, float k_scale, const float* rope, int variant, int head_dim,
                          cudaStream_t st);

}  // namespace pcd

using namespace pcd;

extern "C" int pcd_abi_version(void) { return 4; }
extern "C" unsigned long long pcd_launch_count(void) { return g_launch_count; }
extern "C" const char* pcd_last_error(void) { return g_err; }

extern "C" int pcd_check_device(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    set_error("no CUDA device: %s", cudaGetErrorString(e));
    return PCD_ERR_CUDA;
  }
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (major != 10) {
    set_error("pcd_b200 kernels are built for sm_100a only; device is sm_%d%d", major, minor);
    return PCD_ERR_UNSUPPORTED;
  }
  return PCD_OK;
}

static bool operand_ok(const pcd_attn_operand* o, int align_elems, size_t elem) {
  return o && o->ptr && (reinterpret_cast<uintptr_t>(o->ptr) % (align_elems * elem) == 0) &&
         o->row_stride % align_elems == 0 && o->head_stride % align_elems == 0 &&
         o->batch_stride % align_elems == 0;
}

extern "C" int pcd_attention(const pcd_attn_operand* q, const pcd_attn_operand* k, const pcd_attn_operand* v,
                             void* out, int64_t o_bs, int64_t o_ls, int batch, int heads, int len_q,
                             int len_kv, float q_scale, float k_scale, const float* rope_coords,
                             int precision, int variant, void* stream) {
  PCD_CHECK_ARG(batch > 0 && heads > 0 && len_q > 0 && len_kv > 0, "attention: empty problem");
  PCD_CHECK_ARG(batch <= 65535 && heads <= 65535, "attention: batch/heads exceed grid limits");
  PCD_CHECK_ARG(q_scale > 0.f && k_scale > 0.f, "attention: scales must be positive");
  PCD_CHECK_ARG(rope_coords == nullptr || len_q == len_kv, "attention: rotary needs len_q == len_kv");
  cudaStream_t st = (cudaStream_t)stream;
  if (precision == PCD_F32) {
    PCD_CHECK_ARG(operand_ok(q, 4, 4) && operand_ok(k, 4, 4) && operand_ok(v, 4, 4), "attention(f32): operands must be 16-byte aligned with strides %% 4 == 0");
    PCD_CHECK_ARG(o_ls % 4 == 0 && o_bs % 4 == 0, "attention(f32): output strides must be multiples of 4");
    return launch_attention_f32(q, k, v, (float*)out, o_bs, o_ls, batch, heads, len_q, len_kv, q_scale, k_scale, rope_coords, 64, st);
  }
  if (precision == PCD_BF16) {
    PCD_CHECK_ARG(operand_ok(q, 8, 2) && operand_ok(k, 8, 2) && operand_ok(v, 8, 2), "attention(bf16): operands must be 16-byte aligned with strides %% 8 == 0");
    PCD_CHECK_ARG(o_ls % 8 == 0 && o_bs % 8 == 0 && reinterpret_cast<uintptr_t>(out) % 16 == 0, "attention(bf16): output must be 16-byte aligned with strides %% 8 == 0");
    return launch_attention_bf16(q, k, v, (uint16_t*)out, o_bs, o_ls, batch, heads, len_q, len_kv, q_scale, k_scale,
                                 rope_coords, variant, 64, st);
  }
  PCD_CHECK_ARG(false, "attention: unknown precision %d", precision);
}

extern "C" int pcd_attention_hd32(const pcd_attn_operand* q, const pcd_attn_operand* k, const pcd_attn_operand* v,
                                  float* out, int64_t out_batch_stride, int64_t out_row_stride, int batch, int heads,
                                  int len_q, int len_kv, float q_scale, float k_scale, void* stream) {
  PCD_CHECK_ARG(q != nullptr && k != nullptr && v != nullptr && out != nullptr, "attention_hd32: null argument");
  PCD_CHECK_ARG(batch > 0 && heads > 0 && len_q > 0 && len_kv > 0, "attention_hd32: empty problem");
  PCD_CHECK_ARG(batch <= 65535 && heads <= 65535, "attention_hd32: batch/heads exceed grid limits");
  PCD_CHECK_ARG(q_scale > 0.f && k_scale > 0.f, "attention_hd32: scales must be positive");
  PCD_CHECK_ARG(operand_ok(q, 4, 4) && operand_ok(k, 4, 4) && operand_ok(v, 4, 4), "attention_hd32: operands must be 16-byte aligned with strides %% 4 == 0");
  return launch_attention_f32(q, k, v, out, out_batch_stride, out_row_stride, batch, heads, len_q, len_kv, q_scale, k_scale,
                              nullptr, 32, (cudaStream_t)stream);
}

extern "C" int pcd_attention_hd32_bf16(const pcd_attn_operand* q, const pcd_attn_operand* k, const pcd_attn_operand* v,
                                       void* out, int64_t o_bs, int64_t o_ls, int batch, int heads, int len_q,
                                       int len_kv, float q_scale, float k_scale, int variant, void* stream) {
  PCD_CHECK_ARG(out != nullptr, "attention_hd32_bf16: null argument");
  PCD_CHECK_ARG(batch > 0 && heads > 0 && len_q > 0 && len_kv > 0, "attention_hd32_bf16: empty problem");
  PCD_CHECK_ARG(batch <= 65535 && heads <= 65535, "attention_hd32_bf16: batch/heads exceed grid limits");
  PCD_CHECK_ARG(q_scale > 0.f && k_scale > 0.f, "attention_hd32_bf16: scales must be positive");
  PCD_CHECK_ARG(operand_ok(q, 8, 2) && operand_ok(k, 8, 2) && operand_ok(v, 8, 2), "attention_hd32_bf16: operands must be 16-byte aligned with strides %% 8 == 0");
  PCD_CHECK_ARG(o_ls % 8 == 0 && o_bs % 8 == 0 && reinterpret_cast<uintptr_t>(out) % 16 == 0, "attention_hd32_bf16: output must be 16-byte aligned with strides %% 8 == 0");
  return launch_attention_bf16(q, k, v, (uint16_t*)out, o_bs, o_ls, batch, heads, len_q, len_kv, q_scale, k_scale, nullptr,
                               variant, 32, (cudaStream_t)stream);
}

extern "C" int pcd_rope_bf16(const pcd_attn_operand* x, const float* coords, int batch, int heads, int len,
                             void* stream) {
  PCD_CHECK_ARG(x != nullptr && x->ptr != nullptr && coords != nullptr, "rope_bf16: null argument");
  PCD_CHECK_ARG(batch > 0 && heads > 0 && len > 0, "rope_bf16: empty problem");
  PCD_CHECK_ARG(reinterpret_cast<uintptr_t>(x->ptr) % 4 == 0 && x->row_stride % 2 == 0 && x->head_stride % 2 == 0 &&
                    x->batch_stride % 2 == 0,
                "rope_bf16: operand must be 4-byte aligned with even strides");
  return launch_rope_bf16((uint16_t*)const_cast<void*>(x->ptr), x->batch_stride, x->row_stride, x->head_stride, coords,
                          batch, heads, len, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------
// Whole-denoiser forward
// ---------------------------------------------------------------------------
struct pcd_model {
  pcd_model_desc d;
  std::vector<pcd_block_weights> blocks;
};

namespace {
struct Workspace {
  float *temb, *thid, *tcond, *h, *stats;
  void *xn, *qkv, *att, *hid, *y;
  size_t total;
};

inline size_t align_up(size_t x) { return (x + 255) & ~size_t(255); }

Workspace carve(const pcd_model_desc& d, int seqs, unsigned char* base) {
  Workspace w;
  const size_t M = (size_t)seqs * (d.n_prefix + d.n_points);
  const size_t es = d.precision == PCD_BF16 ? 2 : 4;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    unsigned char* p = base ? base + off : nullptr;
    off += align_up(bytes);
    return p;
  };
  w.temb = (float*)take((size_t)seqs * d.width * 4);
  w.thid = (float*)take((size_t)seqs * d.width * 4 * 4);
  w.tcond = (float*)take((size_t)seqs * d.width * 4);
  w.h = (float*)take(M * d.width * 4);
  w.xn = take(M * d.width * es);
  w.qkv = take(M * d.width * 3 * es);
  w.att = take(M * d.width * es);
  w.hid = take(M * d.width * 4 * es);
  w.y = take(M * d.width * es);
  w.stats = (float*)take(M * (size_t)((d.width + 127) / 128) * 8);  // (mean, M2) per 128 columns (LN-folded path)
  w.total = off;
  return w;
}
}  // namespace

extern "C" int pcd_model_create(const pcd_model_desc* desc, pcd_model** out) {
  PCD_CHECK_ARG(desc != nullptr && out != nullptr, "model_create: null argument");
  PCD_CHECK_ARG(desc->precision == PCD_F32 || desc->precision == PCD_BF16, "model_create: bad precision");
  PCD_CHECK_ARG(desc->heads > 0 && desc->width == desc->heads * 64, "model_create: width must equal heads*64 (width=%d heads=%d)", desc->width, desc->heads);
  PCD_CHECK_ARG(desc->layers > 0 && desc->blocks != nullptr, "model_create: no layers");
  PCD_CHECK_ARG(desc->c_in >= 1 && desc->c_in <= 8 && desc->c_out >= 1 && desc->c_out <= 32, "model_create: unsupported channel counts");
  PCD_CHECK_ARG(desc->n_points > 0 && desc->n_prefix >= 0, "model_create: bad token counts");
  PCD_CHECK_ARG(desc->time_slot < desc->n_prefix, "model_create: time_slot outside the prefix");
  pcd_model* m = new pcd_model();
  m->d = *desc;
  m->blocks.assign(desc->blocks, desc->blocks + desc->layers);
  m->d.blocks = m->blocks.data();
  *out = m;
  return PCD_OK;
}

extern "C" int pcd_model_destroy(pcd_model* m) {
  delete m;
  return PCD_OK;
}

extern "C" size_t pcd_model_workspace_bytes(const pcd_model* m, int seqs) {
  if (m == nullptr || seqs <= 0) return 0;
  return carve(m->d, seqs, nullptr).total;
}

#define PCD_TRY(expr)              \
  do {                             \
    int rc__ = (expr);             \
    if (rc__ != PCD_OK) return rc__; \
  } while (0)

extern "C" int pcd_model_forward(pcd_model* m, const float* x, int x_seqs, const float* t, float* prefix,
                                 const float* add_cond, float* out, int out_channels, void* workspace,
                                 size_t workspace_bytes, int seqs, void* stream) {
  PCD_CHECK_ARG(m != nullptr && x != nullptr && out != nullptr, "model_forward: null argument");
  const pcd_model_desc& d = m->d;
  PCD_CHECK_ARG(seqs > 0 && x_seqs > 0 && seqs % x_seqs == 0, "model_forward: seqs must be a multiple of x_seqs");
  PCD_CHECK_ARG(out_channels >= 1 && out_channels <= d.c_out, "model_forward: out_channels out of range");
  PCD_CHECK_ARG(d.n_prefix == 0 || prefix != nullptr, "model_forward: prefix buffer missing");
  PCD_CHECK_ARG(reinterpret_cast<uintptr_t>(workspace) % 256 == 0, "model_forward: workspace must be 256-byte aligned");
  Workspace w = carve(d, seqs, (unsigned char*)workspace);
  PCD_CHECK_ARG(workspace != nullptr && workspace_bytes >= w.total, "model_forward: workspace too small (%zu < %zu)", workspace_bytes, w.total);
  const int W = d.width, L = d.n_prefix + d.n_points;
  const int64_t M64 = (int64_t)seqs * L;
  PCD_CHECK_ARG(M64 < (int64_t)1 << 31, "model_forward: too many tokens");
  const int M = (int)M64;
  const bool bf = d.precision == PCD_BF16;

  // time embedding MLP (transformer.py:202; models/util.py:72-89) -- always fp32.  With t == NULL
  // the caller has already placed the time token in the prefix slot (or folded it into add_cond):
  // the sampler evaluates every sequence at the same timestep, so the shim computes each distinct
  // time token once per stage instead of once per sequence per evaluation.
  const float* addc = add_cond;
  if (t != nullptr) {
    PCD_TRY(pcd_timestep_embed(t, d.freqs, seqs, W, w.temb, W, stream));
    PCD_TRY(pcd_gemm_f32(w.temb, W, d.time_fc_w, W, d.time_fc_b, nullptr, 0, w.thid, 4 * W, seqs, 4 * W, W, PCD_EPI_BIAS_GELU, stream));
    if (d.time_slot >= 0) {
      PCD_TRY(pcd_gemm_f32(w.thid, 4 * W, d.time_proj_w, 4 * W, d.time_proj_b, nullptr, 0,
                           prefix + (size_t)d.time_slot * W, d.n_prefix * W, seqs, W, 4 * W, PCD_EPI_BIAS, stream));
    } else {
      // time embedding is added to every point token (transformer.py:209-211)
      PCD_TRY(pcd_gemm_f32(w.thid, 4 * W, d.time_proj_w, 4 * W, d.time_proj_b, add_cond, W, w.tcond, W, seqs, W, 4 * W,
                           add_cond ? PCD_EPI_BIAS_RESIDUAL : PCD_EPI_BIAS, stream));
      addc = w.tcond;
    }
  }
  // LayerNorm-folded path (bf16, width % 256 == 0, >= 512 tokens, folded weights supplied): no
  // LayerNorm kernels at all.  The c_proj / mlp.c_proj GEMMs update the fp32 residual stream in their
  // epilogue and emit its bf16 copy + per-row statistics; c_qkv / c_fc consume that copy with
  // W' = gamma o W and apply mean / rstd algebraically in their epilogue (gemm_tc.cu).
  bool fold = bf && W % 256 == 0 && M >= 512 && !(d.flags & PCD_MODEL_SEPARATE_LAYERNORM);
  for (int l = 0; l < d.layers && fold; ++l) {
    const pcd_block_weights& b = m->blocks[l];
    fold = b.w_qkv_ln && b.w_fc_ln && b.qkv_colsum && b.qkv_const && b.fc_colsum && b.fc_const;
  }
  // input_proj + token concat + ln_pre (transformer.py:208-220); on the folded path the same kernel also emits the
  // bf16 copy of the stream and its row statistics for the first c_qkv
  PCD_TRY(pcd_embed_tokens(x, x_seqs, d.c_in, d.n_points, d.in_w, d.in_b, prefix, d.n_prefix, addc, d.ln_pre_g,
                           d.ln_pre_b, d.ln_eps, w.h, seqs, W, fold ? (uint16_t*)w.xn : nullptr, fold ? w.stats : nullptr,
                           stream));

  const float qk_scale = 1.0f / sqrtf(sqrtf(64.0f));  // hd^-1/4 on q and on k (transformer.py:76)
  const int prec = d.precision;
  const size_t es = bf ? 2 : 4;
  // The two residual adds of every block (x = x + attn(..), x = x + mlp(..), transformer.py:113-114)
  // are folded into the LayerNorm that follows them: the c_proj GEMMs write y = A W^T + b and the
  // next LayerNorm kernel does h += y before normalising (one HBM pass, plain GEMM epilogues).
  auto gemm = [&](const void* A, int lda, const void* Wt, const float* bias, void* C, int N, int K, int epi) -> int {
    if (bf)
      return pcd_gemm_bf16((const uint16_t*)A, lda, (const uint16_t*)Wt, K, bias, nullptr, 0, C, N, PCD_BF16, M, N, K, epi, stream);
    return pcd_gemm_f32((const float*)A, lda, (const float*)Wt, K, bias, nullptr, 0, (float*)C, N, M, N, K, epi, stream);
  };
  const int attn_variant = (d.flags >> PCD_MODEL_ATTN_VARIANT_SHIFT) & 0xff;
  if (fold) {
    auto lin_ln = [&](const void* Wt, const float* colsum, const float* cst, void* C, int N, int epi) -> int {
      pcd_gemm_args g = {};
      g.A = w.xn; g.lda = W; g.W = Wt; g.ldw = W; g.bias = cst; g.colsum = colsum; g.stats_in = w.stats;
      g.ln_eps = d.ln_eps; g.C = C; g.ldc = N; g.out_precision = PCD_BF16; g.M = M; g.N = N; g.K = W; g.epilogue = epi;
      return pcd_gemm_bf16_ex(&g, stream);
    };
    auto lin_res = [&](const void* A, int K, const void* Wt, const float* bias) -> int {
      pcd_gemm_args g = {};
      g.A = A; g.lda = K; g.W = Wt; g.ldw = K; g.bias = bias; g.residual = w.h; g.ldr = W; g.C = w.h; g.ldc = W;
      g.out_precision = PCD_F32; g.C2 = w.xn; g.ldc2 = W; g.stats_out = w.stats; g.M = M; g.N = W; g.K = K;
      g.epilogue = PCD_EPI_RESIDUAL_STATS;
      return pcd_gemm_bf16_ex(&g, stream);
    };
    for (int l = 0; l < d.layers; ++l) {
      const pcd_block_weights& b = m->blocks[l];
      PCD_TRY(lin_ln(b.w_qkv_ln, b.qkv_colsum, b.qkv_const, w.qkv, 3 * W, PCD_EPI_LN_BIAS));
      pcd_attn_operand q = {w.qkv, (int64_t)L * 3 * W, 3 * W, 3 * 64};
      pcd_attn_operand k = {(const unsigned char*)w.qkv + 64 * es, (int64_t)L * 3 * W, 3 * W, 3 * 64};
      pcd_attn_operand v = {(const unsigned char*)w.qkv + 128 * es, (int64_t)L * 3 * W, 3 * W, 3 * 64};
      PCD_TRY(pcd_attention(&q, &k, &v, w.att, (int64_t)L * W, W, seqs, d.heads, L, L, qk_scale, qk_scale, nullptr, prec, attn_variant, stream));
      PCD_TRY(lin_res(w.att, W, b.w_proj, b.b_proj));
      PCD_TRY(lin_ln(b.w_fc_ln, b.fc_colsum, b.fc_const, w.hid, 4 * W, PCD_EPI_LN_BIAS_GELU));
      PCD_TRY(lin_res(w.hid, 4 * W, b.w_fc2, b.b_fc2));
    }
    PCD_TRY(pcd_output_proj(w.h, nullptr, prec, seqs, d.n_prefix, d.n_points, W, d.ln_post_g, d.ln_post_b, d.ln_eps,
                            d.out_w, d.out_b, out_channels, out, stream));
    return PCD_OK;
  }
  for (int l = 0; l < d.layers; ++l) {
    const pcd_block_weights& b = m->blocks[l];
    if (l == 0) {
      PCD_TRY(pcd_layernorm(w.h, W, b.ln1_g, b.ln1_b, w.xn, W, prec, M, W, d.ln_eps, stream));
    } else {  // pending MLP output of the previous block
      PCD_TRY(pcd_add_layernorm(w.h, W, w.y, W, prec, b.ln1_g, b.ln1_b, w.xn, W, prec, M, W, d.ln_eps, stream));
    }
    PCD_TRY(gemm(w.xn, W, b.w_qkv, b.b_qkv, w.qkv, 3 * W, W, PCD_EPI_BIAS));
    pcd_attn_operand q = {w.qkv, (int64_t)L * 3 * W, 3 * W, 3 * 64};
    pcd_attn_operand k = {(const unsigned char*)w.qkv + 64 * es, (int64_t)L * 3 * W, 3 * W, 3 * 64};
    pcd_attn_operand v = {(const unsigned char*)w.qkv + 128 * es, (int64_t)L * 3 * W, 3 * W, 3 * 64};
    PCD_TRY(pcd_attention(&q, &k, &v, w.att, (int64_t)L * W, W, seqs, d.heads, L, L, qk_scale, qk_scale, nullptr, prec, attn_variant, stream));
    PCD_TRY(gemm(w.att, W, b.w_proj, b.b_proj, w.y, W, W, PCD_EPI_BIAS));
    PCD_TRY(pcd_add_layernorm(w.h, W, w.y, W, prec, b.ln2_g, b.ln2_b, w.xn, W, prec, M, W, d.ln_eps, stream));
    PCD_TRY(gemm(w.xn, W, b.w_fc, b.b_fc, w.hid, 4 * W, W, PCD_EPI_BIAS_GELU));
    PCD_TRY(gemm(w.hid, 4 * W, b.w_fc2, b.b_fc2, w.y, W, 4 * W, PCD_EPI_BIAS));
  }
  // (+ last MLP output) + ln_post + slice + output_proj + permute (transformer.py:222-226)
  PCD_TRY(pcd_output_proj(w.h, w.y, prec, seqs, d.n_prefix, d.n_points, W, d.ln_post_g, d.ln_post_b, d.ln_eps, d.out_w,
                          d.out_b, out_channels, out, stream));
  return PCD_OK;
}
