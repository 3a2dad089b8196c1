// Host side of the tensor-core attention kernels: TMA tensor maps over the strided q / k / v operands and the
// per-call choice of kernel (include/pcd_b200.h: pcd_attention, `variant`).
#include "common.cuh"
#include "tc_sm100.cuh"

namespace pcd {

int launch_attn_tc5(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, uint16_t* out, int64_t o_bs,
                    int64_t o_ls, int batch, int heads, int len_q, int len_kv, float scale_log2, int poly,
                    cudaStream_t st);  // attn_tc5.cu
int launch_attn_tc8(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& to, int batch, int heads, int len_q, int len_kv, float scale_log2, const float* rope,
                    int mode, cudaStream_t st);  // attn_tc8.cu (mode: 0 default, 1 MUFU token ring, 2 tail keys as a KV step, 3 wide)
int attn_tc8_kv_rows(int mode);                   // keys per K / V tensor-map box of a mode
int launch_rope_bf16(uint16_t* x, int64_t bs, int64_t ls, int64_t hs, const float* coords, int batch, int heads,
                     int len, cudaStream_t st);  // attn_simt.cu

// element (c, h, l, b) of an operand: head_dim contiguous head dims, then heads, rows, sequences at their strides.
// The box is always 64 columns wide (the kernels' tile): with head_dim = 32 its columns 32..63 lie outside the tensor, so
// TMA loads fill them with zeros and TMA stores drop them -- 32-wide heads run on the 64-wide kernel without any padded
// copy in memory (q . k is unchanged by zero columns, the upper half of P V is zero and never written).
static int make_operand_map(CUtensorMap* m, const pcd_attn_operand* op, int batch, int heads, int len, int box_rows,
                            int head_dim = 64) {
  uint64_t dims[4] = {(uint64_t)head_dim, (uint64_t)heads, (uint64_t)len, (uint64_t)batch};
  uint64_t strides[3] = {(uint64_t)op->head_stride * 2, (uint64_t)op->row_stride * 2, (uint64_t)op->batch_stride * 2};
  uint32_t box[4] = {64, 1, (uint32_t)box_rows, 1};
  return encode_tmap_bf16(m, op->ptr, 4, dims, strides, box);
}

int launch_attention_bf16(const pcd_attn_operand* q, const pcd_attn_operand* k, const pcd_attn_operand* v, uint16_t* out,
                          int64_t o_bs, int64_t o_ls, int batch, int heads, int len_q, int len_kv, float q_scale,
                          float k_scale, const float* rope, int variant, int head_dim, cudaStream_t st) {
  if (head_dim != 64 && head_dim != 32) {
    set_error("attention(bf16): head dim %d (64 or 32)", head_dim);
    return PCD_ERR_INVALID;
  }
  if (head_dim == 32 && (variant == PCD_ATTN_DEFAULT || rope != nullptr)) variant = PCD_ATTN_GROUPED;  // TMA-store epilogue
  if (variant == PCD_ATTN_DEFAULT) {
    // three query tiles per CTA need enough (group, head, sequence) items to occupy every SM; small problems keep
    // the finer-grained paired kernel (one query tile per CTA, two CTAs per SM)
    const int64_t nq = (len_q + 127) / 128, grouped_items = ((nq + 2) / 3) * heads * batch;
    variant = (rope != nullptr || grouped_items >= num_sms()) ? PCD_ATTN_GROUPED : PCD_ATTN_PAIRED;
  }
  const bool grouped = variant == PCD_ATTN_GROUPED || variant == PCD_ATTN_GROUPED_TOKEN ||
                       variant == PCD_ATTN_GROUPED_STEPTAIL || variant == PCD_ATTN_GROUPED_WIDE;
  const int mode = variant == PCD_ATTN_GROUPED_TOKEN ? 1 : (variant == PCD_ATTN_GROUPED_STEPTAIL ? 2 : (variant == PCD_ATTN_GROUPED_WIDE ? 3 : 0));
  if (!grouped && variant != PCD_ATTN_PAIRED &&
      variant != PCD_ATTN_PAIRED_POLY4 && variant != PCD_ATTN_PAIRED_POLY2) {
    set_error("attention(bf16): unknown kernel variant %d", variant);
    return PCD_ERR_INVALID;
  }
  CUtensorMap tq, tk, tv;
  int rc;
  if (head_dim == 32 && !grouped) {
    set_error("attention(bf16): 32-wide heads need a grouped kernel variant (its output leaves through TMA)");
    return PCD_ERR_UNSUPPORTED;
  }
  if ((rc = make_operand_map(&tq, q, batch, heads, len_q, 128, head_dim)) != PCD_OK) return rc;
  const int kv_rows = grouped ? attn_tc8_kv_rows(mode) : 64;
  if ((rc = make_operand_map(&tk, k, batch, heads, len_kv, kv_rows, head_dim)) != PCD_OK) return rc;
  if ((rc = make_operand_map(&tv, v, batch, heads, len_kv, kv_rows, head_dim)) != PCD_OK) return rc;
  const float scale_log2 = q_scale * k_scale * 1.4426950408889634f;
  if (grouped) {
    // the output through the same kind of map: [batch, len_q, heads, 64] at the caller's strides, 32-row boxes
    CUtensorMap to;
    const pcd_attn_operand oo = {out, o_bs, o_ls, head_dim};
    if ((rc = make_operand_map(&to, &oo, batch, heads, len_q, 32, head_dim)) != PCD_OK) return rc;
    return launch_attn_tc8(tq, tk, tv, to, batch, heads, len_q, len_kv, scale_log2, rope, mode, st);
  }
  if (rope != nullptr) {
    set_error("attention(bf16): the paired kernel takes pre-rotated operands (pcd_rope_bf16); use PCD_ATTN_GROUPED");
    return PCD_ERR_UNSUPPORTED;
  }
  return launch_attn_tc5(tq, tk, tv, out, o_bs, o_ls, batch, heads, len_q, len_kv, scale_log2,
                         variant == PCD_ATTN_PAIRED_POLY4 ? 2 : (variant == PCD_ATTN_PAIRED_POLY2 ? 1 : 0), st);
}

}  // namespace pcd
