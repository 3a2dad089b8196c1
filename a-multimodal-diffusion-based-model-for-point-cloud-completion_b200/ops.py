"""Tensor-level wrappers over the C ABI (one ctypes call per op, current CUDA stream).

PyTorch is used here for device memory and streams only; all arithmetic happens
in libpcd_b200.so.  Every function raises ``PcdError`` instead of falling back.
"""
import ctypes as C
import math
from typing import Optional

import torch

from . import _lib
from ._lib import (AttnOperand, EPI_BIAS, EPI_BIAS_GELU, EPI_BIAS_RESIDUAL, PCD_BF16, PCD_F32,
                   StepScalars, check, ptr, require_cuda, stream_ptr)

_PREC = {torch.float32: PCD_F32, torch.bfloat16: PCD_BF16}


def _rowmajor2d(t: torch.Tensor):
    assert t.dim() == 2 and t.stride(1) == 1, "expected a row-major 2-D tensor"
    return t.stride(0)


def timestep_freqs(dim: int, max_period: float = 10000.0, device=None) -> torch.Tensor:
    """Frequency table of ``timestep_embedding`` computed on the CPU exactly like the
    reference does (models/util.py:81-83) and moved to the device."""
    half = dim // 2
    f = torch.exp(-math.log(max_period) * torch.arange(start=0, end=half, dtype=torch.float32) / half)
    return f.to(device) if device is not None else f


def timestep_embedding(t: torch.Tensor, dim: int, freqs: Optional[torch.Tensor] = None) -> torch.Tensor:
    """reference models/util.py:72-89."""
    require_cuda(t)
    tf = t.to(torch.float32).contiguous()
    if freqs is None:
        freqs = timestep_freqs(dim, device=t.device)
    out = torch.empty(t.shape[0], dim, device=t.device, dtype=torch.float32)
    check(_lib.load().pcd_timestep_embed(ptr(tf), ptr(freqs), t.shape[0], dim, ptr(out), dim, stream_ptr()),
          "timestep_embed")
    return out


def layernorm(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, eps: float = 1e-5,
              out_dtype=torch.float32, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """nn.LayerNorm over the last dim of an fp32 tensor -> fp32 or bf16."""
    require_cuda(x, weight, bias)
    assert x.dtype == torch.float32
    dim = x.shape[-1]
    x2 = x.reshape(-1, dim)
    if x2.stride(1) != 1:
        x2 = x2.contiguous()
    if out is None:
        out = torch.empty(x2.shape, device=x.device, dtype=out_dtype)
    o2 = out.reshape(-1, dim)
    check(_lib.load().pcd_layernorm(ptr(x2), x2.stride(0), ptr(weight), ptr(bias), ptr(o2), o2.stride(0),
                                    _PREC[out.dtype], x2.shape[0], dim, float(eps), stream_ptr()), "layernorm")
    return out.reshape(*x.shape[:-1], dim)


def add_layernorm(h: torch.Tensor, y: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor,
                  eps: float = 1e-5, out_dtype=torch.float32) -> torch.Tensor:
    """Fused residual update + LayerNorm: h += y IN PLACE (h fp32, y bf16/fp32), returns LN(h)."""
    require_cuda(h, y, weight, bias)
    assert h.dtype == torch.float32 and h.is_contiguous() and y.is_contiguous() and y.shape == h.shape
    dim = h.shape[-1]
    h2, y2 = h.view(-1, dim), y.view(-1, dim)
    out = torch.empty(h2.shape, device=h.device, dtype=out_dtype)
    check(_lib.load().pcd_add_layernorm(ptr(h2), dim, ptr(y2), dim, _PREC[y.dtype], ptr(weight), ptr(bias),
                                        ptr(out), dim, _PREC[out_dtype], h2.shape[0], dim, float(eps),
                                        stream_ptr()), "add_layernorm")
    return out.view(h.shape)


def linear(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None, *,
           epilogue: int = EPI_BIAS, residual: Optional[torch.Tensor] = None,
           out_dtype=None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """y = epi(x W^T + b).  fp32 operands -> CUDA-core GEMM; bf16 operands -> tcgen05 GEMM
    (bias / residual stay fp32, output bf16 or fp32)."""
    require_cuda(x, weight)
    assert x.dtype == weight.dtype and x.dtype in _PREC
    K = x.shape[-1]
    N = weight.shape[0]
    assert weight.shape[1] == K
    x2 = x.reshape(-1, K)
    if x2.stride(1) != 1:
        x2 = x2.contiguous()
    M = x2.shape[0]
    if out_dtype is None:
        out_dtype = x.dtype
    if out is None:
        out = torch.empty(M, N, device=x.device, dtype=out_dtype)
    o2 = out.reshape(-1, N) if out.dim() != 2 else out
    r2 = None
    ldr = 0
    if residual is not None:
        assert residual.dtype == torch.float32
        if epilogue == EPI_BIAS_GELU:
            raise ValueError("linear: a residual cannot be combined with the GELU epilogue")
        r2 = residual.reshape(-1, N) if residual.dim() != 2 else residual
        ldr = _rowmajor2d(r2)
        epilogue = EPI_BIAS_RESIDUAL
    lib = _lib.load()
    if x.dtype == torch.float32:
        assert o2.dtype == torch.float32
        check(lib.pcd_gemm_f32(ptr(x2), x2.stride(0), ptr(weight), _rowmajor2d(weight), ptr(bias), ptr(r2), ldr,
                               ptr(o2), _rowmajor2d(o2), M, N, K, epilogue, stream_ptr()), "gemm_f32")
    else:
        check(lib.pcd_gemm_bf16(ptr(x2), x2.stride(0), ptr(weight), _rowmajor2d(weight), ptr(bias), ptr(r2), ldr,
                                ptr(o2), _rowmajor2d(o2), _PREC[o2.dtype], M, N, K, epilogue, stream_ptr()),
              "gemm_bf16")
    return out.reshape(*x.shape[:-1], N) if out.dim() == 2 and x.dim() != 2 else out


def launch_count() -> int:
    """Kernels this library has launched so far in the process (pcd_launch_count)."""
    return int(_lib.load().pcd_launch_count())


def add(a: torch.Tensor, b: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out = a + b (fp32, same shape, contiguous; ``out`` may be ``a``)."""
    require_cuda(a, b)
    assert a.dtype == torch.float32 and b.dtype == torch.float32 and a.shape == b.shape
    a, b = a.contiguous(), b.contiguous()
    if out is None:
        out = torch.empty_like(a)
    check(_lib.load().pcd_add_f32(ptr(a), ptr(b), ptr(out), a.numel(), stream_ptr()), "add_f32")
    return out


def attention_hd32(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, heads: int) -> torch.Tensor:
    """softmax(q k^T * 32^-1/2) v per head for head dim 32 (TwoStream CrossAttention, reference
    models/modules.py:40-63): q [B, Lq, H*32], k / v [B, Lkv, H*32] fp32 (views with a unit last stride are fine)."""
    require_cuda(q, k, v)
    assert q.dtype == k.dtype == v.dtype == torch.float32
    B, Lq, W = q.shape
    Lkv = k.shape[1]
    assert W == heads * 32 and k.shape == v.shape == (B, Lkv, W)
    for t in (q, k, v):
        assert t.stride(2) == 1
    out = torch.empty(B, Lq, W, device=q.device, dtype=torch.float32)
    s = 32.0 ** -0.25
    opnd = lambda t: _operand(t, 0, t.stride(0), t.stride(1), 32)
    qo, ko, vo = opnd(q), opnd(k), opnd(v)
    check(_lib.load().pcd_attention_hd32(C.byref(qo), C.byref(ko), C.byref(vo), ptr(out), out.stride(0), out.stride(1),
                                         B, heads, Lq, Lkv, s, s, stream_ptr()), "attention_hd32")
    return out


def embed_tokens(x: torch.Tensor, w_in: torch.Tensor, b_in: torch.Tensor, prefix: Optional[torch.Tensor],
                 add_cond: Optional[torch.Tensor], ln_g: torch.Tensor, ln_b: torch.Tensor, eps: float = 1e-5,
                 seqs: Optional[int] = None, with_stats: bool = False):
    """input_proj + token concat + ln_pre (reference transformer.py:208-220).  x [x_seqs, c_in, n_points] (NCL),
    prefix [seqs, n_prefix, dim] or None, add_cond [seqs, dim] or None -> h [seqs, n_prefix + n_points, dim] fp32;
    with_stats also returns (bf16 copy of h, per-128-column (mean, M2) [rows, dim/128, 2]) from the same launch."""
    require_cuda(x, w_in, b_in, ln_g, ln_b)
    x_seqs, c_in, n_points = x.shape
    dim = w_in.shape[0]
    seqs = x_seqs if seqs is None else seqs
    n_prefix = 0 if prefix is None else prefix.shape[1]
    assert x.is_contiguous() and w_in.is_contiguous() and (prefix is None or prefix.is_contiguous())
    assert add_cond is None or (add_cond.is_contiguous() and add_cond.shape == (seqs, dim))
    L = n_prefix + n_points
    h = torch.empty(seqs, L, dim, device=x.device, dtype=torch.float32)
    hb = stats = None
    if with_stats:
        hb = torch.empty(seqs, L, dim, device=x.device, dtype=torch.bfloat16)
        stats = torch.empty(seqs * L, dim // 128, 2, device=x.device, dtype=torch.float32)
    opt = lambda t: ptr(t) if t is not None else None
    check(_lib.load().pcd_embed_tokens(ptr(x), x_seqs, c_in, n_points, ptr(w_in), ptr(b_in), opt(prefix), n_prefix,
                                       opt(add_cond), ptr(ln_g), ptr(ln_b), float(eps), ptr(h), seqs, dim, opt(hb),
                                       opt(stats), stream_ptr()), "embed_tokens")
    return (h, hb, stats) if with_stats else h


def output_proj(h: torch.Tensor, n_prefix: int, ln_g: torch.Tensor, ln_b: torch.Tensor, w_out: torch.Tensor,
                b_out: torch.Tensor, eps: float = 1e-5, y: Optional[torch.Tensor] = None) -> torch.Tensor:
    """ln_post + drop the prefix tokens + output_proj + permute (reference transformer.py:222-226).
    h [seqs, L, dim] fp32 (+ y, a pending residual of the same shape, bf16 or fp32) -> [seqs, c_out, L - n_prefix]."""
    require_cuda(h, ln_g, ln_b, w_out, b_out, y)
    assert h.dtype == torch.float32 and h.is_contiguous() and w_out.is_contiguous()
    seqs, L, dim = h.shape
    c_out = w_out.shape[0]
    out = torch.empty(seqs, c_out, L - n_prefix, device=h.device, dtype=torch.float32)
    yprec = _PREC[y.dtype] if y is not None else PCD_F32
    check(_lib.load().pcd_output_proj(ptr(h), ptr(y), yprec, seqs, n_prefix, L - n_prefix, dim, ptr(ln_g), ptr(ln_b),
                                      float(eps), ptr(w_out), ptr(b_out), c_out, ptr(out), stream_ptr()), "output_proj")
    return out


def cast_rowstats(h: torch.Tensor):
    """bf16 copy of the fp32 residual stream h [rows, dim] + per-128-column (mean, M2) of the rounded
    values [rows, dim/128, 2]: the inputs of the LayerNorm-folded projections."""
    require_cuda(h)
    assert h.dtype == torch.float32 and h.dim() == 2 and h.shape[1] % 128 == 0
    rows, dim = h.shape
    hb = torch.empty(rows, dim, device=h.device, dtype=torch.bfloat16)
    stats = torch.empty(rows, dim // 128, 2, device=h.device, dtype=torch.float32)
    check(_lib.load().pcd_cast_rowstats(ptr(h), _rowmajor2d(h), ptr(hb), dim, ptr(stats), rows, dim, stream_ptr()),
          "cast_rowstats")
    return hb, stats


def linear_residual_stats(a: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, h: torch.Tensor):
    """h <- h + a W^T + b in place (fp32 residual stream, transformer.py:113-114); returns
    (bf16 copy of the updated h, its row statistics [rows, N/128, 2])."""
    require_cuda(a, weight, h)
    assert a.dtype == torch.bfloat16 and weight.dtype == torch.bfloat16 and h.dtype == torch.float32
    M, K = a.shape
    N = weight.shape[0]
    assert h.shape == (M, N)
    hb = torch.empty(M, N, device=a.device, dtype=torch.bfloat16)
    stats = torch.empty(M, N // 128, 2, device=a.device, dtype=torch.float32)
    g = _lib.GemmArgs()
    g.A, g.lda, g.W, g.ldw, g.bias = ptr(a), _rowmajor2d(a), ptr(weight), _rowmajor2d(weight), ptr(bias)
    g.residual, g.ldr, g.C, g.ldc, g.out_precision = ptr(h), _rowmajor2d(h), ptr(h), _rowmajor2d(h), _lib.PCD_F32
    g.C2, g.ldc2, g.stats_out = ptr(hb), N, ptr(stats)
    g.M, g.N, g.K, g.epilogue = M, N, K, _lib.EPI_RESIDUAL_STATS
    check(_lib.load().pcd_gemm_bf16_ex(C.byref(g), stream_ptr()), "gemm_bf16_ex(residual+stats)")
    return hb, stats


def linear_layernorm_folded(hb: torch.Tensor, stats: torch.Tensor, w_folded: torch.Tensor, colsum: torch.Tensor,
                            const: torch.Tensor, eps: float = 1e-5, gelu: bool = False) -> torch.Tensor:
    """(gelu)(LayerNorm(h) W^T + b) evaluated from the bf16 copy hb of h, its row statistics and the
    folded weights: rstd (hb (gamma o W)^T - mu colsum) + const."""
    require_cuda(hb, w_folded)
    assert hb.dtype == torch.bfloat16 and w_folded.dtype == torch.bfloat16
    M, K = hb.shape
    N = w_folded.shape[0]
    out = torch.empty(M, N, device=hb.device, dtype=torch.bfloat16)
    g = _lib.GemmArgs()
    g.A, g.lda, g.W, g.ldw, g.bias = ptr(hb), _rowmajor2d(hb), ptr(w_folded), _rowmajor2d(w_folded), ptr(const)
    g.colsum, g.stats_in, g.ln_eps = ptr(colsum), ptr(stats), eps
    g.C, g.ldc, g.out_precision = ptr(out), N, _lib.PCD_BF16
    g.M, g.N, g.K, g.epilogue = M, N, K, (_lib.EPI_LN_BIAS_GELU if gelu else _lib.EPI_LN_BIAS)
    check(_lib.load().pcd_gemm_bf16_ex(C.byref(g), stream_ptr()), "gemm_bf16_ex(layernorm-folded)")
    return out


def _operand(t: torch.Tensor, offset: int, batch_stride: int, row_stride: int, head_stride: int):
    return AttnOperand(C.c_void_p(t.data_ptr() + offset * t.element_size()), batch_stride, row_stride, head_stride)


def attention_packed(q_op, k_op, v_op, out: torch.Tensor, batch: int, heads: int, len_q: int, len_kv: int,
                     q_scale: float, k_scale: float, rope_coords: Optional[torch.Tensor] = None, variant: int = 0):
    """Low-level call: operands are ``AttnOperand`` views; ``out`` is [batch, len_q, heads*64]; ``variant`` picks
    the tensor-core kernel (``_lib.ATTN_*``, 0 = default) for this call."""
    require_cuda(out)
    check(_lib.load().pcd_attention(C.byref(q_op), C.byref(k_op), C.byref(v_op), ptr(out), out.stride(0),
                                    out.stride(1), batch, heads, len_q, len_kv, float(q_scale), float(k_scale),
                                    ptr(rope_coords), _PREC[out.dtype], int(variant), stream_ptr()), "attention")
    return out


def self_attention(qkv: torch.Tensor, heads: int, variant: int = 0) -> torch.Tensor:
    """QKVMultiheadAttention (reference models/transformer.py:65-84): qkv [B, L, H*3*64]
    laid out [H][q|k|v][64]; q and k each scaled by 64**-0.25."""
    B, L, W3 = qkv.shape
    hd = W3 // heads // 3
    assert hd == 64, "this build supports head dim 64 (all registered configs)"
    qkv = qkv.contiguous()
    out = torch.empty(B, L, heads * hd, device=qkv.device, dtype=qkv.dtype)
    s = 1.0 / math.sqrt(math.sqrt(hd))
    ops = [_operand(qkv, i * hd, L * W3, W3, 3 * hd) for i in range(3)]
    return attention_packed(ops[0], ops[1], ops[2], out, B, heads, L, L, s, s, variant=variant)


def cross_attention(q: torch.Tensor, kv: torch.Tensor, heads: int, variant: int = 0) -> torch.Tensor:
    """QKVMultiheadCrossAttention (reference models/perceiver.py:46-67): q [B, Lq, H*64],
    kv [B, Lkv, H*2*64] laid out [H][k|v][64]."""
    B, Lq, W = q.shape
    _, Lkv, W2 = kv.shape
    hd = W2 // heads // 2
    assert hd == 64 and W == heads * hd
    q, kv = q.contiguous(), kv.contiguous()
    out = torch.empty(B, Lq, W, device=q.device, dtype=q.dtype)
    s = 1.0 / math.sqrt(math.sqrt(hd))
    qo = _operand(q, 0, Lq * W, W, hd)
    ko = _operand(kv, 0, Lkv * W2, W2, 2 * hd)
    vo = _operand(kv, hd, Lkv * W2, W2, 2 * hd)
    return attention_packed(qo, ko, vo, out, B, heads, Lq, Lkv, s, s, variant=variant)


def attention_views(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, heads: int, q_scale: float,
                    k_scale: float, variant: int = 0, head_dim: int = 64) -> torch.Tensor:
    """softmax((q_scale q) (k_scale k)^T) v per head over strided views: q [B, Lq, H*head_dim], k / v [B, Lkv, H*head_dim]
    (any batch / row strides, unit last stride; e.g. column blocks of one fused projection output).  head_dim 64, or 32
    in bf16 (TwoStream's heads: the 64-wide tensor-core kernel over tensor maps that zero-fill the missing columns)."""
    B, Lq, W = q.shape
    Lkv = k.shape[1]
    assert W == heads * head_dim and k.shape == v.shape == (B, Lkv, W) and q.dtype == k.dtype == v.dtype
    for t in (q, k, v):
        assert t.stride(2) == 1
    out = torch.empty(B, Lq, W, device=q.device, dtype=q.dtype)
    opnd = lambda t: _operand(t, 0, t.stride(0), t.stride(1), head_dim)
    if head_dim == 64:
        return attention_packed(opnd(q), opnd(k), opnd(v), out, B, heads, Lq, Lkv, q_scale, k_scale, None, variant)
    assert head_dim == 32 and q.dtype == torch.bfloat16, "32-wide heads: bf16 here, fp32 through attention_hd32"
    require_cuda(q, k, v)
    qo, ko, vo = opnd(q), opnd(k), opnd(v)
    check(_lib.load().pcd_attention_hd32_bf16(C.byref(qo), C.byref(ko), C.byref(vo), ptr(out), out.stride(0), out.stride(1),
                                              B, heads, Lq, Lkv, float(q_scale), float(k_scale), int(variant), stream_ptr()),
          "attention_hd32_bf16")
    return out


def rotary_attention(qkv: torch.Tensor, coords: torch.Tensor, heads: int) -> torch.Tensor:
    """RotarySelfAttention core (reference models/rotaryencoderpcd.py:68-84): qkv [B, N, 3*D]
    laid out [3][H][64]; 3-axis RoPE on head dims 0..5 of q and k; logits * D**-0.5.  One launch in either
    precision: the rotation happens inside the attention kernel (fp32: in registers while the tiles are loaded;
    bf16: on the Q / K tiles in shared memory between their TMA arrival and the first MMA)."""
    B, N, D3 = qkv.shape
    D = D3 // 3
    assert D // heads == 64
    qkv = qkv.contiguous()
    coords = coords.to(torch.float32).contiguous()
    assert coords.shape == (B, N, 3)
    out = torch.empty(B, N, D, device=qkv.device, dtype=qkv.dtype)
    ops = [_operand(qkv, i * D, N * D3, D3, 64) for i in range(3)]
    return attention_packed(ops[0], ops[1], ops[2], out, B, heads, N, N, D ** -0.5, 1.0, coords)


def rope_bf16_(x: torch.Tensor, coords: torch.Tensor, heads: int) -> torch.Tensor:
    """Stand-alone rotation (pcd_rope_bf16): rotate head dims 0..5 of a bf16 [B, N, H*64] view IN PLACE."""
    require_cuda(x, coords)
    assert x.dtype == torch.bfloat16 and x.stride(2) == 1
    B, N, _ = x.shape
    coords = coords.to(torch.float32).contiguous()
    o = _operand(x, 0, x.stride(0), x.stride(1), 64)
    check(_lib.load().pcd_rope_bf16(C.byref(o), ptr(coords), B, heads, N, stream_ptr()), "rope_bf16")
    return x


def chamfer_distance_xyz(p1: torch.Tensor, p2: torch.Tensor) -> torch.Tensor:
    """reference models/util.py:265-295: squared-L2 Chamfer on channels 0:3 -> [B]."""
    require_cuda(p1, p2)
    p1, p2 = p1.float().contiguous(), p2.float().contiguous()
    B, c1, n1 = p1.shape
    _, c2, n2 = p2.shape
    out = torch.empty(B, device=p1.device, dtype=torch.float32)
    ws = torch.empty(B * (n1 + n2), device=p1.device, dtype=torch.float32)
    check(_lib.load().pcd_chamfer(ptr(p1), c1, n1, ptr(p2), c2, n2, B, ptr(out), ptr(ws), stream_ptr()), "chamfer")
    return out


def farthest_point_sample(points: torch.Tensor, n_samples: int, init_idx) -> torch.Tensor:
    """Greedy farthest-point sampling (reference util/point_cloud.py:82-118, evaluation.py:40-48):
    points [B, N, 3] fp32 -> int64 indices [B, n_samples]; init_idx int or [B] tensor."""
    require_cuda(points)
    assert points.dim() == 3 and points.shape[2] == 3 and points.dtype == torch.float32
    B, N, _ = points.shape
    points = points.contiguous()
    if not torch.is_tensor(init_idx):
        if not 0 <= int(init_idx) < N:
            raise IndexError(f"farthest_point_sample: init_idx {int(init_idx)} outside a cloud of {N} points")
        init_idx = torch.full((B,), int(init_idx), dtype=torch.int32, device=points.device)
    init_idx = init_idx.to(device=points.device, dtype=torch.int32).contiguous()  # (device tensors are clamped by the kernel)
    out = torch.empty(B, n_samples, dtype=torch.int64, device=points.device)
    ws = torch.empty(B, N, dtype=torch.float32, device=points.device) if N > 8192 else None  # running distances
    check(_lib.load().pcd_farthest_point_sample(ptr(points), B, N, n_samples, ptr(init_idx), ptr(out), ptr(ws), stream_ptr()),
          "farthest_point_sample")
    return out


def nearest_points(queries: torch.Tensor, cloud: torch.Tensor, form: int = 1):
    """For every query [B, Nq, 3] the index of / squared distance to its nearest cloud point [B, Nc, 3]
    (PointCloud.nearest_points, reference util/point_cloud.py:148-165)."""
    require_cuda(queries, cloud)
    assert queries.dtype == torch.float32 and cloud.dtype == torch.float32
    queries, cloud = queries.contiguous(), cloud.contiguous()
    B, Nq, _ = queries.shape
    idx = torch.empty(B, Nq, dtype=torch.int64, device=queries.device)
    d2 = torch.empty(B, Nq, dtype=torch.float32, device=queries.device)
    check(_lib.load().pcd_nearest_points(ptr(queries), Nq, ptr(cloud), cloud.shape[1], B, form, ptr(d2), ptr(idx),
                                         stream_ptr()), "nearest_points")
    return idx, d2


def fscore_point_cloud_batch(pred: torch.Tensor, gt: torch.Tensor, threshold: float = 0.03, squared: bool = False):
    """F-score / precision / recall of batched clouds (reference models/util.py:195-262; squared=True is
    fscore_point_cloud_batch_squared): pred [B, N, 3], gt [B, M, 3] -> three [B] tensors."""
    require_cuda(pred, gt)
    pred, gt = pred.float().contiguous(), gt.float().contiguous()
    B, N, _ = pred.shape
    M = gt.shape[1]
    out = torch.empty(3, B, dtype=torch.float32, device=pred.device)
    ws = torch.empty(B * (N + M), dtype=torch.float32, device=pred.device)
    check(_lib.load().pcd_fscore(ptr(pred), N, ptr(gt), M, B, float(threshold), int(squared), ptr(out), ptr(ws),
                                 stream_ptr()), "fscore")
    return out[0], out[1], out[2]


def step_scalars(**kw) -> StepScalars:
    s = StepScalars()
    for k, v in kw.items():
        setattr(s, k, float(v))
    return s
