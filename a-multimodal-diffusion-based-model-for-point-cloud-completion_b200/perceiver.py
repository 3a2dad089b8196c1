"""Cross-attention blocks on the kernels (reference models/perceiver.py:11-146).

Same module tree / ``state_dict`` keys as the reference's ``SimplePerceiver``.  A block is

    x += c_proj(attn(c_q(ln_1 x), c_kv(ln_2 data)));  x += mlp(ln_3 x)          (perceiver.py:70-104)

and runs, in bf16 mode, without a single LayerNorm launch on the query stream: ``ln_1`` and ``ln_3`` are folded
into the ``c_q`` / ``mlp.c_fc`` projections (``PCD_EPI_LN_BIAS`` / ``_GELU`` on the bf16 copy of the stream and its
row statistics), ``c_proj`` / ``mlp.c_proj`` update the fp32 stream in place and emit that copy + statistics
(``PCD_EPI_RESIDUAL_STATS``) -- five launches per block (q, flash cross-attention, proj, fc, proj).  The key / value
side ``c_kv(ln_2(data))`` does not depend on the query stream: inside a sampler the conditioning ``data`` (e.g. the
77 CLIP text tokens of BASELINE config 3) is the same tensor at every one of the 254 evaluations, so it is computed
once per distinct ``data`` tensor and cached (the reference recomputes it in every call).  Weight copies (bf16,
LayerNorm-folded) are made once per parameter version.
"""
import math
from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn

from . import _lib, ops
from .transformer import LN_EPS, MLP, fold_layernorm_into_linear, init_linear


def _fp32_linear(fan_in: int, fan_out: int, device, init_scale: float) -> nn.Linear:
    """Parameters live in fp32 whatever the compute dtype; the kernels take bf16 copies of the matrices."""
    layer = nn.Linear(fan_in, fan_out, device=device, dtype=torch.float32)
    init_linear(layer, init_scale)
    return layer


def _fp32_norm(width: int, device) -> nn.LayerNorm:
    return nn.LayerNorm(width, device=device, dtype=torch.float32)


class MultiheadCrossAttention(nn.Module):
    """Parameter container: queries from the stream (c_q), keys | values from the data (c_kv), c_proj."""

    def __init__(self, *, device, dtype, n_data: int, width: int, heads: int, init_scale: float,
                 data_width: Optional[int] = None):
        super().__init__()
        self.n_data, self.width, self.heads = n_data, width, heads
        self.data_width = data_width or width
        for name, fan_in, fan_out in (("c_q", width, width), ("c_kv", self.data_width, 2 * width), ("c_proj", width, width)):
            setattr(self, name, _fp32_linear(fan_in, fan_out, device, init_scale))


class ResidualCrossAttentionBlock(nn.Module):
    """x += c_proj(attn(c_q(ln_1 x), c_kv(ln_2 data))); x += mlp(ln_3 x)   (perceiver.py:70-104)."""

    def __init__(self, *, device, dtype, n_data: int, width: int, heads: int,
                 data_width: Optional[int] = None, init_scale: float = 1.0):
        super().__init__()
        if width != heads * 64:
            raise ValueError("the attention kernels are built for head dim 64")
        data_width = data_width or width
        self.compute_dtype = dtype
        self.attn = MultiheadCrossAttention(device=device, dtype=dtype, n_data=n_data, width=width, heads=heads,
                                            data_width=data_width, init_scale=init_scale)
        self.ln_1, self.ln_2 = _fp32_norm(width, device), _fp32_norm(data_width, device)
        self.mlp = MLP(device=device, dtype=dtype, width=width, init_scale=init_scale)
        self.ln_3 = _fp32_norm(width, device)
        self._packed: Optional[Tuple] = None        # (parameter versions, dict of kernel-ready weights)
        self._kv_cache: Dict[Tuple, Tuple] = {}     # data identity -> (data kept alive, c_kv(ln_2 data))

    # ---- weights, once per parameter version ----
    def _weights(self) -> dict:
        key = tuple((p.data_ptr(), p._version) for p in self.parameters())
        if self._packed is None or self._packed[0] != key:
            cd = self.compute_dtype
            cast = lambda lin: lin.weight.detach().to(cd).contiguous()
            w = dict(q=cast(self.attn.c_q), kv=cast(self.attn.c_kv), proj=cast(self.attn.c_proj),
                     fc=cast(self.mlp.c_fc), fc2=cast(self.mlp.c_proj))
            if cd == torch.bfloat16 and self.attn.width % 256 == 0:
                d = lambda t: t.detach()
                w["q_ln"] = fold_layernorm_into_linear(d(self.attn.c_q.weight), d(self.attn.c_q.bias),
                                                       d(self.ln_1.weight), d(self.ln_1.bias))
                w["fc_ln"] = fold_layernorm_into_linear(d(self.mlp.c_fc.weight), d(self.mlp.c_fc.bias),
                                                        d(self.ln_3.weight), d(self.ln_3.bias))
            self._packed = (key, w)
            self._kv_cache.clear()
        return self._packed[1]

    # ---- key / value side: step-invariant ----
    def keys_values(self, data: torch.Tensor) -> torch.Tensor:
        """c_kv(ln_2(data)) -> [B, Lkv, 2 * width] in the compute dtype, cached per distinct ``data`` tensor
        (identity, shape and version counter; the tensor is kept alive so its address cannot be recycled)."""
        w = self._weights()
        key = (data.data_ptr(), tuple(data.shape), tuple(data.stride()), data._version, data.dtype)
        hit = self._kv_cache.get(key)
        if hit is None:
            batch, n_kv, _ = data.shape
            normed = ops.layernorm(data.float().contiguous(), self.ln_2.weight, self.ln_2.bias, LN_EPS,
                                   out_dtype=self.compute_dtype)
            kv = ops.linear(normed, w["kv"], self.attn.c_kv.bias).view(batch, n_kv, 2 * self.attn.width)
            for old in list(self._kv_cache)[:-3]:  # a few conditioning tensors at most (cond / uncond halves)
                del self._kv_cache[old]
            hit = (data, kv)
            self._kv_cache[key] = hit
        return hit[1]

    # ---- query stream ----
    def forward(self, x: torch.Tensor, data: torch.Tensor) -> torch.Tensor:
        batch, n_q, width = x.shape
        stream = x.float().contiguous().clone().view(batch * n_q, width)
        state = ops.cast_rowstats(stream) if self._folded(batch * n_q) else None
        stream, _ = self.step(stream, state, self.keys_values(data), batch)
        return stream.view(batch, n_q, width)

    def _folded(self, rows: int) -> bool:
        return self.compute_dtype == torch.bfloat16 and self.attn.width % 256 == 0 and rows >= 512

    def step(self, stream: torch.Tensor, state, keys_values: torch.Tensor, batch: int):
        """One block on the fp32 stream [rows, width], updated IN PLACE.  ``state`` = (bf16 copy of the stream, its
        row statistics) on the LayerNorm-folded path (returned for the next block), None otherwise."""
        w = self._weights()
        rows, width = stream.shape
        n_q = rows // batch
        heads = self.attn.heads
        if state is not None:
            copy, stats = state
            queries = ops.linear_layernorm_folded(copy, stats, *w["q_ln"], eps=LN_EPS)
            attended = ops.cross_attention(queries.view(batch, n_q, width), keys_values, heads)
            copy, stats = ops.linear_residual_stats(attended.view(rows, width), w["proj"], self.attn.c_proj.bias, stream)
            hidden = ops.linear_layernorm_folded(copy, stats, *w["fc_ln"], eps=LN_EPS, gelu=True)
            return stream, ops.linear_residual_stats(hidden, w["fc2"], self.mlp.c_proj.bias, stream)
        cd = self.compute_dtype
        normed = ops.layernorm(stream, self.ln_1.weight, self.ln_1.bias, LN_EPS, out_dtype=cd)
        queries = ops.linear(normed, w["q"], self.attn.c_q.bias)
        attended = ops.cross_attention(queries.view(batch, n_q, width), keys_values, heads)
        stream = ops.linear(attended.view(rows, width), w["proj"], self.attn.c_proj.bias, residual=stream,
                            out_dtype=torch.float32)
        normed = ops.layernorm(stream, self.ln_3.weight, self.ln_3.bias, LN_EPS, out_dtype=cd)
        hidden = ops.linear(normed, w["fc"], self.mlp.c_fc.bias, epilogue=_lib.EPI_BIAS_GELU)
        return ops.linear(hidden, w["fc2"], self.mlp.c_proj.bias, residual=stream, out_dtype=torch.float32), None


class SimplePerceiver(nn.Module):
    """A stack of cross-attention blocks over fixed data tokens (reference models/perceiver.py:107-146)."""

    def __init__(self, *, device, dtype=torch.bfloat16, n_data: int, width: int, layers: int, heads: int,
                 init_scale: float = 0.25, data_width: Optional[int] = None):
        super().__init__()
        self.width, self.layers = width, layers
        block_args = dict(device=device, dtype=dtype, n_data=n_data, width=width, heads=heads, data_width=data_width,
                          init_scale=init_scale * math.sqrt(1.0 / width))
        self.resblocks = nn.ModuleList(ResidualCrossAttentionBlock(**block_args) for _ in range(layers))

    @torch.no_grad()
    def forward(self, x: torch.Tensor, data: torch.Tensor):
        batch, n_q, width = x.shape
        stream = x.float().contiguous().clone().view(batch * n_q, width)
        first = self.resblocks[0]
        state = ops.cast_rowstats(stream) if first._folded(batch * n_q) else None
        for block in self.resblocks:
            stream, state = block.step(stream, state, block.keys_values(data), batch)
        return stream.view(batch, n_q, width)
