"""Cross-attention blocks on the kernels (reference models/perceiver.py:11-146).

Same module tree / ``state_dict`` keys as the reference's ``SimplePerceiver``; each
block = LayerNorm kernels -> c_q / c_kv projections -> flash cross-attention
(q [B,Lq,H,64], kv [B,Lkv,H,(k|v),64]) -> c_proj + residual -> MLP.
"""
import math
from typing import Optional

import torch
import torch.nn as nn

from . import _lib, ops
from .transformer import LN_EPS, MLP, init_linear


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

    def _project(self, acts, layer: nn.Linear, **kw):
        return ops.linear(acts, layer.weight.to(self.compute_dtype), layer.bias, **kw)

    def _normed(self, stream, norm: nn.LayerNorm):
        return ops.layernorm(stream, norm.weight, norm.bias, LN_EPS, out_dtype=self.compute_dtype)

    def forward(self, x: torch.Tensor, data: torch.Tensor) -> torch.Tensor:
        batch, n_q, width = x.shape
        stream = x.float().contiguous()
        queries = self._project(self._normed(stream, self.ln_1), self.attn.c_q).view(batch, n_q, width)
        keys_values = self._project(self._normed(data.float().contiguous(), self.ln_2), self.attn.c_kv)
        attended = ops.cross_attention(queries, keys_values.view(batch, data.shape[1], 2 * width), self.attn.heads)
        stream = self._project(attended, self.attn.c_proj, residual=stream, out_dtype=torch.float32).view(batch, n_q, width)
        hidden = self._project(self._normed(stream, self.ln_3), self.mlp.c_fc, epilogue=_lib.EPI_BIAS_GELU)
        return self._project(hidden, self.mlp.c_proj, residual=stream, out_dtype=torch.float32).view(batch, n_q, width)


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
        for block in self.resblocks:
            x = block(x, data)
        return x
