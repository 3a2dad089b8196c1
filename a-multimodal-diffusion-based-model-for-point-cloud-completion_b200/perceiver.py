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


class MultiheadCrossAttention(nn.Module):
    def __init__(self, *, device, dtype, n_data: int, width: int, heads: int, init_scale: float,
                 data_width: Optional[int] = None):
        super().__init__()
        self.n_data, self.width, self.heads = n_data, width, heads
        self.data_width = width if data_width is None else data_width
        self.c_q = nn.Linear(width, width, device=device, dtype=torch.float32)
        self.c_kv = nn.Linear(self.data_width, width * 2, device=device, dtype=torch.float32)
        self.c_proj = nn.Linear(width, width, device=device, dtype=torch.float32)
        init_linear(self.c_q, init_scale)
        init_linear(self.c_kv, init_scale)
        init_linear(self.c_proj, init_scale)


class ResidualCrossAttentionBlock(nn.Module):
    def __init__(self, *, device, dtype, n_data: int, width: int, heads: int,
                 data_width: Optional[int] = None, init_scale: float = 1.0):
        super().__init__()
        if data_width is None:
            data_width = width
        if width != heads * 64:
            raise ValueError("the attention kernels are built for head dim 64")
        self.compute_dtype = dtype
        self.attn = MultiheadCrossAttention(device=device, dtype=dtype, n_data=n_data, width=width,
                                            heads=heads, data_width=data_width, init_scale=init_scale)
        self.ln_1 = nn.LayerNorm(width, device=device, dtype=torch.float32)
        self.ln_2 = nn.LayerNorm(data_width, device=device, dtype=torch.float32)
        self.mlp = MLP(device=device, dtype=dtype, width=width, init_scale=init_scale)
        self.ln_3 = nn.LayerNorm(width, device=device, dtype=torch.float32)

    def forward(self, x: torch.Tensor, data: torch.Tensor) -> torch.Tensor:
        cd = self.compute_dtype
        w = lambda lin: lin.weight.to(cd)
        B, Lq, W = x.shape
        x = x.float().contiguous()
        q_in = ops.layernorm(x, self.ln_1.weight, self.ln_1.bias, LN_EPS, out_dtype=cd)
        kv_in = ops.layernorm(data.float().contiguous(), self.ln_2.weight, self.ln_2.bias, LN_EPS, out_dtype=cd)
        q = ops.linear(q_in, w(self.attn.c_q), self.attn.c_q.bias).view(B, Lq, W)
        kv = ops.linear(kv_in, w(self.attn.c_kv), self.attn.c_kv.bias).view(B, data.shape[1], 2 * W)
        att = ops.cross_attention(q, kv, self.attn.heads)
        x = ops.linear(att, w(self.attn.c_proj), self.attn.c_proj.bias, residual=x,
                       out_dtype=torch.float32).view(B, Lq, W)
        m = ops.layernorm(x, self.ln_3.weight, self.ln_3.bias, LN_EPS, out_dtype=cd)
        hdn = ops.linear(m, w(self.mlp.c_fc), self.mlp.c_fc.bias, epilogue=_lib.EPI_BIAS_GELU)
        return ops.linear(hdn, w(self.mlp.c_proj), self.mlp.c_proj.bias, residual=x,
                          out_dtype=torch.float32).view(B, Lq, W)


class SimplePerceiver(nn.Module):
    """Only does cross attention (reference models/perceiver.py:107-146)."""

    def __init__(self, *, device, dtype=torch.bfloat16, n_data: int, width: int, layers: int, heads: int,
                 init_scale: float = 0.25, data_width: Optional[int] = None):
        super().__init__()
        self.width, self.layers = width, layers
        init_scale = init_scale * math.sqrt(1.0 / width)
        self.resblocks = nn.ModuleList([
            ResidualCrossAttentionBlock(device=device, dtype=dtype, n_data=n_data, width=width, heads=heads,
                                        init_scale=init_scale, data_width=data_width)
            for _ in range(layers)])

    @torch.no_grad()
    def forward(self, x: torch.Tensor, data: torch.Tensor):
        for block in self.resblocks:
            x = block(x, data)
        return x
