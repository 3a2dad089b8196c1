"""Rotary point encoding (reference models/rotaryencoderpcd.py:6-27, 58-84).

``RotarySelfAttention`` keeps the reference's parameter names (``qkv``, ``out_proj``).
The 3-axis rotation of head dims 0..5 is applied to q and k in registers while the
tiles are loaded by the fp32 attention kernel (no rotated copy is materialised); with
``dtype=torch.bfloat16`` the projections and the attention run on the tensor cores and the
rotation is applied in place to the projection output (``pcd_rope_bf16``) in between."""
import torch
import torch.nn as nn

from . import ops


class RotarySelfAttention(nn.Module):
    def __init__(self, dim, heads=8, dropout=0.0, dtype=torch.float32):
        super().__init__()
        if dtype not in (torch.float32, torch.bfloat16):
            raise ValueError("dtype must be torch.float32 (parity mode) or torch.bfloat16 (tensor-core mode)")
        self.compute_dtype = dtype
        if dim != heads * 64:
            raise ValueError("the attention kernels are built for head dim 64")
        if dropout != 0.0:
            raise NotImplementedError("inference path: dropout must be 0")
        self.heads = heads
        self.dim = dim
        self.scale = dim ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.out_proj = nn.Linear(dim, dim)

    @torch.no_grad()
    def forward(self, x, pos):
        B, N, D = x.shape
        dt = self.compute_dtype
        w = lambda lin: lin.weight.to(dt)
        qkv = ops.linear(x.to(dt).contiguous(), w(self.qkv), self.qkv.bias.float()).view(B, N, 3 * D)
        out = ops.rotary_attention(qkv, pos, self.heads)
        return ops.linear(out, w(self.out_proj), self.out_proj.bias.float(), out_dtype=torch.float32).view(B, N, D)
