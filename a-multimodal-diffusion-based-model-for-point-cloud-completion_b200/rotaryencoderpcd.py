"""Rotary point encoding (reference models/rotaryencoderpcd.py:6-27, 58-84).

``RotarySelfAttention`` keeps the reference's parameter names (``qkv``, ``out_proj``).
The 3-axis rotation of head dims 0..5 is applied to q and k in registers while the
tiles are loaded by the fp32 attention kernel (no rotated copy is materialised)."""
import torch
import torch.nn as nn

from . import ops


class RotarySelfAttention(nn.Module):
    def __init__(self, dim, heads=8, dropout=0.0):
        super().__init__()
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
        qkv = ops.linear(x.float().contiguous(), self.qkv.weight, self.qkv.bias).view(B, N, 3 * D)
        out = ops.rotary_attention(qkv, pos, self.heads)
        return ops.linear(out, self.out_proj.weight, self.out_proj.bias).view(B, N, D)
