"""TwoStreamDenoiser: the recurrent-interface (read / compute / write) point-cloud denoiser the reference
trains and evaluates (models/model.py:437-547, models/modules.py:17-244), behind the same constructor,
``forward(x, t, class_labels, viewpoints, partial_pcd, depth_maps, prev_latent) -> (x_denoised, latent)``
signature and ``state_dict`` keys, so a reference checkpoint loads unchanged and the model drops into
``PointCloudSampler`` (the sampler threads ``prev_latent`` through the guided denoiser, k_diffusion.py:190-203).

All arithmetic runs in libpcd_b200.so: projections through ``pcd_gemm_f32`` / ``pcd_gemm_bf16`` (bias, GELU and
residual epilogues), the head-dim-32 cross-attention through ``pcd_attention_hd32``, LayerNorm, timestep
embedding and the few plain stream additions through their kernels; PyTorch only owns the tensors (and the
embedding-table gathers / concatenations, which are memory movement).  The nn.Module tree below is a parameter
container with the reference's names -- it is never called.

Scope of this round: the backbone and the "class" / "view" condition encoders.  The partial-cloud and depth-map
encoders (nn.TransformerEncoder / Decoder stacks, model.py:262-434) are not built yet: constructing the model
with those modalities raises NotImplementedError instead of silently running something else.
"""
from typing import List, Optional

import torch
import torch.nn as nn

from . import ops
from ._lib import EPI_BIAS, EPI_BIAS_GELU

LN_EPS = 1e-5


class _Mlp(nn.Module):  # timm.models.vision_transformer.Mlp: fc1 -> GELU -> fc2
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.fc2 = nn.Linear(hidden, dim)


class _CrossAttention(nn.Module):  # models/modules.py:17-38
    def __init__(self, dim, kv_dim, qkv_bias):
        super().__init__()
        self.wq = nn.Linear(dim, dim, bias=qkv_bias)
        self.wk = nn.Linear(kv_dim, dim, bias=qkv_bias)
        self.wv = nn.Linear(kv_dim, dim, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)


class _ComputeBlock(nn.Module):  # models/modules.py:65-80
    def __init__(self, z_dim, mlp_ratio, qkv_bias):
        super().__init__()
        self.norm_z1 = nn.LayerNorm(z_dim)
        self.attn = _CrossAttention(z_dim, z_dim, qkv_bias)
        self.norm_z2 = nn.LayerNorm(z_dim)
        self.mlp = _Mlp(z_dim, int(z_dim * mlp_ratio))


class _ReadBlock(nn.Module):  # models/modules.py:82-99
    def __init__(self, z_dim, x_dim, mlp_ratio, qkv_bias):
        super().__init__()
        self.norm_x = nn.LayerNorm(x_dim)
        self.norm_z1 = nn.LayerNorm(z_dim)
        self.attn = _CrossAttention(z_dim, x_dim, qkv_bias)
        self.norm_z2 = nn.LayerNorm(z_dim)
        self.mlp = _Mlp(z_dim, int(z_dim * mlp_ratio))


class _WriteBlock(nn.Module):  # models/modules.py:101-120
    def __init__(self, z_dim, x_dim, mlp_ratio, qkv_bias):
        super().__init__()
        self.norm_z = nn.LayerNorm(z_dim)
        self.norm_x1 = nn.LayerNorm(x_dim)
        self.attn = _CrossAttention(x_dim, z_dim, qkv_bias)
        self.norm_x2 = nn.LayerNorm(x_dim)
        self.mlp = _Mlp(x_dim, int(x_dim * mlp_ratio))


class _RCWBlock(nn.Module):  # models/modules.py:122-146
    def __init__(self, z_dim, x_dim, num_compute_layers, mlp_ratio, qkv_bias):
        super().__init__()
        self.read = _ReadBlock(z_dim, x_dim, mlp_ratio, qkv_bias)
        self.write = _WriteBlock(z_dim, x_dim, mlp_ratio, qkv_bias)
        self.compute = nn.ModuleList([_ComputeBlock(z_dim, mlp_ratio, qkv_bias) for _ in range(num_compute_layers)])


class _Backbone(nn.Module):  # Denoiser_backbone, models/modules.py:148-196
    def __init__(self, input_channels, output_channels, num_z, num_x, z_dim, x_dim, num_blocks, num_compute_layers,
                 num_heads, mlp_ratio=4.0, qkv_bias=True):
        super().__init__()
        self.num_z, self.num_x, self.z_dim, self.x_dim, self.num_heads = num_z, num_x, z_dim, x_dim, num_heads
        self.input_proj = nn.Linear(input_channels, x_dim)
        self.ln_pre = nn.LayerNorm(x_dim)
        self.z_init = nn.Parameter(torch.zeros(1, num_z, z_dim))
        self.time_embed = _Mlp(z_dim, int(z_dim * mlp_ratio))
        self.latent_mlp = _Mlp(z_dim, int(z_dim * mlp_ratio))
        self.ln_latent = nn.LayerNorm(z_dim)
        self.blocks = nn.ModuleList([_RCWBlock(z_dim, x_dim, num_compute_layers, mlp_ratio, qkv_bias)
                                     for _ in range(num_blocks)])
        self.ln_post = nn.LayerNorm(x_dim)
        self.output_proj = nn.Linear(x_dim, output_channels)
        # reference initialisation (modules.py:179-196)
        nn.init.normal_(self.z_init, std=0.02)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.LayerNorm):
                nn.init.constant_(m.bias, 0)
                nn.init.constant_(m.weight, 1.0)
        nn.init.constant_(self.ln_latent.weight, 0)
        nn.init.constant_(self.ln_latent.bias, 0)


class _ClassEmbedding(nn.Module):  # models/model.py:217-232
    def __init__(self, num_classes, embed_dim):
        super().__init__()
        self.embedding = nn.Embedding(num_classes, embed_dim)
        self.norm = nn.LayerNorm(embed_dim)
        nn.init.normal_(self.embedding.weight, std=0.02)


class _ViewAngleEmbedding(nn.Module):  # models/model.py:235-259
    def __init__(self, input_dim, embed_dim):
        super().__init__()
        self.mlp = nn.Sequential(nn.Linear(input_dim, embed_dim // 2), nn.GELU(), nn.Linear(embed_dim // 2, embed_dim),
                                 nn.GELU(), nn.Linear(embed_dim, embed_dim), nn.LayerNorm(embed_dim))
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight)
                nn.init.constant_(m.bias, 0)


_TOKEN_TYPE = {"class": 0, "view": 1, "partial_pcd": 2, "depth": 3}


class TwoStreamDenoiser(nn.Module):
    def __init__(self, num_points: int = 1024, num_latents: int = 256, cond_drop_prob: float = 0.1,
                 input_channels: int = 3, output_channels: int = 3, latent_dim: int = 768, x_dim: int = 512,
                 num_blocks: int = 6, num_compute_layers: int = 4, num_classes: int = 16, num_heads: int = 8,
                 num_tokens_ppcd: int = 64, num_tokens_depth: int = 32,
                 active_modalities: List[str] = ("class", "view", "partial_pcd", "depth"),
                 device: Optional[torch.device] = None, dtype: torch.dtype = torch.float32):
        super().__init__()
        unsupported = [m for m in active_modalities if m in ("partial_pcd", "depth")]
        if unsupported:
            raise NotImplementedError(f"condition encoders {unsupported} are not built yet (models/model.py:262-434); "
                                      "construct the model with active_modalities=['class', 'view']")
        if latent_dim % num_heads or x_dim != latent_dim or latent_dim // num_heads != 32:
            # CrossAttention projects both streams to the query stream's width and splits it into heads
            # (modules.py:30-37); the attention kernel behind it is built for head dim 32 (the shipped config.yaml)
            raise ValueError("this build supports x_dim == latent_dim == 32 * num_heads (reference config.yaml:30-34)")
        if dtype not in (torch.float32, torch.bfloat16):
            raise ValueError("dtype must be torch.float32 (parity mode) or torch.bfloat16 (tensor-core projections)")
        if dtype == torch.bfloat16 and latent_dim % 128:
            raise ValueError("bf16 mode needs latent_dim % 128 == 0")
        self.num_points, self.latent_dim, self.cond_drop_prob = num_points, latent_dim, cond_drop_prob
        self.active_modalities = list(active_modalities)
        self.compute_dtype = dtype
        self.denoiser_backbone = _Backbone(input_channels, output_channels, num_latents, num_points, latent_dim, x_dim,
                                           num_blocks, num_compute_layers, num_heads)
        self.encoders = nn.ModuleDict()
        types = []
        for m in self.active_modalities:
            self.encoders[m] = (_ClassEmbedding(num_classes, latent_dim) if m == "class"
                                else _ViewAngleEmbedding(3, latent_dim))
            types.append(_TOKEN_TYPE[m])  # one token per class / view modality
        self.token_type_embeddings = nn.Embedding(4, latent_dim)
        nn.init.normal_(self.token_type_embeddings.weight, std=0.005)
        self.register_buffer("token_types_template", torch.tensor(types, dtype=torch.long))
        self._bf16 = {}
        if device is not None:
            self.to(device)
        self.eval()

    def cached_model_kwargs(self, batch_size, model_kwargs):  # model.py:486-487
        return model_kwargs

    # ---- kernels ---------------------------------------------------------------------------------
    def _w(self, lin: nn.Linear) -> torch.Tensor:
        """Weight in the compute dtype (bf16 copies are cached per parameter version)."""
        w = lin.weight
        if self.compute_dtype == torch.float32:
            return w.detach()
        key = (id(w), w._version, w.data_ptr())
        hit = self._bf16.get(id(w))
        if hit is None or hit[0] != key:
            hit = (key, w.detach().to(torch.bfloat16).contiguous())
            self._bf16[id(w)] = hit
        return hit[1]

    def _ln(self, x2: torch.Tensor, norm: nn.LayerNorm, act: bool = True) -> torch.Tensor:
        """LayerNorm of an fp32 [M, d] stream; act=True -> in the dtype the next projection consumes."""
        dt = self.compute_dtype if act else torch.float32
        return ops.layernorm(x2, norm.weight.detach(), norm.bias.detach(), eps=LN_EPS, out_dtype=dt)

    def _lin(self, a2: torch.Tensor, lin: nn.Linear, gelu: bool = False, residual: Optional[torch.Tensor] = None,
             out_fp32: bool = True) -> torch.Tensor:
        """Projection of [M, K] activations.  fp32 activations use the CUDA-core GEMM (parity mode and the tiny
        layers whose K is not a multiple of 8); bf16 activations use the tcgen05 GEMM."""
        bias = lin.bias.detach().float() if lin.bias is not None else None
        if a2.dtype == torch.float32:
            return ops.linear(a2, lin.weight.detach().float(), bias, epilogue=EPI_BIAS_GELU if gelu else EPI_BIAS,
                              residual=residual)
        return ops.linear(a2, self._w(lin), bias, epilogue=EPI_BIAS_GELU if gelu else EPI_BIAS, residual=residual,
                          out_dtype=torch.float32 if (out_fp32 or residual is not None) else torch.bfloat16)

    @staticmethod
    def _lin_k3(a2: torch.Tensor, lin: nn.Linear, gelu: bool = False) -> torch.Tensor:
        """fp32 projection of an [M, 3] input (xyz / view angles): K zero-padded to the GEMM's multiple of 4."""
        M, K = a2.shape
        Kp = (K + 3) // 4 * 4
        a = torch.zeros(M, Kp, device=a2.device)
        a[:, :K].copy_(a2)
        w = torch.zeros(lin.weight.shape[0], Kp, device=a2.device)
        w[:, :K].copy_(lin.weight.detach())
        return ops.linear(a, w, lin.bias.detach().float(), epilogue=EPI_BIAS_GELU if gelu else EPI_BIAS)

    def _act(self, a2: torch.Tensor) -> torch.Tensor:
        """fp32 [M, d] -> the dtype the next projection consumes."""
        return a2 if self.compute_dtype == torch.float32 else ops.cast_rowstats(a2)[0]

    def _mlp(self, stream2: torch.Tensor, norm: nn.LayerNorm, mlp: _Mlp) -> torch.Tensor:
        """stream + fc2(gelu(fc1(LN(stream)))) (residual add in the fc2 epilogue)."""
        hid = self._lin(self._ln(stream2, norm), mlp.fc1, gelu=True, out_fp32=False)
        return self._lin(hid, mlp.fc2, residual=stream2)

    def _attend(self, q_stream: torch.Tensor, kv_stream: torch.Tensor, B: int, attn: _CrossAttention,
                residual: torch.Tensor) -> torch.Tensor:
        """residual + proj(softmax(q k^T / sqrt(32)) v) with q from q_stream, k / v from kv_stream (modules.py:40-63)."""
        heads, d = self.denoiser_backbone.num_heads, residual.shape[1]
        q = self._lin(q_stream, attn.wq).view(B, -1, d)
        k = self._lin(kv_stream, attn.wk).view(B, -1, d)
        v = self._lin(kv_stream, attn.wv).view(B, -1, d)
        a = ops.attention_hd32(q, k, v, heads).view(-1, d)
        return self._lin(self._act(a), attn.proj, residual=residual)

    # ---- conditioning (model.py:489-538, eval branch) ------------------------------------------------
    def _cond_tokens(self, B, dev, class_labels, viewpoints) -> torch.Tensor:
        d = self.latent_dim
        cond = torch.zeros(B, len(self.active_modalities), d, device=dev)
        te = self.token_type_embeddings.weight.detach()[self.token_types_template]  # gather: [n_cond, d]
        for i, m in enumerate(self.active_modalities):
            value = class_labels if m == "class" else viewpoints
            if value is None or bool(torch.all(value == 0)):
                continue  # zero token, masked type embedding
            if m == "class":
                enc = self.encoders["class"]
                rows = enc.embedding.weight.detach()[value.to(dev).long()].float().contiguous()  # table gather
                tok = ops.layernorm(rows, enc.norm.weight.detach(), enc.norm.bias.detach(), eps=LN_EPS)
            else:
                mlp = self.encoders["view"].mlp
                h = value.to(dev).float().contiguous()
                h = self._lin_k3(h, mlp[0], gelu=True)
                h = ops.linear(h, mlp[2].weight.detach(), mlp[2].bias.detach(), epilogue=EPI_BIAS_GELU)
                h = ops.linear(h, mlp[4].weight.detach(), mlp[4].bias.detach())
                tok = ops.layernorm(h, mlp[5].weight.detach(), mlp[5].bias.detach(), eps=LN_EPS)
            cond[:, i].copy_(ops.add(tok, te[i].expand(B, d).contiguous()))
        return cond

    # ---- forward ---------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, x, t, class_labels=None, viewpoints=None, partial_pcd=None, depth_maps=None, prev_latent=None):
        assert x.shape[-1] == self.num_points, \
            f"Input point cloud must have {self.num_points} points, got {x.shape[-1]} points."
        assert partial_pcd is None and depth_maps is None, "partial_pcd / depth encoders are not built yet"
        if self.training:
            raise NotImplementedError("inference path only (the reference's training branch draws dropout masks)")
        bb = self.denoiser_backbone
        B, dev = x.shape[0], x.device
        d, n_cond = self.latent_dim, len(self.active_modalities)
        n_lat = bb.num_z + n_cond + 1
        cond = self._cond_tokens(B, dev, class_labels, viewpoints)

        # timestep token (modules.py:220): Mlp(timestep_embedding(t)) -- two tiny fp32 projections
        te = ops.timestep_embedding(t, d)
        te = ops.linear(te, bb.time_embed.fc1.weight.detach(), bb.time_embed.fc1.bias.detach(), epilogue=EPI_BIAS_GELU)
        te = ops.linear(te, bb.time_embed.fc2.weight.detach(), bb.time_embed.fc2.bias.detach())

        # x stream: input_proj + ln_pre (modules.py:223-224); K = 3 -> CUDA-core GEMM
        pts = x.float().permute(0, 2, 1).contiguous().view(B * self.num_points, -1)
        xs = self._lin_k3(pts, bb.input_proj)
        xs = self._ln(xs, bb.ln_pre, act=False)

        # latent stream with self-conditioning (modules.py:226-229)
        z = torch.empty(B, n_lat, d, device=dev)
        z[:, :bb.num_z].copy_(bb.z_init.detach().expand(B, -1, -1))
        z[:, bb.num_z:bb.num_z + n_cond].copy_(cond)
        z[:, -1].copy_(te)
        z = z.view(B * n_lat, d)
        if prev_latent is None:
            prev = torch.zeros(B * n_lat, d, device=dev)
        else:
            assert prev_latent.shape == (B, n_lat, d)
            prev = prev_latent.float().contiguous().view(B * n_lat, d)
        hid = self._lin(self._act(prev), bb.latent_mlp.fc1, gelu=True, out_fp32=False)
        prev = self._lin(hid, bb.latent_mlp.fc2, residual=prev)
        z = ops.add(z, self._ln(prev, bb.ln_latent, act=False))

        for blk in bb.blocks:
            z = self._attend(self._ln(z, blk.read.norm_z1), self._ln(xs, blk.read.norm_x), B, blk.read.attn, z)
            z = self._mlp(z, blk.read.norm_z2, blk.read.mlp)
            for cb in blk.compute:
                zn = self._ln(z, cb.norm_z1)
                z = self._attend(zn, zn, B, cb.attn, z)
                z = self._mlp(z, cb.norm_z2, cb.mlp)
            xs = self._attend(self._ln(xs, blk.write.norm_x1), self._ln(z, blk.write.norm_z), B, blk.write.attn, xs)
            xs = self._mlp(xs, blk.write.norm_x2, blk.write.mlp)

        out = ops.linear(self._ln(xs, bb.ln_post, act=False), bb.output_proj.weight.detach(), bb.output_proj.bias.detach())
        return out.view(B, self.num_points, -1).permute(0, 2, 1).contiguous(), z.view(B, n_lat, d)
