"""Point-cloud diffusion transformer family on the B200 kernels.

Same class names, constructor keywords, ``state_dict`` keys and
``model(x, t, **kwargs)`` / ``cached_model_kwargs`` signatures as the reference's
``models/transformer.py``; the forward pass is one call into the C ABI
(``pcd_model_forward``), which runs the whole denoiser with hand-written sm_100a
kernels.  Parameters are kept as fp32 ``nn.Parameter``s (so a reference checkpoint
loads with ``load_state_dict``) and packed for the kernels on first use.

Precision: ``dtype=torch.bfloat16`` -> tcgen05 bf16 GEMMs + tcgen05 flash attention
(fp32 residual stream, LayerNorm, softmax and accumulation);
``dtype=torch.float32`` -> fp32 CUDA-core kernels (1e-4 parity mode).
"""
import ctypes as C
import math
from typing import Any, Dict, Iterable, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from . import _lib, ops
from ._lib import BlockWeights, ModelDesc, PCD_BF16, PCD_F32, check, ptr, require_cuda, stream_ptr
from .pretrained_clip import FrozenImageCLIP, ImageCLIP

LN_EPS = 1e-5


def init_linear(l, stddev):
    nn.init.normal_(l.weight, std=stddev)
    if l.bias is not None:
        nn.init.constant_(l.bias, 0.0)


def fold_layernorm_into_linear(weight: torch.Tensor, bias: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor):
    """LayerNorm(x; gamma, beta) followed by Linear(weight, bias) (reference transformer.py:108-114:
    ``attn(ln_1(x))``, ``mlp(ln_2(x))``) as ONE projection of the un-normalised row:

        LN(x) W^T + b = rstd * (x W'^T - mu * colsum) + const,

    W' = bf16(gamma o W) (what the tensor cores multiply by), colsum[n] = sum_k W'[n, k] (of the ROUNDED
    folded weights, so that the mean correction cancels exactly) and const[n] = beta . W[n, :] + b[n].
    mu / rstd are the statistics of the row the GEMM actually reads (its bf16 copy)."""
    w32, g, be = weight.float(), gamma.float(), beta.float()
    wf = (w32 * g[None, :]).to(torch.bfloat16).contiguous()
    colsum = wf.float().sum(dim=1).contiguous()
    const = (w32.double() @ be.double() + bias.double()).float().contiguous()
    return wf, colsum, const


def _compute_dtype(t: torch.Tensor, precision: int):
    return t.to(torch.bfloat16) if precision == PCD_BF16 else t.float()


# ---------------------------------------------------------------------------
# Parameter containers with the reference's module tree (state_dict key parity).
# Their forwards run on the kernels too, so blocks can be used / tested alone.
# ---------------------------------------------------------------------------
class MLP(nn.Module):
    """reference models/transformer.py:51-62"""

    def __init__(self, *, device, dtype, width: int, init_scale: float):
        super().__init__()
        self.width = width
        self.c_fc = nn.Linear(width, width * 4, device=device, dtype=torch.float32)
        self.c_proj = nn.Linear(width * 4, width, device=device, dtype=torch.float32)
        init_linear(self.c_fc, init_scale)
        init_linear(self.c_proj, init_scale)

    def forward(self, x):  # fp32 CUDA-core path (tiny per-sequence MLPs)
        h = ops.linear(x.float(), self.c_fc.weight, self.c_fc.bias, epilogue=_lib.EPI_BIAS_GELU)
        return ops.linear(h, self.c_proj.weight, self.c_proj.bias)


class MultiheadAttention(nn.Module):
    """reference models/transformer.py:23-48"""

    def __init__(self, *, device, dtype, n_ctx: int, width: int, heads: int, init_scale: float):
        super().__init__()
        self.n_ctx, self.width, self.heads = n_ctx, width, heads
        self.c_qkv = nn.Linear(width, width * 3, device=device, dtype=torch.float32)
        self.c_proj = nn.Linear(width, width, device=device, dtype=torch.float32)
        init_linear(self.c_qkv, init_scale)
        init_linear(self.c_proj, init_scale)


class ResidualAttentionBlock(nn.Module):
    """reference models/transformer.py:87-115"""

    def __init__(self, *, device, dtype, n_ctx: int, width: int, heads: int, init_scale: float = 1.0):
        super().__init__()
        self.compute_dtype = dtype
        self.attn = MultiheadAttention(device=device, dtype=dtype, n_ctx=n_ctx, width=width, heads=heads,
                                       init_scale=init_scale)
        self.ln_1 = nn.LayerNorm(width, device=device, dtype=torch.float32)
        self.mlp = MLP(device=device, dtype=dtype, width=width, init_scale=init_scale)
        self.ln_2 = nn.LayerNorm(width, device=device, dtype=torch.float32)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """x fp32 [B, L, width] -> fp32, composed from the individual kernels."""
        cd = self.compute_dtype
        w = lambda lin: lin.weight.to(cd)
        a = ops.layernorm(x, self.ln_1.weight, self.ln_1.bias, LN_EPS, out_dtype=cd)
        qkv = ops.linear(a, w(self.attn.c_qkv), self.attn.c_qkv.bias)
        att = ops.self_attention(qkv.view(x.shape[0], x.shape[1], -1), self.attn.heads)
        x = ops.linear(att, w(self.attn.c_proj), self.attn.c_proj.bias, residual=x.contiguous(),
                       out_dtype=torch.float32).view_as(x)
        m = ops.layernorm(x, self.ln_2.weight, self.ln_2.bias, LN_EPS, out_dtype=cd)
        hdn = ops.linear(m, w(self.mlp.c_fc), self.mlp.c_fc.bias, epilogue=_lib.EPI_BIAS_GELU)
        return ops.linear(hdn, w(self.mlp.c_proj), self.mlp.c_proj.bias, residual=x.contiguous(),
                          out_dtype=torch.float32).view_as(x)


class Transformer(nn.Module):
    """reference models/transformer.py:118-152"""

    def __init__(self, *, device, dtype, n_ctx: int, width: int, layers: int, heads: int,
                 init_scale: float = 0.25):
        super().__init__()
        self.n_ctx, self.width, self.layers = n_ctx, width, layers
        init_scale = init_scale * math.sqrt(1.0 / width)
        self.resblocks = nn.ModuleList([
            ResidualAttentionBlock(device=device, dtype=dtype, n_ctx=n_ctx, width=width, heads=heads,
                                   init_scale=init_scale) for _ in range(layers)])

    def forward(self, x: torch.Tensor):
        for block in self.resblocks:
            x = block(x)
        return x


# ---------------------------------------------------------------------------
class PointDiffusionTransformer(nn.Module):
    """reference models/transformer.py:155-226"""

    pcd_native = True

    def __init__(self, *, device: torch.device, dtype: torch.dtype = torch.bfloat16, input_channels: int = 3,
                 output_channels: int = 3, n_ctx: int = 1024, width: int = 512, layers: int = 12,
                 heads: int = 8, init_scale: float = 0.25, time_token_cond: bool = False):
        super().__init__()
        if dtype not in (torch.float32, torch.bfloat16):
            raise ValueError("dtype must be torch.float32 (parity mode) or torch.bfloat16 (tensor-core mode)")
        if width != heads * 64:
            raise ValueError("the attention kernels are built for head dim 64 (every registered config)")
        self.compute_dtype = dtype
        self.input_channels = input_channels
        self.output_channels = output_channels
        self.n_ctx = n_ctx
        self.time_token_cond = time_token_cond
        self.time_embed = MLP(device=device, dtype=dtype, width=width,
                              init_scale=init_scale * math.sqrt(1.0 / width))
        self.ln_pre = nn.LayerNorm(width, device=device, dtype=torch.float32)
        self.backbone = Transformer(device=device, dtype=dtype, n_ctx=n_ctx + int(time_token_cond),
                                    width=width, layers=layers, heads=heads, init_scale=init_scale)
        self.ln_post = nn.LayerNorm(width, device=device, dtype=torch.float32)
        self.input_proj = nn.Linear(input_channels, width, device=device, dtype=torch.float32)
        self.output_proj = nn.Linear(width, output_channels, device=device, dtype=torch.float32)
        with torch.no_grad():
            self.output_proj.weight.zero_()
            self.output_proj.bias.zero_()
        self.heads = heads
        self._handle = None
        self._packed_key = None
        self._keep = []          # packed tensors referenced by the C handle
        self._ws: Dict[int, torch.Tensor] = {}
        self._prefix: Dict[int, torch.Tensor] = {}
        self._cond_key: Dict[int, Any] = {}
        self._cond_refs: Dict[int, Any] = {}
        self._addc: Dict[int, Optional[torch.Tensor]] = {}
        # per-handle options (pcd_model_desc.flags): LayerNorm folded into the projections (bf16 mode) and the
        # tensor-core attention kernel of the forward (_lib.ATTN_*, 0 = default)
        self.fold_layernorm = True
        self.attention_variant = 0
        self._pack_version = 0   # bumped whenever the packed weights move (CUDA graphs captured before are stale)
        self._time_tok: Dict[float, torch.Tensor] = {}
        self._tcond: Dict[int, torch.Tensor] = {}
        self._out: Dict[Tuple[int, int], torch.Tensor] = {}

    # ---- token layout (subclasses override) ----
    def _prefix_layout(self) -> Tuple[int, int]:
        """(number of prefix tokens, index of the time token or -1)."""
        return (1, 0) if self.time_token_cond else (0, -1)

    def _fill_cond(self, seqs: int, kw: Dict[str, Any], prefix: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
        """Write the step-invariant conditioning tokens into ``prefix`` and return the
        non-token conditioning vector [seqs, width] (or None)."""
        return None

    # ---- packing ----
    def _version_key(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters()) + tuple(
            (b.data_ptr(), b._version) for b in self.buffers()) + (bool(self.fold_layernorm), int(self.attention_variant))

    def _ensure_handle(self):
        key = self._version_key()
        if self._handle is not None and key == self._packed_key:
            return
        require_cuda(self.ln_pre.weight)
        lib = _lib.load()
        check(lib.pcd_check_device(), "device check")
        self._destroy_handle()
        prec = PCD_BF16 if self.compute_dtype == torch.bfloat16 else PCD_F32
        keep: List[torch.Tensor] = []

        def f32(p):
            t = p.detach().float().contiguous()
            keep.append(t)
            return ptr(t)

        def mat(p):
            t = _compute_dtype(p.detach(), prec).contiguous()
            keep.append(t)
            return ptr(t)

        blocks = (BlockWeights * len(self.backbone.resblocks))()
        for i, blk in enumerate(self.backbone.resblocks):
            b = blocks[i]
            b.ln1_g, b.ln1_b = f32(blk.ln_1.weight), f32(blk.ln_1.bias)
            b.ln2_g, b.ln2_b = f32(blk.ln_2.weight), f32(blk.ln_2.bias)
            b.w_qkv, b.b_qkv = mat(blk.attn.c_qkv.weight), f32(blk.attn.c_qkv.bias)
            b.w_proj, b.b_proj = mat(blk.attn.c_proj.weight), f32(blk.attn.c_proj.bias)
            b.w_fc, b.b_fc = mat(blk.mlp.c_fc.weight), f32(blk.mlp.c_fc.bias)
            b.w_fc2, b.b_fc2 = mat(blk.mlp.c_proj.weight), f32(blk.mlp.c_proj.bias)
            if prec == PCD_BF16 and self.backbone.width % 256 == 0:
                # LayerNorm folded into the projection that follows it (ln_1 -> c_qkv, ln_2 -> c_fc,
                # transformer.py:108-114):  LN(x) W^T + b = rstd (x (gamma o W)^T - mu s) + c  with
                # s_n = sum_k (gamma o W)[n, k] of the bf16-ROUNDED folded weights (what the tensor
                # cores multiply by) and c_n = beta . W[n, :] + b_n.  Step-invariant: done once here.
                def fold(lin, ln):
                    wf, colsum, const = fold_layernorm_into_linear(lin.weight.detach(), lin.bias.detach(),
                                                                   ln.weight.detach(), ln.bias.detach())
                    keep.extend((wf, colsum, const))
                    return ptr(wf), ptr(colsum), ptr(const)
                b.w_qkv_ln, b.qkv_colsum, b.qkv_const = fold(blk.attn.c_qkv, blk.ln_1)
                b.w_fc_ln, b.fc_colsum, b.fc_const = fold(blk.mlp.c_fc, blk.ln_2)
        n_prefix, time_slot = self._prefix_layout()
        width = self.backbone.width
        freqs = ops.timestep_freqs(width, device=self.ln_pre.weight.device)
        keep.append(freqs)
        d = ModelDesc()
        d.precision, d.width, d.heads, d.layers = prec, width, self.heads, len(self.backbone.resblocks)
        d.c_in, d.c_out, d.n_points = self.input_channels, self.output_channels, self.n_ctx
        d.n_prefix, d.time_slot, d.ln_eps = n_prefix, time_slot, LN_EPS
        d.flags = ((0 if self.fold_layernorm else _lib.MODEL_SEPARATE_LAYERNORM)
                   | (int(self.attention_variant) << _lib.MODEL_ATTN_VARIANT_SHIFT))
        d.time_fc_w, d.time_fc_b = f32(self.time_embed.c_fc.weight), f32(self.time_embed.c_fc.bias)
        d.time_proj_w, d.time_proj_b = f32(self.time_embed.c_proj.weight), f32(self.time_embed.c_proj.bias)
        d.freqs = ptr(freqs)
        d.ln_pre_g, d.ln_pre_b = f32(self.ln_pre.weight), f32(self.ln_pre.bias)
        d.ln_post_g, d.ln_post_b = f32(self.ln_post.weight), f32(self.ln_post.bias)
        d.in_w, d.in_b = f32(self.input_proj.weight), f32(self.input_proj.bias)
        d.out_w, d.out_b = f32(self.output_proj.weight), f32(self.output_proj.bias)
        d.blocks = blocks
        handle = C.c_void_p()
        check(lib.pcd_model_create(C.byref(d), C.byref(handle)), "model_create")
        self._handle, self._packed_key, self._keep = handle, key, keep
        self._pack_version += 1
        self._n_prefix, self._time_slot = n_prefix, time_slot
        self._cond_key.clear()
        self._time_tok.clear()

    def graph_key(self):
        """Changes whenever a CUDA graph captured over this model would read freed / repacked memory."""
        self._ensure_handle()
        return self._pack_version

    def _destroy_handle(self):
        if getattr(self, "_handle", None) is not None:
            _lib.load().pcd_model_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self._destroy_handle()
        except Exception:
            pass

    # ---- execution ----
    def _get_buffers(self, seqs: int):
        dev = self.ln_pre.weight.device
        if seqs not in self._ws:
            nbytes = _lib.load().pcd_model_workspace_bytes(self._handle, seqs)
            self._ws[seqs] = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            self._prefix[seqs] = (torch.zeros(seqs, self._n_prefix, self.backbone.width, device=dev)
                                  if self._n_prefix else None)
        return self._ws[seqs], self._prefix[seqs]

    def _run(self, x: torch.Tensor, t: Optional[torch.Tensor], kw: Dict[str, Any], seqs: int, out_channels: int,
             out: Optional[torch.Tensor] = None, add_cond_override: Optional[torch.Tensor] = None) -> torch.Tensor:
        require_cuda(x, t)
        assert x.dim() == 3 and x.shape[1] == self.input_channels and x.shape[2] == self.n_ctx, \
            f"expected x of shape [B, {self.input_channels}, {self.n_ctx}], got {tuple(x.shape)}"
        self._ensure_handle()
        ws, prefix = self._get_buffers(seqs)
        add_cond = self.prepare_cond(seqs, kw)
        if add_cond_override is not None:
            add_cond = add_cond_override
        x = x.float().contiguous()
        if t is not None:
            t = t.to(torch.float32).contiguous()
            assert t.shape == (seqs,)
        if out is None:
            out = torch.empty(seqs, out_channels, self.n_ctx, device=x.device, dtype=torch.float32)
        check(_lib.load().pcd_model_forward(self._handle, ptr(x), x.shape[0], ptr(t), ptr(prefix), ptr(add_cond),
                                            ptr(out), out_channels, ptr(ws), ws.numel(), seqs, stream_ptr()),
              "model_forward")
        return out

    def prepare_cond(self, seqs: int, kw: Dict[str, Any]) -> Optional[torch.Tensor]:
        """Compute the step-invariant conditioning tokens for ``kw`` once (the reference
        recomputes them on every forward, e.g. transformer.py:351-352,404-409) and keep them
        in the persistent prefix buffer.  Cached on tensor identity + version; the keyed
        tensors are kept alive so their addresses cannot be recycled under the cache."""
        self._ensure_handle()
        _, prefix = self._get_buffers(seqs)
        tensors = {k: v for k, v in kw.items() if torch.is_tensor(v)}
        ckey = tuple(sorted((k, v.data_ptr(), v._version, tuple(v.shape)) for k, v in tensors.items()))
        if self._cond_key.get(seqs) != ckey or len(tensors) != len(kw):
            addc = self._fill_cond(seqs, kw, prefix)
            if addc is not None:
                # persistent per-seqs buffer: a captured CUDA graph holds this address, so new conditioning is
                # copied INTO it (like the prefix tokens), never swapped for a fresh tensor
                buf = self._addc.get(seqs)
                if buf is None or buf.shape != addc.shape:
                    buf = torch.empty_like(addc)
                buf.copy_(addc)
                addc = buf
            self._addc[seqs] = addc
            self._cond_key[seqs] = ckey
            self._cond_refs[seqs] = list(tensors.values())
        return self._addc.get(seqs)

    def _model_forward(self, x: torch.Tensor, t: torch.Tensor, **kw) -> torch.Tensor:
        assert x.shape[-1] == self.n_ctx
        return self._run(x, t, kw, x.shape[0], self.output_channels)

    def forward(self, x: torch.Tensor, t: torch.Tensor):
        """:param x: [N x C x T]  :param t: [N]  :return: [N x C' x T] (fp32)."""
        return self._model_forward(x, t)

    @torch.no_grad()
    def forward_cfg(self, x: torch.Tensor, t: int, model_kwargs: Dict[str, Any], doubled: bool,
                    out_channels: Optional[int] = None) -> torch.Tensor:
        """Sampler fast path: one forward over ``seqs`` = B (or 2B with classifier-free
        guidance: kwargs rows [:B] conditional, [B:] unconditional, sampler.py:133-136)
        sequences that share the same B inputs ``x``; ``t`` is the common integer timestep.
        Returns a persistent [seqs, out_channels, N] buffer (overwritten by the next call)."""
        B = x.shape[0]
        seqs = 2 * B if doubled else B
        oc = out_channels or self.output_channels
        okey = (seqs, oc)
        if okey not in self._out:
            self._out[okey] = torch.empty(seqs, oc, self.n_ctx, device=x.device, dtype=torch.float32)
        kw = {k: v for k, v in (model_kwargs or {}).items() if k != "prev_latent"}
        # every sequence shares the timestep: the time token is computed once per distinct t
        # (time_embed MLP on ONE row) and broadcast into the prefix slot
        self._ensure_handle()
        tok = self.time_token(t)
        _, prefix = self._get_buffers(seqs)
        if self._time_slot >= 0:
            prefix[:, self._time_slot].copy_(tok.expand(seqs, -1))
            return self._run(x, None, kw, seqs, oc, out=self._out[okey])
        addc = self.prepare_cond(seqs, kw)
        tcond = self._tcond.setdefault(seqs, torch.empty(seqs, tok.shape[-1], device=x.device))
        if addc is None:
            tcond.copy_(tok.expand(seqs, -1))
        else:
            torch.add(addc, tok, out=tcond)
        return self._run(x, None, kw, seqs, oc, out=self._out[okey], add_cond_override=tcond)

    @torch.no_grad()
    def time_token(self, t) -> torch.Tensor:
        """time_embed(timestep_embedding(t)) for ONE timestep value -> [1, width], cached."""
        key = float(t)
        tok = self._time_tok.get(key)
        if tok is None:
            dev = self.ln_pre.weight.device
            tt = torch.full((1,), key, device=dev, dtype=torch.float32)
            tok = self.time_embed(ops.timestep_embedding(tt, self.backbone.width))
            self._time_tok[key] = tok
        return tok

    @torch.no_grad()
    def prepare_time_tokens(self, timesteps) -> None:
        """Batch-compute the time tokens of a whole sampling schedule (distinct values only)."""
        self._ensure_handle()
        todo = sorted({float(t) for t in timesteps} - set(self._time_tok))
        if not todo:
            return
        dev = self.ln_pre.weight.device
        tt = torch.tensor(todo, device=dev, dtype=torch.float32)
        toks = self.time_embed(ops.timestep_embedding(tt, self.backbone.width))
        for i, k in enumerate(todo):
            self._time_tok[k] = toks[i:i + 1].clone()

    # helpers for subclasses -------------------------------------------------
    def _embed_grid(self, grid: torch.Tensor) -> torch.Tensor:
        """[S, 1024, 256] -> LN(1024) -> Linear(1024 -> width): [S, 256, width] (transformer.py:351-352)."""
        ln, lin = self.clip_embed[0], self.clip_embed[1]
        g = grid.float().permute(0, 2, 1).contiguous()
        g = ops.layernorm(g, ln.weight, ln.bias, LN_EPS)
        return ops.linear(g, lin.weight, lin.bias)

    def _embed_low_res(self, x: torch.Tensor) -> torch.Tensor:
        """[S, C, P] -> channel scale/bias -> Linear(C -> width): [S, P, width] (transformer.py:404-409)."""
        x = x.float()
        if self.channel_scales is not None:
            x = x * self.channel_scales[None, :, None]
        if self.channel_biases is not None:
            x = x + self.channel_biases[None, :, None]
        xp = x.permute(0, 2, 1)
        cin = xp.shape[-1]
        pad = (-cin) % 4  # the CUDA-core GEMM wants K % 4 == 0: zero-pad the channel axis
        xp = torch.nn.functional.pad(xp, (0, pad)).contiguous()
        w = torch.nn.functional.pad(self.cond_point_proj.weight.detach().float(), (0, pad)).contiguous()
        return ops.linear(xp, w, self.cond_point_proj.bias)


class CLIPImagePointDiffusionTransformer(PointDiffusionTransformer):
    """reference models/transformer.py:229-287 (configs base40M-imagevec / -textvec)."""

    def __init__(self, *, device, dtype=torch.bfloat16, n_ctx: int = 1024, token_cond: bool = False,
                 cond_drop_prob: float = 0.0, frozen_clip: bool = True, cache_dir: Optional[str] = None,
                 **kwargs):
        super().__init__(device=device, dtype=dtype, n_ctx=n_ctx + int(token_cond), **kwargs)
        self.n_ctx = n_ctx
        self.token_cond = token_cond
        self.clip = (FrozenImageCLIP if frozen_clip else ImageCLIP)(device, cache_dir=cache_dir)
        self.clip_embed = nn.Linear(self.clip.feature_dim, self.backbone.width, device=device, dtype=torch.float32)
        self.cond_drop_prob = cond_drop_prob

    def _prefix_layout(self):
        # token order = order of the cond list: [clip, t] (transformer.py:286, :212-218)
        n = int(self.token_cond) + int(self.time_token_cond)
        return n, (int(self.token_cond) if self.time_token_cond else -1)

    def cached_model_kwargs(self, batch_size: int, model_kwargs: Dict[str, Any]) -> Dict[str, Any]:
        with torch.no_grad():
            return dict(embeddings=self.clip(batch_size, **model_kwargs))

    def _fill_cond(self, seqs, kw, prefix):
        clip_out = self.clip(batch_size=seqs, images=kw.get("images"), texts=kw.get("texts"),
                             embeddings=kw.get("embeddings"))
        assert clip_out.dim() == 2 and clip_out.shape[0] == seqs
        clip_out = math.sqrt(clip_out.shape[1]) * clip_out.float()
        emb = ops.linear(clip_out.contiguous(), self.clip_embed.weight, self.clip_embed.bias)
        if self.token_cond:
            prefix[:, 0].copy_(emb)
            return None
        return emb

    def forward(self, x, t, images=None, texts=None, embeddings=None):
        kw = {k: v for k, v in dict(images=images, texts=texts, embeddings=embeddings).items() if v is not None}
        return self._model_forward(x, t, **kw)


class CLIPImageGridPointDiffusionTransformer(PointDiffusionTransformer):
    """reference models/transformer.py:290-355 (configs base40M / base300M / base1B)."""

    def __init__(self, *, device, dtype=torch.bfloat16, n_ctx: int = 1024, cond_drop_prob: float = 0.0,
                 frozen_clip: bool = True, cache_dir: Optional[str] = None, **kwargs):
        clip = (FrozenImageCLIP if frozen_clip else ImageCLIP)(device, cache_dir=cache_dir)
        super().__init__(device=device, dtype=dtype, n_ctx=n_ctx + clip.grid_size ** 2, **kwargs)
        self.n_ctx = n_ctx
        self.clip = clip
        self.clip_embed = nn.Sequential(
            nn.LayerNorm(normalized_shape=(self.clip.grid_feature_dim,), device=device, dtype=torch.float32),
            nn.Linear(self.clip.grid_feature_dim, self.backbone.width, device=device, dtype=torch.float32))
        self.cond_drop_prob = cond_drop_prob

    def _prefix_layout(self):
        g = self.clip.grid_size ** 2
        return int(self.time_token_cond) + g, (0 if self.time_token_cond else -1)

    def cached_model_kwargs(self, batch_size: int, model_kwargs: Dict[str, Any]) -> Dict[str, Any]:
        # the reference requires `images` here (transformer.py:317-320); precomputed grid
        # embeddings are accepted as an extension since CLIP weights are not bundled.
        if "images" not in model_kwargs and "embeddings" in model_kwargs:
            return dict(embeddings=model_kwargs["embeddings"])
        with torch.no_grad():
            return dict(embeddings=self.clip.embed_images_grid(model_kwargs["images"]))

    def _fill_cond(self, seqs, kw, prefix):
        images, embeddings = kw.get("images"), kw.get("embeddings")
        assert images is not None or embeddings is not None, "must specify images or embeddings"
        assert images is None or embeddings is None, "cannot specify both images and embeddings"
        clip_out = self.clip.embed_images_grid(images) if images is not None else embeddings
        s0 = int(self.time_token_cond)
        prefix[:, s0:s0 + self.clip.grid_size ** 2].copy_(self._embed_grid(clip_out))
        return None

    def forward(self, x, t, images=None, embeddings=None):
        kw = {k: v for k, v in dict(images=images, embeddings=embeddings).items() if v is not None}
        return self._model_forward(x, t, **kw)


class UpsamplePointDiffusionTransformer(PointDiffusionTransformer):
    """reference models/transformer.py:358-409."""

    def __init__(self, *, device, dtype=torch.bfloat16, cond_input_channels: Optional[int] = None,
                 cond_ctx: int = 1024, n_ctx: int = 4096 - 1024,
                 channel_scales: Optional[Sequence[float]] = None,
                 channel_biases: Optional[Sequence[float]] = None, **kwargs):
        super().__init__(device=device, dtype=dtype, n_ctx=n_ctx + cond_ctx, **kwargs)
        self.n_ctx = n_ctx
        self.cond_ctx = cond_ctx
        self.cond_input_channels = cond_input_channels or self.input_channels
        self.cond_point_proj = nn.Linear(self.cond_input_channels, self.backbone.width, device=device,
                                         dtype=torch.float32)
        self.register_buffer("channel_scales", torch.tensor(channel_scales, dtype=torch.float32, device=device)
                             if channel_scales is not None else None)
        self.register_buffer("channel_biases", torch.tensor(channel_biases, dtype=torch.float32, device=device)
                             if channel_biases is not None else None)

    def _prefix_layout(self):
        return int(self.time_token_cond) + self.cond_ctx, (0 if self.time_token_cond else -1)

    def _fill_cond(self, seqs, kw, prefix):
        low = self._embed_low_res(kw["low_res"])
        assert low.shape[1] == self.cond_ctx, f"low_res must have {self.cond_ctx} points"
        s0 = int(self.time_token_cond)
        prefix[:, s0:s0 + self.cond_ctx].copy_(low)
        return None

    def forward(self, x, t, *, low_res):
        return self._model_forward(x, t, low_res=low_res)


class CLIPImageGridUpsamplePointDiffusionTransformer(UpsamplePointDiffusionTransformer):
    """reference models/transformer.py:412-494 (config "upsample")."""

    def __init__(self, *, device, dtype=torch.bfloat16, n_ctx: int = 4096 - 1024, cond_drop_prob: float = 0.0,
                 frozen_clip: bool = True, cache_dir: Optional[str] = None, **kwargs):
        clip = (FrozenImageCLIP if frozen_clip else ImageCLIP)(device, cache_dir=cache_dir)
        super().__init__(device=device, dtype=dtype, n_ctx=n_ctx + clip.grid_size ** 2, **kwargs)
        self.n_ctx = n_ctx
        self.clip = clip
        self.clip_embed = nn.Sequential(
            nn.LayerNorm(normalized_shape=(self.clip.grid_feature_dim,), device=device, dtype=torch.float32),
            nn.Linear(self.clip.grid_feature_dim, self.backbone.width, device=device, dtype=torch.float32))
        self.cond_drop_prob = cond_drop_prob
        # The reference's cached_model_kwargs replaces a passed `embeddings` by zeros when no
        # `images` are given (transformer.py:440-446).  Set True to keep precomputed grids.
        self.accept_grid_embeddings = False

    def _prefix_layout(self):
        g = self.clip.grid_size ** 2
        return int(self.time_token_cond) + g + self.cond_ctx, (0 if self.time_token_cond else -1)

    def cached_model_kwargs(self, batch_size: int, model_kwargs: Dict[str, Any]) -> Dict[str, Any]:
        if "images" not in model_kwargs:
            if self.accept_grid_embeddings and "embeddings" in model_kwargs:
                return dict(embeddings=model_kwargs["embeddings"], low_res=model_kwargs["low_res"])
            zero_emb = torch.zeros([batch_size, self.clip.grid_feature_dim, self.clip.grid_size ** 2],
                                   device=next(self.parameters()).device)
            return dict(embeddings=zero_emb, low_res=model_kwargs["low_res"])
        with torch.no_grad():
            return dict(embeddings=self.clip.embed_images_grid(model_kwargs["images"]),
                        low_res=model_kwargs["low_res"])

    def _fill_cond(self, seqs, kw, prefix):
        low = self._embed_low_res(kw["low_res"])
        assert low.shape[1] == self.cond_ctx, f"low_res must have {self.cond_ctx} points"
        images, embeddings = kw.get("images"), kw.get("embeddings")
        g = self.clip.grid_size ** 2
        if images is not None:
            clip_out = self.clip.embed_images_grid(images)
        elif embeddings is not None:
            clip_out = embeddings
        else:  # unconditional generation (transformer.py:476-484)
            clip_out = torch.zeros([seqs, self.clip.grid_feature_dim, g], device=low.device)
        s0 = int(self.time_token_cond)
        prefix[:, s0:s0 + g].copy_(self._embed_grid(clip_out))
        prefix[:, s0 + g:s0 + g + self.cond_ctx].copy_(low)
        return None

    def forward(self, x, t, *, low_res, images=None, embeddings=None):
        kw = {k: v for k, v in dict(low_res=low_res, images=images, embeddings=embeddings).items()
              if v is not None}
        return self._model_forward(x, t, **kw)
