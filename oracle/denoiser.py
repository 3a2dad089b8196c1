"""Oracle restatement of the reference denoiser family (CPU, fp32, functional).

Weights are passed as a ``state_dict``-style mapping using the reference's own key
names (SURVEY.md 8a "weights contract").  Nothing here is used by the product.
"""
import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
LN_EPS = 1e-5  # nn.LayerNorm default, reference models/transformer.py:108,110,176,186


def timestep_embedding(t: Tensor, dim: int, max_period: float = 10000.0) -> Tensor:
    """reference models/util.py:72-89 -- [cos(t f_k) || sin(t f_k)], f_k = exp(-ln(P) k/half)."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32, device=t.device) / half)
    args = t[:, None].to(t.dtype) * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


def linear(x: Tensor, sd: Dict[str, Tensor], name: str) -> Tensor:
    return F.linear(x, sd[name + ".weight"], sd.get(name + ".bias"))


def layer_norm(x: Tensor, sd: Dict[str, Tensor], name: str) -> Tensor:
    w = sd[name + ".weight"]
    return F.layer_norm(x, (w.shape[0],), w, sd[name + ".bias"], LN_EPS)


def mlp(x: Tensor, sd, name: str) -> Tensor:
    """reference models/transformer.py:51-62 -- c_proj(gelu_erf(c_fc(x)))."""
    return linear(F.gelu(linear(x, sd, name + ".c_fc")), sd, name + ".c_proj")


def qkv_attention(qkv: Tensor, heads: int) -> Tensor:
    """reference models/transformer.py:65-84.

    qkv rows are laid out [H][q|k|v][hd]; q and k are each scaled by hd**-0.25,
    softmax is taken in fp32, output is [B, L, H*hd].
    """
    bs, n_ctx, width = qkv.shape
    hd = width // heads // 3
    scale = 1 / math.sqrt(math.sqrt(hd))
    qkv = qkv.view(bs, n_ctx, heads, 3 * hd)
    q, k, v = qkv[..., :hd], qkv[..., hd:2 * hd], qkv[..., 2 * hd:]
    w = torch.einsum("bthc,bshc->bhts", q * scale, k * scale)
    w = torch.softmax(w.float(), dim=-1).type(w.dtype)
    return torch.einsum("bhts,bshc->bthc", w, v).reshape(bs, n_ctx, -1)


def resblock(x: Tensor, sd, pre: str, heads: int) -> Tensor:
    """reference models/transformer.py:87-115 -- pre-LN residual attention + MLP."""
    a = qkv_attention(linear(layer_norm(x, sd, pre + ".ln_1"), sd, pre + ".attn.c_qkv"), heads)
    x = x + linear(a, sd, pre + ".attn.c_proj")
    x = x + mlp(layer_norm(x, sd, pre + ".ln_2"), sd, pre + ".mlp")
    return x


def forward_with_cond(sd, cfg, x: Tensor, cond_as_token) -> Tensor:
    """reference models/transformer.py:205-226."""
    h = linear(x.permute(0, 2, 1), sd, "input_proj")
    for emb, as_token in cond_as_token:
        if not as_token:
            h = h + emb[:, None]
    extra = [(e[:, None] if e.dim() == 2 else e) for e, as_token in cond_as_token if as_token]
    if extra:
        h = torch.cat(extra + [h], dim=1)
    h = layer_norm(h, sd, "ln_pre")
    for i in range(cfg["layers"]):
        h = resblock(h, sd, f"backbone.resblocks.{i}", cfg["heads"])
    h = layer_norm(h, sd, "ln_post")
    if extra:
        h = h[:, sum(e.shape[1] for e in extra):]
    h = linear(h, sd, "output_proj")
    return h.permute(0, 2, 1)


def embed_low_res(sd, low_res: Tensor) -> Tensor:
    """reference models/transformer.py:404-409."""
    x = low_res
    if sd.get("channel_scales") is not None:
        x = x * sd["channel_scales"][None, :, None]
    if sd.get("channel_biases") is not None:
        x = x + sd["channel_biases"][None, :, None]
    return linear(x.permute(0, 2, 1), sd, "cond_point_proj")


def embed_grid(sd, grid: Tensor) -> Tensor:
    """reference models/transformer.py:351-352 -- [B,1024,256] -> LN(1024) -> Linear -> 256 tokens."""
    g = grid.permute(0, 2, 1)
    return linear(layer_norm(g, sd, "clip_embed.0"), sd, "clip_embed.1")


def denoiser_forward(sd, cfg, x: Tensor, t: Tensor, *, embeddings: Optional[Tensor] = None,
                     low_res: Optional[Tensor] = None) -> Tensor:
    """Dispatch on cfg["name"] to the five reference classes (models/transformer.py)."""
    name = cfg["name"]
    width = cfg["width"]
    time_tok = cfg.get("time_token_cond", False)
    t_embed = mlp(timestep_embedding(t, width), sd, "time_embed")
    if name == "PointDiffusionTransformer":  # :195-203
        cond = [(t_embed, time_tok)]
    elif name == "CLIPImagePointDiffusionTransformer":  # :255-287
        if embeddings is None:
            embeddings = torch.zeros(x.shape[0], 768, device=x.device)
        clip_out = math.sqrt(embeddings.shape[1]) * embeddings
        cond = [(linear(clip_out, sd, "clip_embed"), cfg.get("token_cond", False)),
                (t_embed, time_tok)]
    elif name == "CLIPImageGridPointDiffusionTransformer":  # :323-355
        cond = [(t_embed, time_tok), (embed_grid(sd, embeddings), True)]
    elif name == "UpsamplePointDiffusionTransformer":  # :389-402
        cond = [(t_embed, time_tok), (embed_low_res(sd, low_res), True)]
    elif name == "CLIPImageGridUpsamplePointDiffusionTransformer":  # :453-494
        if embeddings is None:
            embeddings = torch.zeros(x.shape[0], 1024, 256, dtype=x.dtype, device=x.device)
        cond = [(t_embed, time_tok), (embed_grid(sd, embeddings), True),
                (embed_low_res(sd, low_res), True)]
    else:
        raise ValueError(name)
    return forward_with_cond(sd, cfg, x, cond)


# ---------------------------------------------------------------------------
# cross attention (reference models/perceiver.py)
# ---------------------------------------------------------------------------
def qkv_cross_attention(q: Tensor, kv: Tensor, heads: int) -> Tensor:
    """reference models/perceiver.py:46-67 -- kv rows laid out [H][k|v][hd]."""
    _, n_ctx, _ = q.shape
    bs, n_data, width = kv.shape
    hd = width // heads // 2
    scale = 1 / math.sqrt(math.sqrt(hd))
    q = q.view(bs, n_ctx, heads, -1)
    kv = kv.view(bs, n_data, heads, -1)
    k, v = kv[..., :hd], kv[..., hd:]
    w = torch.einsum("bthc,bshc->bhts", q * scale, k * scale)
    w = torch.softmax(w.float(), dim=-1).type(w.dtype)
    return torch.einsum("bhts,bshc->bthc", w, v).reshape(bs, n_ctx, -1)


def cross_resblock(x: Tensor, data: Tensor, sd, pre: str, heads: int) -> Tensor:
    """reference models/perceiver.py:70-104."""
    q = linear(layer_norm(x, sd, pre + ".ln_1"), sd, pre + ".attn.c_q")
    kv = linear(layer_norm(data, sd, pre + ".ln_2"), sd, pre + ".attn.c_kv")
    x = x + linear(qkv_cross_attention(q, kv, heads), sd, pre + ".attn.c_proj")
    x = x + mlp(layer_norm(x, sd, pre + ".ln_3"), sd, pre + ".mlp")
    return x


def perceiver_forward(sd, layers: int, heads: int, x: Tensor, data: Tensor) -> Tensor:
    """reference models/perceiver.py:107-146 (SimplePerceiver.forward)."""
    for i in range(layers):
        x = cross_resblock(x, data, sd, f"resblocks.{i}", heads)
    return x


# ---------------------------------------------------------------------------
# rotary point encoding (reference models/rotaryencoderpcd.py)
# ---------------------------------------------------------------------------
def apply_rotary(x: Tensor, coords: Tensor) -> Tensor:
    """reference models/rotaryencoderpcd.py:6-27 on one of q/k ([B,H,N,hd]).

    theta = pi * coords; out[0:3] = x_even*cos - x_odd*sin, out[3:6] = x_even*sin +
    x_odd*cos over head dims 0..5 (de-interleaved output); dims >= 6 untouched.
    """
    theta = coords * math.pi
    sin, cos = theta.sin().unsqueeze(1), theta.cos().unsqueeze(1)
    x1, x2 = x[..., 0:6:2], x[..., 1:6:2]
    return torch.cat([x1 * cos - x2 * sin, x1 * sin + x2 * cos, x[..., 6:]], dim=-1)


def rotary_self_attention(sd, heads: int, x: Tensor, pos: Tensor) -> Tensor:
    """reference models/rotaryencoderpcd.py:58-84 (qkv laid out [3][H][hd];
    logits scaled by MODEL width ** -0.5; softmax in the input dtype)."""
    B, N, D = x.shape
    hd = D // heads
    qkv = linear(x, sd, "qkv").reshape(B, N, 3, heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = apply_rotary(qkv[0], pos), apply_rotary(qkv[1], pos), qkv[2]
    attn = ((q @ k.transpose(-2, -1)) * (D ** -0.5)).softmax(dim=-1)
    out = (attn @ v).transpose(1, 2).reshape(B, N, D)
    return linear(out, sd, "out_proj")


def chamfer_distance_xyz(p1: Tensor, p2: Tensor) -> Tensor:
    """reference models/util.py:265-295 -- squared-L2 Chamfer on channels 0:3, [B]."""
    a = p1[:, :3, :].transpose(1, 2)
    b = p2[:, :3, :].transpose(1, 2)
    dist = torch.cdist(a, b, p=2).pow(2)
    return dist.min(dim=2)[0].mean(dim=1) + dist.min(dim=1)[0].mean(dim=1)


def cached_model_kwargs(cfg, batch_size: int, kw: Dict[str, Tensor]) -> Dict[str, Tensor]:
    """What the reference's ``cached_model_kwargs`` does for tensor-only kwargs.

    * CLIPImagePointDiffusionTransformer (transformer.py:251-253): runs the CLIP
      wrapper, which for ``embeddings=`` is a pass-through stack (pretrained_clip.py:95-107).
    * CLIPImageGridUpsamplePointDiffusionTransformer (transformer.py:440-451): with no
      ``images`` key the grid embedding is REPLACED by zeros (a passed ``embeddings``
      is dropped) and ``low_res`` is forwarded.
    * CLIPImageGridPointDiffusionTransformer (transformer.py:317-320) needs real images
      (KeyError otherwise) -- not reachable with synthetic embeddings.
    * The other classes have no ``cached_model_kwargs``.
    """
    name = cfg["name"]
    if name == "CLIPImagePointDiffusionTransformer":
        emb = kw.get("embeddings")
        out = torch.zeros(batch_size, 768)
        if emb is not None:
            out = torch.stack([e.to(out) for e in emb])
        return dict(embeddings=out)
    if name == "CLIPImageGridUpsamplePointDiffusionTransformer":
        return dict(embeddings=torch.zeros(batch_size, 1024, 256), low_res=kw["low_res"])
    if name == "CLIPImageGridPointDiffusionTransformer":
        raise KeyError("images")
    return kw
