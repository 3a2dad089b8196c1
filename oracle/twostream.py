"""CPU restatement (functional PyTorch, fp32) of the reference's TwoStreamDenoiser in eval mode
(models/model.py:437-547) and its recurrent-interface backbone (models/modules.py:17-244), over a
``state_dict`` with the reference's key names.  TEST INFRASTRUCTURE ONLY.

Modalities covered: "class" (ClassEmbedding, model.py:217-232) and "view" (ViewAngleEmbedding,
model.py:235-259); the partial-cloud and depth encoders (nn.TransformerEncoder/Decoder stacks) are not
restated yet -- their tokens can be passed in precomputed through ``extra_cond``.
Pinned against the unmodified reference by tests/golden/twostream_*.npz (oracle/make_golden_twostream.py)."""
import math

import torch
import torch.nn.functional as F

from oracle.denoiser import timestep_embedding


def _lin(sd, p, x):
    return F.linear(x, sd[p + ".weight"], sd.get(p + ".bias"))


def _ln(sd, p, x, eps=1e-5):
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], eps)


def _mlp(sd, p, x):
    """timm Mlp: fc1 -> GELU (exact erf) -> fc2 (models/modules.py:14; dropout is identity in eval)."""
    return _lin(sd, p + ".fc2", F.gelu(_lin(sd, p + ".fc1", x)))


def cross_attention(sd, p, x_q, x_kv, heads):
    """CrossAttention.forward (models/modules.py:40-63): separate wq / wk / wv, softmax(q k^T * hd^-1/2) v, proj."""
    B, Nq, C = x_q.shape
    Nkv = x_kv.shape[1]
    hd = C // heads
    q = _lin(sd, p + ".wq", x_q).reshape(B, Nq, heads, hd).permute(0, 2, 1, 3)
    k = _lin(sd, p + ".wk", x_kv).reshape(B, Nkv, heads, hd).permute(0, 2, 1, 3)
    v = _lin(sd, p + ".wv", x_kv).reshape(B, Nkv, heads, hd).permute(0, 2, 1, 3)
    attn = ((q @ k.transpose(-2, -1)) * hd ** -0.5).softmax(dim=-1)
    return _lin(sd, p + ".proj", (attn @ v).transpose(1, 2).reshape(B, Nq, C))


def backbone_forward(sd, cfg, x, t, cond, prev_latent, p="denoiser_backbone"):
    """Denoiser_backbone.forward (models/modules.py:198-244): x [B, N, C_in] -> ([B, N, C_out], latent)."""
    B = x.shape[0]
    zd, heads = cfg["latent_dim"], cfg["num_heads"]
    n_lat = cfg["num_latents"] + cond.shape[1] + 1
    if prev_latent is None:
        prev_latent = torch.zeros(B, n_lat, zd)
    t_embed = _mlp(sd, p + ".time_embed", timestep_embedding(t, zd)).unsqueeze(1)
    x = _ln(sd, p + ".ln_pre", _lin(sd, p + ".input_proj", x))
    z = torch.cat([sd[p + ".z_init"].repeat(B, 1, 1), cond, t_embed], dim=1)
    prev_latent = prev_latent + _mlp(sd, p + ".latent_mlp", prev_latent)
    z = z + _ln(sd, p + ".ln_latent", prev_latent)
    for i in range(cfg["num_blocks"]):
        b = f"{p}.blocks.{i}"
        # read (modules.py:96-99)
        z = z + cross_attention(sd, b + ".read.attn", _ln(sd, b + ".read.norm_z1", z), _ln(sd, b + ".read.norm_x", x), heads)
        z = z + _mlp(sd, b + ".read.mlp", _ln(sd, b + ".read.norm_z2", z))
        # compute (modules.py:76-80)
        for j in range(cfg["num_compute_layers"]):
            c = f"{b}.compute.{j}"
            zn = _ln(sd, c + ".norm_z1", z)
            z = z + cross_attention(sd, c + ".attn", zn, zn, heads)
            z = z + _mlp(sd, c + ".mlp", _ln(sd, c + ".norm_z2", z))
        # write (modules.py:117-120)
        x = x + cross_attention(sd, b + ".write.attn", _ln(sd, b + ".write.norm_x1", x), _ln(sd, b + ".write.norm_z", z), heads)
        x = x + _mlp(sd, b + ".write.mlp", _ln(sd, b + ".write.norm_x2", x))
    return _lin(sd, p + ".output_proj", _ln(sd, p + ".ln_post", x)), z


def cond_tokens(sd, cfg, B, class_labels=None, viewpoints=None, extra_cond=None):
    """Eval-mode conditioning of TwoStreamDenoiser.forward (model.py:489-538) for the modalities in
    cfg["active_modalities"]: encoder tokens (zeros when the input is None or all zero) + masked token-type
    embeddings."""
    zd = cfg["latent_dim"]
    toks, types, masks = [], [], []
    for m in cfg["active_modalities"]:
        if m == "class":
            use = class_labels is not None and not bool(torch.all(class_labels == 0))
            tk = (_ln(sd, "encoders.class.norm", sd["encoders.class.embedding.weight"][class_labels]).unsqueeze(1)
                  if use else torch.zeros(B, 1, zd))
            tid = 0
        elif m == "view":
            use = viewpoints is not None and not bool(torch.all(viewpoints == 0))
            if use:
                h = F.gelu(_lin(sd, "encoders.view.mlp.0", viewpoints))
                h = F.gelu(_lin(sd, "encoders.view.mlp.2", h))
                tk = _ln(sd, "encoders.view.mlp.5", _lin(sd, "encoders.view.mlp.4", h)).unsqueeze(1)
            else:
                tk = torch.zeros(B, 1, zd)
            tid = 1
        else:
            tk, use = extra_cond[m]
            tid = {"partial_pcd": 2, "depth": 3}[m]
        toks.append(tk)
        types += [tid] * tk.shape[1]
        masks.append(torch.full((B, tk.shape[1], 1), 1.0 if use else 0.0))
    te = sd["token_type_embeddings.weight"][torch.tensor(types)].unsqueeze(0).expand(B, -1, -1)
    return torch.cat(toks, dim=1) + te * torch.cat(masks, dim=1)


def twostream_forward(sd, cfg, x, t, class_labels=None, viewpoints=None, prev_latent=None, extra_cond=None):
    """TwoStreamDenoiser.forward in eval mode: x [B, C, N] -> (x_denoised [B, C_out, N], latent)."""
    cond = cond_tokens(sd, cfg, x.shape[0], class_labels, viewpoints, extra_cond)
    y, z = backbone_forward(sd, cfg, x.permute(0, 2, 1).contiguous(), t, cond, prev_latent)
    return y.permute(0, 2, 1).contiguous(), z
