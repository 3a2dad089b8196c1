"""CPU restatement (functional PyTorch, fp32) of the reference's TwoStreamDenoiser in eval mode
(models/model.py:437-547) and its recurrent-interface backbone (models/modules.py:17-244), over a
``state_dict`` with the reference's key names.  TEST INFRASTRUCTURE ONLY.

Modalities covered: "class" (ClassEmbedding, model.py:217-232), "view" (ViewAngleEmbedding, model.py:235-259),
"partial_pcd" (PartialPointCloudEncoder, model.py:262-338) and "depth" (DepthMapEncoder, model.py:341-434); the
last two are stacks of torch's nn.TransformerEncoderLayer / nn.TransformerDecoderLayer (norm_first, GELU,
batch_first), restated here from their documented arithmetic over the modules' own state_dict keys.
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
        prev_latent = torch.zeros(B, n_lat, zd, device=x.device)
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


def multihead_attention(sd, p, x_q, x_kv, heads):
    """torch.nn.MultiheadAttention (batch_first, no masks, eval): packed in_proj_weight [3d, d] = [Wq; Wk; Wv],
    softmax(q k^T / sqrt(hd)) v, out_proj."""
    B, Lq, d = x_q.shape
    Lk = x_kv.shape[1]
    hd = d // heads
    w, b = sd[p + ".in_proj_weight"], sd[p + ".in_proj_bias"]
    q = F.linear(x_q, w[:d], b[:d]).reshape(B, Lq, heads, hd).permute(0, 2, 1, 3)
    k = F.linear(x_kv, w[d:2 * d], b[d:2 * d]).reshape(B, Lk, heads, hd).permute(0, 2, 1, 3)
    v = F.linear(x_kv, w[2 * d:], b[2 * d:]).reshape(B, Lk, heads, hd).permute(0, 2, 1, 3)
    a = ((q @ k.transpose(-2, -1)) / math.sqrt(hd)).softmax(dim=-1) @ v
    return _lin(sd, p + ".out_proj", a.permute(0, 2, 1, 3).reshape(B, Lq, d))


def _ffn(sd, p, x):
    return _lin(sd, p + ".linear2", F.gelu(_lin(sd, p + ".linear1", x)))


def encoder_stack(sd, p, x, layers, heads):
    """nn.TransformerEncoder of norm_first GELU layers, no final norm: x += SA(LN1 x); x += FFN(LN2 x)."""
    for i in range(layers):
        l = f"{p}.layers.{i}"
        h = _ln(sd, l + ".norm1", x)
        x = x + multihead_attention(sd, l + ".self_attn", h, h, heads)
        x = x + _ffn(sd, l, _ln(sd, l + ".norm2", x))
    return x


def decoder_stack(sd, p, x, memory, layers, heads):
    """nn.TransformerDecoder of norm_first GELU layers, no masks, no final norm:
    x += SA(LN1 x); x += MHA(LN2 x, memory, memory); x += FFN(LN3 x)  (the memory is not normalised)."""
    for i in range(layers):
        l = f"{p}.layers.{i}"
        h = _ln(sd, l + ".norm1", x)
        x = x + multihead_attention(sd, l + ".self_attn", h, h, heads)
        x = x + multihead_attention(sd, l + ".multihead_attn", _ln(sd, l + ".norm2", x), memory, heads)
        x = x + _ffn(sd, l, _ln(sd, l + ".norm3", x))
    return x


ENC_LAYERS, ENC_HEADS = 8, 8  # constructor defaults TwoStreamDenoiser never overrides (model.py:264-266,343-344,355,369)


def _tokens_from_sequence(sd, p, x, B, stack):
    """Shared tail of both encoders (model.py:321-338, 417-434): CLS + sequence through the encoder stack, learned
    queries decoded against the non-CLS outputs, refined, CLS output prepended, proj_out + ln_out."""
    x = torch.cat([sd[p + ".cls_token"].expand(B, -1, -1), x], dim=1)
    x = encoder_stack(sd, f"{p}.{stack}", x, ENC_LAYERS, ENC_HEADS)
    q = sd[p + ".token_queries"].expand(B, -1, -1)
    tok = decoder_stack(sd, p + ".decoder", q, x[:, 1:], ENC_LAYERS // 2, ENC_HEADS)
    tok = tok + encoder_stack(sd, p + ".query_refiner", tok, ENC_LAYERS // 2, ENC_HEADS)
    return _ln(sd, p + ".ln_out", _lin(sd, p + ".proj_out", torch.cat([x[:, :1], tok], dim=1)))


def partial_pcd_encoder(sd, pcd, p="encoders.partial_pcd"):
    """PartialPointCloudEncoder.forward (model.py:318-338): pcd [B, N, 3] -> [B, num_tokens, d]."""
    return _tokens_from_sequence(sd, p, _lin(sd, p + ".input_proj", pcd), pcd.shape[0], "encoder")


def sincos_2d(h, w, dim, temperature=10000.0):
    """build_2d_sincos_position_embedding (model.py:190-213): [sin(x f) | cos(x f) | sin(y f) | cos(y f)],
    f_k = T^(-2k / (dim/2)), k < dim/4, over the row-major h x w grid."""
    yy, xx = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
    yy, xx = yy.reshape(-1, 1).float(), xx.reshape(-1, 1).float()
    f = torch.exp(torch.arange(0, dim // 2, 2).float() * -(math.log(temperature) / (dim // 4)))
    return torch.cat([torch.sin(xx * f), torch.cos(xx * f), torch.sin(yy * f), torch.cos(yy * f)], dim=1)


def depth_encoder(sd, depth, p="encoders.depth"):
    """DepthMapEncoder.forward (model.py:404-434): depth [B, 1, 512, 512] -> non-overlapping 32 x 32 patches
    (Conv2d with kernel = stride = patch, i.e. one matmul per patch) + the pos_embed buffer -> tokens."""
    w = sd[p + ".proj.weight"]                         # [d, C, P, P]
    B, C, H, W = depth.shape
    P = w.shape[-1]
    patches = depth.reshape(B, C, H // P, P, W // P, P).permute(0, 2, 4, 1, 3, 5).reshape(B, (H // P) * (W // P), C * P * P)
    x = F.linear(patches, w.reshape(w.shape[0], -1), sd[p + ".proj.bias"]) + sd[p + ".pos_embed"][None]
    return _tokens_from_sequence(sd, p, x, B, "mixer")


def cond_tokens(sd, cfg, B, class_labels=None, viewpoints=None, extra_cond=None, partial_pcd=None, depth_maps=None):
    """Eval-mode conditioning of TwoStreamDenoiser.forward (model.py:489-538) for the modalities in
    cfg["active_modalities"]: encoder tokens (zeros when the input is None or all zero) + masked token-type
    embeddings."""
    zd = cfg["latent_dim"]
    dev = sd["token_type_embeddings.weight"].device
    toks, types, masks = [], [], []
    for m in cfg["active_modalities"]:
        if m == "class":
            use = class_labels is not None and not bool(torch.all(class_labels == 0))
            tk = (_ln(sd, "encoders.class.norm", sd["encoders.class.embedding.weight"][class_labels]).unsqueeze(1)
                  if use else torch.zeros(B, 1, zd, device=dev))
            tid = 0
        elif m == "view":
            use = viewpoints is not None and not bool(torch.all(viewpoints == 0))
            if use:
                h = F.gelu(_lin(sd, "encoders.view.mlp.0", viewpoints))
                h = F.gelu(_lin(sd, "encoders.view.mlp.2", h))
                tk = _ln(sd, "encoders.view.mlp.5", _lin(sd, "encoders.view.mlp.4", h)).unsqueeze(1)
            else:
                tk = torch.zeros(B, 1, zd, device=dev)
            tid = 1
        elif extra_cond is not None and m in extra_cond:
            tk, use = extra_cond[m]
            tid = {"partial_pcd": 2, "depth": 3}[m]
        else:
            value, enc, tid = ((partial_pcd, partial_pcd_encoder, 2) if m == "partial_pcd" else (depth_maps, depth_encoder, 3))
            use = value is not None and not bool(torch.all(value == 0))
            n_tok = sd[f"encoders.{m}.token_queries"].shape[1] + 1
            tk = enc(sd, value) if use else torch.zeros(B, n_tok, zd, device=dev)
        toks.append(tk)
        types += [tid] * tk.shape[1]
        masks.append(torch.full((B, tk.shape[1], 1), 1.0 if use else 0.0, device=dev))
    te = sd["token_type_embeddings.weight"][torch.tensor(types, device=dev)].unsqueeze(0).expand(B, -1, -1)
    return torch.cat(toks, dim=1) + te * torch.cat(masks, dim=1)


def twostream_forward(sd, cfg, x, t, class_labels=None, viewpoints=None, prev_latent=None, extra_cond=None,
                      partial_pcd=None, depth_maps=None):
    """TwoStreamDenoiser.forward in eval mode: x [B, C, N] -> (x_denoised [B, C_out, N], latent)."""
    cond = cond_tokens(sd, cfg, x.shape[0], class_labels, viewpoints, extra_cond, partial_pcd, depth_maps)
    y, z = backbone_forward(sd, cfg, x.permute(0, 2, 1).contiguous(), t, cond, prev_latent)
    return y.permute(0, 2, 1).contiguous(), z
