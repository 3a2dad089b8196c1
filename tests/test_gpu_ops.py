"""Kernel-level parity (through the C ABI) against the CPU oracle / golden vectors."""
import math

import numpy as np
import pytest
import torch

from conftest import load_golden
from gpu_util import DEV, describe, rel
from oracle import denoiser as D
from oracle import cases, det

import pcd_b200 as P
ops = P.ops

pytestmark = pytest.mark.gpu


def bf16_round(t):
    return t.to(torch.bfloat16).float()


def test_device_is_b200():
    assert P._lib.load().pcd_check_device() == 0, P._lib.load().pcd_last_error()


def test_timestep_embedding_matches_reference():
    g = load_golden("ops")
    t_int = torch.tensor([0, 1, 17, 511, 1017, 1023], dtype=torch.long, device=DEV)
    t_flt = torch.tensor([0.5, 250.25, 999.75], dtype=torch.float32, device=DEV)
    for d in (128, 512):
        a = ops.timestep_embedding(t_int, d).cpu()
        assert (a - torch.from_numpy(g[f"temb_int_{d}"])).abs().max() < 2e-6, describe(a, g[f"temb_int_{d}"])
        b = ops.timestep_embedding(t_flt, d).cpu()
        assert (b - torch.from_numpy(g[f"temb_flt_{d}"])).abs().max() < 2e-6


@pytest.mark.parametrize("dim", [128, 192, 512, 768, 1024, 2048])
def test_layernorm(dim):
    x = det.normal((37, dim), 401 + dim, std=2.0) + 0.5
    w = 1.0 + det.uniform((dim,), 402, 0.2)
    b = det.uniform((dim,), 403, 0.2)
    want = torch.nn.functional.layer_norm(x, (dim,), w, b, 1e-5)
    got = ops.layernorm(x.to(DEV), w.to(DEV), b.to(DEV))
    assert rel(got, want) < 2e-6, describe(got, want)
    got16 = ops.layernorm(x.to(DEV), w.to(DEV), b.to(DEV), out_dtype=torch.bfloat16)
    assert got16.dtype == torch.bfloat16
    assert rel(got16.float(), want) < 4e-3, describe(got16.float(), want)


@pytest.mark.parametrize("ydtype", [torch.float32, torch.bfloat16])
def test_add_layernorm(ydtype):
    dim = 512
    h = det.normal((37, dim), 411, std=2.0)
    y = det.normal((37, dim), 412).to(ydtype)
    w = 1.0 + det.uniform((dim,), 413, 0.2)
    b = det.uniform((dim,), 414, 0.2)
    hs = h + y.float()
    want = torch.nn.functional.layer_norm(hs, (dim,), w, b, 1e-5)
    hd = h.to(DEV).clone()
    got = ops.add_layernorm(hd, y.to(DEV), w.to(DEV), b.to(DEV))
    assert rel(hd, hs) < 1e-7, describe(hd, hs, "h += y")
    assert rel(got, want) < 2e-6, describe(got, want)


GEMM_SHAPES = [  # M, N, K
    (128, 128, 64), (128, 256, 128), (256, 512, 512), (300, 384, 128), (1026, 1536, 512),
    (77, 256, 192), (5, 2048, 512), (1000, 128, 2048), (640, 72, 64), (129, 520, 72),
]


@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
@pytest.mark.parametrize("epi", [0, 1, 2])
def test_gemm_f32(M, N, K, epi):
    if K % 4:
        pytest.skip("K % 4")
    a = det.normal((M, K), 500 + M)
    w = det.uniform((N, K), 501 + N, 1 / math.sqrt(K))
    bias = det.uniform((N,), 502, 0.5)
    res = det.normal((M, N), 503)
    want = a.double() @ w.double().t() + bias.double()
    if epi == 1:
        want = torch.nn.functional.gelu(want)
    if epi == 2:
        want = want + res.double()
    got = ops.linear(a.to(DEV), w.to(DEV), bias.to(DEV), epilogue=epi, residual=res.to(DEV) if epi == 2 else None)
    assert got.shape == (M, N)
    assert rel(got, want) < 2e-6, describe(got, want, f"gemm_f32 {M}x{N}x{K} epi{epi}")


@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
@pytest.mark.parametrize("epi,out_dtype", [(0, torch.bfloat16), (0, torch.float32), (1, torch.bfloat16),
                                           (2, torch.float32)])
def test_gemm_bf16_tcgen05(M, N, K, epi, out_dtype):
    a = bf16_round(det.normal((M, K), 600 + M))
    w = bf16_round(det.uniform((N, K), 601 + N, 1 / math.sqrt(K)))
    bias = det.uniform((N,), 602, 0.5)
    res = det.normal((M, N), 603)
    want = a.double() @ w.double().t() + bias.double()
    if epi == 1:
        want = torch.nn.functional.gelu(want)
    if epi == 2:
        want = want + res.double()
    got = ops.linear(a.to(DEV).bfloat16(), w.to(DEV).bfloat16(), bias.to(DEV), epilogue=epi,
                     residual=res.to(DEV) if epi == 2 else None, out_dtype=out_dtype)
    torch.cuda.synchronize()
    assert got.dtype == out_dtype and got.shape == (M, N)
    tol = 3e-3 if out_dtype == torch.bfloat16 else 1e-5  # bf16 output rounding vs fp32 accumulate
    assert rel(got.float(), want) < tol, describe(got.float(), want, f"gemm_bf16 {M}x{N}x{K} epi{epi}")


def test_gemm_bf16_large_persistent():
    """More tiles than SMs, several accumulator phases per CTA, K > one stage ring."""
    M, N, K = 128 * 150 + 17, 512, 2048
    a = bf16_round(det.normal((M, K), 610))
    w = bf16_round(det.uniform((N, K), 611, 1 / math.sqrt(K)))
    bias = det.uniform((N,), 612, 0.5)
    got = ops.linear(a.to(DEV).bfloat16(), w.to(DEV).bfloat16(), bias.to(DEV), out_dtype=torch.float32)
    want = (a.to(DEV) @ w.to(DEV).t() + bias.to(DEV))
    assert rel(got, want) < 1e-4, describe(got, want)


def _slot_stats(hb):
    x = hb.float().view(hb.shape[0], -1, 128)
    mean = x.mean(dim=2)
    return torch.stack([mean, ((x - mean[..., None]) ** 2).sum(dim=2)], dim=2)


def test_cast_rowstats():
    """bf16 copy of the residual stream + (mean, M2) per 128 columns of the rounded values."""
    h = (det.normal((777, 512), 801) * 3.0 + 0.7).to(DEV)
    hb, stats = ops.cast_rowstats(h)
    torch.cuda.synchronize()
    assert torch.equal(hb, h.bfloat16())
    want = _slot_stats(hb)
    assert rel(stats, want) < 1e-5, describe(stats, want)


@pytest.mark.parametrize("dim,c_in,n_prefix,x_seqs,with_cond", [(512, 6, 2, 3, False), (512, 3, 0, 6, True),
                                                                 (1024, 6, 258, 2, False), (384, 6, 1, 6, False)])
def test_embed_tokens(dim, c_in, n_prefix, x_seqs, with_cond):
    """input_proj + concat + ln_pre (reference transformer.py:208-220) against torch, and the bf16 copy + row statistics
    the same launch emits for the LayerNorm-folded forward: bit-identical to pcd_cast_rowstats of its own fp32 output."""
    seqs, n = 6, 333
    x = det.normal((x_seqs, c_in, n), 840).to(DEV)
    w = det.uniform((dim, c_in), 841, 1 / math.sqrt(c_in)).to(DEV)
    b = det.uniform((dim,), 842, 0.5).to(DEV)
    prefix = det.normal((seqs, n_prefix, dim), 843).to(DEV) if n_prefix else None
    cond = det.normal((seqs, dim), 844).to(DEV) if with_cond else None
    g = (1.0 + 0.1 * det.normal((dim,), 845)).to(DEV)
    beta = (0.1 * det.normal((dim,), 846)).to(DEV)
    stats_ok = dim % 128 == 0
    if stats_ok:
        h, hb, stats = ops.embed_tokens(x, w, b, prefix, cond, g, beta, 1e-5, seqs=seqs, with_stats=True)
    else:
        h = ops.embed_tokens(x, w, b, prefix, cond, g, beta, 1e-5, seqs=seqs)
    torch.cuda.synchronize()
    xt = x.repeat(seqs // x_seqs, 1, 1).double()
    tok = xt.permute(0, 2, 1) @ w.double().t() + b.double()
    if cond is not None:
        tok = tok + cond.double()[:, None]
    if prefix is not None:
        tok = torch.cat([prefix.double(), tok], dim=1)
    want = torch.nn.functional.layer_norm(tok, (dim,), g.double(), beta.double(), 1e-5)
    assert rel(h, want) < 2e-6, describe(h, want)
    assert torch.equal(h, ops.embed_tokens(x, w, b, prefix, cond, g, beta, 1e-5, seqs=seqs)), "same fp32 stream either way"
    if stats_ok:
        hb2, stats2 = ops.cast_rowstats(h.view(-1, dim))
        assert torch.equal(hb.view(-1, dim), hb2) and torch.equal(stats, stats2)
    else:
        with pytest.raises(RuntimeError, match="128"):
            ops.embed_tokens(x, w, b, prefix, cond, g, beta, 1e-5, seqs=seqs, with_stats=True)


@pytest.mark.parametrize("dim,c_out,n_prefix,with_y", [(512, 6, 2, False), (512, 3, 0, False), (384, 6, 1, False),
                                                        (512, 12, 2, False), (512, 6, 2, True), (1024, 6, 258, False)])
def test_output_proj(dim, c_out, n_prefix, with_y):
    """ln_post + slice + output_proj + permute (reference transformer.py:222-226) against torch in fp64, for the
    register-resident kernel (no pending residual, width <= 512, 3 / 6 channels) and the general one."""
    seqs, n = 5, 333
    h = (det.normal((seqs, n_prefix + n, dim), 850) * 2.0 + 0.3).to(DEV)
    y = det.normal((seqs, n_prefix + n, dim), 851).to(DEV) if with_y else None
    g = (1.0 + 0.1 * det.normal((dim,), 852)).to(DEV)
    beta = (0.1 * det.normal((dim,), 853)).to(DEV)
    w = det.uniform((c_out, dim), 854, 1 / math.sqrt(dim)).to(DEV)
    b = det.uniform((c_out,), 855, 0.5).to(DEV)
    got = ops.output_proj(h, n_prefix, g, beta, w, b, 1e-5, y=y)
    torch.cuda.synchronize()
    hh = h.double() + (y.double() if with_y else 0.0)
    xn = torch.nn.functional.layer_norm(hh, (dim,), g.double(), beta.double(), 1e-5)[:, n_prefix:]
    want = (xn @ w.double().t() + b.double()).permute(0, 2, 1)
    assert got.shape == want.shape and rel(got, want) < 2e-6, describe(got, want)


@pytest.mark.parametrize("M,N,K", [(128 * 37 + 5, 512, 512), (2048, 512, 2048), (1500, 1024, 1024), (700, 2048, 512)])
def test_gemm_residual_stats(M, N, K):
    """PCD_EPI_RESIDUAL_STATS: h <- h + A W^T + b in place, bf16 copy and row statistics
    (reference transformer.py:113-114, the residual update of a block)."""
    a = bf16_round(det.normal((M, K), 810)).to(DEV).bfloat16()
    w = bf16_round(det.uniform((N, K), 811, 1 / math.sqrt(K))).to(DEV).bfloat16()
    bias = det.uniform((N,), 812, 0.5).to(DEV)
    h0 = (det.normal((M, N), 813) * 2.0).to(DEV)
    want = h0.double() + a.double() @ w.double().t() + bias.double()
    h = h0.clone()
    hb, stats = ops.linear_residual_stats(a, w, bias, h)
    torch.cuda.synchronize()
    assert rel(h, want) < 1e-5, describe(h, want)
    assert torch.equal(hb, h.bfloat16()), "bf16 copy must be the rounding of the fp32 result"
    ws = _slot_stats(hb)
    assert rel(stats[..., 0], ws[..., 0]) < 1e-4 and rel(stats[..., 1], ws[..., 1]) < 1e-4, describe(stats, ws)


@pytest.mark.parametrize("gelu", [False, True])
@pytest.mark.parametrize("M,N,K", [(128 * 9 + 77, 1536, 512), (1026 * 2, 2048, 512), (600, 1024, 1024),
                                   (520, 512, 2048)])   # width 2048 (base1B): 16 statistics slots, un-prefetched path
def test_gemm_layernorm_folded(M, N, K, gelu):
    """PCD_EPI_LN_BIAS(_GELU): LayerNorm folded into the projection (transformer.py:108-114) against
    torch's LayerNorm -> Linear (-> GELU) in fp64 on the same bf16 activations."""
    h = (det.normal((M, K), 820) * 1.7 + det.normal((M, 1), 821) * 0.5).to(DEV)   # rows with non-zero means
    gamma = (1.0 + 0.3 * det.normal((K,), 822)).to(DEV)
    beta = (0.2 * det.normal((K,), 823)).to(DEV)
    w = bf16_round(det.uniform((N, K), 824, 1 / math.sqrt(K))).to(DEV)
    b = det.uniform((N,), 825, 0.5).to(DEV)
    hb, stats = ops.cast_rowstats(h)
    wf = (w * gamma[None, :]).bfloat16().contiguous()
    colsum = wf.float().sum(dim=1).contiguous()
    const = (w.double() @ beta.double() + b.double()).float().contiguous()
    got = ops.linear_layernorm_folded(hb, stats, wf, colsum, const, eps=1e-5, gelu=gelu)
    torch.cuda.synchronize()
    xn = torch.nn.functional.layer_norm(hb.double(), (K,), gamma.double(), beta.double(), 1e-5)
    want = xn @ w.double().t() + b.double()
    if gelu:
        want = torch.nn.functional.gelu(want)
    # bf16 output rounding + bf16 rounding of gamma o W (vs. rounding LayerNorm's output in the unfused path)
    assert rel(got.float(), want) < 6e-3, describe(got.float(), want, f"lnfold {M}x{N}x{K} gelu={gelu}")


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 5e-6), (torch.bfloat16, 1e-2)])
@pytest.mark.parametrize("variant", [0, 8, 9, 5])
def test_self_attention_golden(dtype, tol, variant):
    if dtype == torch.float32 and variant != 0:
        pytest.skip("variant only affects the tensor-core kernel")
    g = load_golden("ops")
    for name, shape, seed, std, heads in (("self_attn", (2, 70, 384), 301, 1.0, 2),
                                          ("self_attn_h8", (1, 300, 1536), 302, 2.0, 8)):
        qkv = det.normal(shape, seed, std=std)
        want = torch.from_numpy(g[name])
        if dtype == torch.bfloat16:
            want = D.qkv_attention(bf16_round(qkv), heads)
        got = ops.self_attention(qkv.to(DEV).to(dtype), heads, variant=variant)
        torch.cuda.synchronize()
        assert rel(got.float(), want) < tol, describe(got.float(), want, f"{name} {dtype} v{variant}")


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 5e-6), (torch.bfloat16, 1e-2)])
@pytest.mark.parametrize("L", [1, 16, 127, 128, 129, 256, 1026])
def test_self_attention_lengths(dtype, tol, L):
    """Tile tails: L = 1 .. 1026 (the north-star L = 1024 + 2 prefix tokens)."""
    heads, B = 2, 2
    qkv = det.normal((B, L, heads * 192), 700 + L, std=1.5)
    src = bf16_round(qkv) if dtype == torch.bfloat16 else qkv
    want = D.qkv_attention(src, heads)
    got = ops.self_attention(qkv.to(DEV).to(dtype), heads)
    torch.cuda.synchronize()
    assert rel(got.float(), want) < tol, describe(got.float(), want, f"L={L} {dtype}")


@pytest.mark.parametrize("variant", [8, 11, 10, 9, 7, 6, 5])
@pytest.mark.parametrize("B,heads,L,std", [(2, 2, 1, 1.5), (2, 2, 63, 1.5), (3, 2, 64, 1.5), (2, 3, 65, 1.5),
                                           (2, 3, 66, 1.5), (2, 3, 68, 2.5),  # 1 full KV tile + a 2- / 4-key tail
                                           (2, 3, 69, 1.5), (3, 2, 196, 1.5),  # 5-key tail (16-column form); 3 x 64 + 4 keys
                                           (2, 3, 80, 1.5), (2, 3, 81, 1.5), (3, 2, 141, 2.5),  # 16 keys; 17 (a step); 13
                                           (2, 2, 127, 1.5), (2, 2, 129, 1.5), (1, 2, 1026, 1.5),
                                           (3, 2, 257, 1.5),     # 3 query tiles, the third holding one row
                                           (2, 2, 385, 1.5),     # a second query group with a single one-row tile
                                           (2, 2, 640, 1.5),     # groups of 3 + 2 query tiles
                                           (40, 8, 200, 1.0),    # 640 items on the persistent CTAs
                                           (33, 8, 1026, 2.5),   # 9 query tiles x 17 KV tiles (odd), several items / CTA
                                           (2, 8, 4353, 1.0),    # upsampler length: 35 query tiles (11 groups of 3 + 2)
                                           (5, 16, 1281, 4.0)])  # large logits: lazy rescaling path
def test_attention_bf16_variants_vs_torch(variant, B, heads, L, std):
    """Tensor-core attention variants against an fp32 torch evaluation of the same bf16 inputs, at
    shapes that give a persistent CTA several (query tile, head, sequence) items, odd KV-tile
    counts and ragged query / key tails."""
    lib = P._lib.load()
    gen = torch.Generator(device="cpu").manual_seed(1000 + L + heads)
    qkv = (torch.randn(B, L, heads * 192, generator=gen) * std).to(DEV).bfloat16()
    x = qkv.float().view(B, L, heads, 192)
    q, k, v = x[..., :64], x[..., 64:128], x[..., 128:]
    want = torch.empty(B, L, heads * 64, device=DEV)
    for b0 in range(0, B, 8):
        w = torch.einsum("bthc,bshc->bhts", q[b0:b0 + 8], k[b0:b0 + 8]) / 8.0
        want[b0:b0 + 8] = torch.einsum("bhts,bshc->bthc", torch.softmax(w, -1), v[b0:b0 + 8]).reshape(-1, L, heads * 64)
    got = ops.self_attention(qkv, heads, variant=variant)
    torch.cuda.synchronize()
    assert torch.isfinite(got.float()).all()
    assert rel(got.float(), want) < 1e-2, describe(got.float(), want, f"v{variant} B{B} H{heads} L{L}")
    # per-sequence check so that one bad item cannot hide in the norm of many good ones
    per = ((got.float() - want).flatten(1).norm(dim=1) / want.flatten(1).norm(dim=1)).max()
    assert float(per) < 1e-2, float(per)


@pytest.mark.parametrize("variant", [0, 8, 10, 11])
@pytest.mark.parametrize("B,H,Lq,Lkv", [(3, 8, 643, 643), (2, 8, 1024, 643), (2, 8, 643, 1024), (4, 2, 70, 5), (1, 1, 1, 64),
                                        (64, 8, 257, 129)])
def test_attention_hd32_bf16(variant, B, H, Lq, Lkv):
    """32-wide heads (TwoStream CrossAttention, reference models/modules.py:17-63) on the 64-wide tensor-core kernel:
    compact [.., H, 32] operands, the missing 32 columns of every tile zero-filled by TMA, output written compactly.
    Against an fp32 torch evaluation of the same bf16 inputs; operands are column blocks of one fused buffer."""
    gen = torch.Generator(device="cpu").manual_seed(5000 + Lq + Lkv)
    D = H * 32
    qb = (torch.randn(B, Lq, D, generator=gen) * 1.4).to(DEV).bfloat16()
    kvb = (torch.randn(B, Lkv, 2 * D, generator=gen) * 1.4).to(DEV).bfloat16()
    q, k, v = qb, kvb[..., :D], kvb[..., D:]
    s = 32.0 ** -0.25
    w = torch.einsum("bthc,bshc->bhts", q.float().view(B, Lq, H, 32), k.float().reshape(B, Lkv, H, 32)) * (s * s)
    want = torch.einsum("bhts,bshc->bthc", torch.softmax(w, -1), v.float().reshape(B, Lkv, H, 32)).reshape(B, Lq, D)
    got = ops.attention_views(q, k, v, H, s, s, variant=variant, head_dim=32)
    torch.cuda.synchronize()
    assert got.shape == (B, Lq, D) and torch.isfinite(got.float()).all()
    assert rel(got.float(), want) < 1e-2, describe(got.float(), want, f"hd32 v{variant} B{B} H{H} Lq{Lq} Lkv{Lkv}")
    per = ((got.float() - want).flatten(1).norm(dim=1) / want.flatten(1).norm(dim=1)).max()
    assert float(per) < 1e-2, float(per)
    with pytest.raises(RuntimeError):
        ops.attention_views(q, k, v, H, s, s, variant=5, head_dim=32)   # the paired kernel stores from registers


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_attention_bf16_random_lengths(seed):
    """The dispatcher over random (len_q, len_kv, sequences, heads): every tail length 0..63 class (tail-key forms for
    0..4 and 5..16 keys, masked last step beyond), 1..n KV steps, ragged query tiles, cross-attention shapes; the default
    kernel choice and the forced grouped kernel against an fp32 torch evaluation of the same bf16 inputs."""
    rng = torch.Generator(device="cpu").manual_seed(4242 + seed)
    ri = lambda lo, hi: int(torch.randint(lo, hi + 1, (1,), generator=rng))
    for case in range(14):
        B, H = ri(1, 5), ri(1, 4)
        lq = ri(1, 700)
        lkv = (ri(1, 9) * 64 + (ri(0, 20) if case % 2 == 0 else ri(0, 63))) if case % 5 else ri(1, 63)
        q = (torch.randn(B, lq, H * 64, generator=rng) * 1.3).to(DEV).bfloat16()
        k = (torch.randn(B, lkv, H * 64, generator=rng) * 1.3).to(DEV).bfloat16()
        v = torch.randn(B, lkv, H * 64, generator=rng).to(DEV).bfloat16()
        w = torch.einsum("bthc,bshc->bhts", q.float().view(B, lq, H, 64), k.float().view(B, lkv, H, 64)) / 8.0
        want = torch.einsum("bhts,bshc->bthc", torch.softmax(w, -1), v.float().view(B, lkv, H, 64)).reshape(B, lq, H * 64)
        for variant in (0, 8, 11):
            got = ops.attention_views(q, k, v, H, 64 ** -0.25, 64 ** -0.25, variant=variant)
            torch.cuda.synchronize()
            assert torch.isfinite(got.float()).all(), (B, H, lq, lkv, variant)
            assert rel(got.float(), want) < 1e-2, describe(got.float(), want, f"B{B} H{H} Lq{lq} Lkv{lkv} v{variant}")


@pytest.mark.parametrize("variant", [8, 11, 10, 9, 5])
@pytest.mark.parametrize("late", [1, 5, 16])
def test_attention_rereferences_rows_when_later_tiles_dominate(variant, late):
    """Online softmax with a lazily updated reference point: keys from tile `late` on are scaled so that their logits
    sit tens of binades above everything seen before (and some rows overflow an un-rescaled exponential), for every
    row at a different tile.  The kernels must re-reference (rescale O and l) and still match an fp32 softmax."""
    B, H, L = 2, 4, 1026
    gen = torch.Generator(device="cpu").manual_seed(77 + late)
    qkv = torch.randn(B, L, H * 192, generator=gen)
    x = qkv.view(B, L, H, 192)
    boost = torch.ones(L)
    boost[late * 64:] = 6.0
    boost[(late + 3) * 64 + 7:] = 30.0      # a second jump in the middle of a tile (if it exists)
    x[..., 64:128] *= boost[None, :, None, None]
    qkv = x.reshape(B, L, H * 192).to(DEV).bfloat16()
    xf = qkv.float().view(B, L, H, 192)
    q, k, v = xf[..., :64], xf[..., 64:128], xf[..., 128:]
    w = torch.einsum("bthc,bshc->bhts", q, k) / 8.0
    want = torch.einsum("bhts,bshc->bthc", torch.softmax(w, -1), v).reshape(B, L, H * 64)
    got = ops.self_attention(qkv, H, variant=variant)
    torch.cuda.synchronize()
    assert torch.isfinite(got.float()).all()
    assert rel(got.float(), want) < 1e-2, describe(got.float(), want, f"v{variant} late={late}")


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 5e-6), (torch.bfloat16, 1e-2)])
def test_cross_attention_golden(dtype, tol):
    g = load_golden("ops")
    q, kv = det.normal((2, 70, 128), 303), det.normal((2, 77, 256), 304)
    want = torch.from_numpy(g["cross_attn"])
    if dtype == torch.bfloat16:
        want = D.qkv_cross_attention(bf16_round(q), bf16_round(kv), 2)
    got = ops.cross_attention(q.to(DEV).to(dtype), kv.to(DEV).to(dtype), 2)
    assert rel(got.float(), want) < tol, describe(got.float(), want)


def test_rotary_attention_golden():
    g = load_golden("ops")
    coords = det.uniform((2, 70, 3), 307, std=0.5 / 3 ** 0.5)
    from test_oracle_golden import TOL  # noqa: F401
    sd = det.fill_state_dict({"qkv.weight": torch.zeros(384, 128), "qkv.bias": torch.zeros(384),
                              "out_proj.weight": torch.zeros(128, 128), "out_proj.bias": torch.zeros(128)}, 308)
    mod = P.rotaryencoderpcd.RotarySelfAttention(128, heads=2).to(DEV)
    mod.load_state_dict(sd)
    got = mod(det.normal((2, 70, 128), 309).to(DEV), coords.to(DEV))
    assert rel(got, g["rotary_self_attn"]) < 1e-5, describe(got, g["rotary_self_attn"])
    # q-side rotation alone: attention with V = identity-like probes is indirect; check the
    # rotation itself through a 1-key problem where softmax == 1 and out == v
    qkv = det.normal((1, 1, 384), 310)
    out = ops.rotary_attention(qkv.to(DEV), coords[:1, :1].to(DEV), 2)
    assert rel(out, qkv[..., 256:]) < 1e-6


def test_rotary_attention_bf16():
    """bf16 mode of RotarySelfAttention (rotaryencoderpcd.py:58-84): tensor-core projections and ONE attention launch
    that rotates q / k inside the kernel (shared-memory tiles, between TMA arrival and the first MMA), against the
    reference golden, against the fp32 rotary kernel on the same values, and bit for bit against the stand-alone
    rotation (pcd_rope_bf16) followed by plain attention."""
    g = load_golden("ops")
    lib = P._lib.load()
    coords = det.uniform((2, 70, 3), 307, std=0.5 / 3 ** 0.5)
    sd = det.fill_state_dict({"qkv.weight": torch.zeros(384, 128), "qkv.bias": torch.zeros(384),
                              "out_proj.weight": torch.zeros(128, 128), "out_proj.bias": torch.zeros(128)}, 308)
    mod = P.rotaryencoderpcd.RotarySelfAttention(128, heads=2, dtype=torch.bfloat16).to(DEV)
    mod.load_state_dict(sd)
    got = mod(det.normal((2, 70, 128), 309).to(DEV), coords.to(DEV))
    assert got.dtype == torch.float32
    assert rel(got, g["rotary_self_attn"]) < 2e-2, describe(got, g["rotary_self_attn"])
    # kernel level: several query groups, ragged tails, bf16 rotary attention vs the fp32 rotary kernel
    for B, N, H in ((3, 333, 4), (2, 1026, 2), (1, 64, 1), (2, 131, 2), (2, 256, 2), (40, 1025, 4)):
        qkv = bf16_round(det.normal((B, N, 3 * H * 64), 330 + N, std=1.2)).to(DEV)
        pos = det.uniform((B, N, 3), 331 + N, std=0.4).to(DEV)
        want = ops.rotary_attention(qkv, pos, H)
        keep = qkv.bfloat16()
        snapshot = keep.clone()
        n0 = lib.pcd_launch_count()
        got2 = ops.rotary_attention(keep, pos, H)
        assert lib.pcd_launch_count() - n0 == 1, "rotary attention must be a single launch"
        torch.cuda.synchronize()
        assert torch.equal(keep, snapshot), "the caller's qkv must not be modified"
        assert rel(got2.float(), want) < 1e-2, describe(got2.float(), want, f"B{B} N{N} H{H}")
        # the fused rotation computes exactly what the stand-alone one does
        D = H * 64
        rot = keep.clone()
        ops.rope_bf16_(rot[..., :D], pos, H)
        ops.rope_bf16_(rot[..., D:2 * D], pos, H)
        # (same kernel on both sides: rotary attention always takes the grouped one, a small plain problem would go to the
        # paired kernel, whose tail keys are summed in a different order)
        plain = ops.attention_views(rot[..., :D], rot[..., D:2 * D], rot[..., 2 * D:], H, D ** -0.5, 1.0,
                                    variant=P._lib.ATTN_GROUPED)
        torch.cuda.synchronize()
        assert torch.equal(plain, got2), describe(got2.float(), plain.float(), "fused vs stand-alone rotation")


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-5), (torch.bfloat16, 2e-2)])
def test_perceiver_golden(dtype, tol):
    from test_oracle_golden import perceiver_shapes
    g = load_golden("ops")
    sd = det.fill_state_dict(perceiver_shapes(128, 2, 192), 310)
    per = P.perceiver.SimplePerceiver(device=DEV, dtype=dtype, n_data=77, width=128, layers=2, heads=2,
                                      data_width=192)
    per.load_state_dict(sd)
    got = per(det.normal((2, 70, 128), 311).to(DEV), det.normal((2, 77, 192), 312).to(DEV))
    assert rel(got, g["perceiver"]) < tol, describe(got, g["perceiver"], f"perceiver {dtype}")


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-4), (torch.bfloat16, 2e-2)])
def test_perceiver_text_shape(dtype, tol):
    """BASELINE config 3 (ii): SimplePerceiver with the 1026 denoiser tokens (width 512, 8 heads) as
    queries over 77 CLIP text tokens of width 768, against the unmodified reference.  bf16 runs the LayerNorm-folded
    path (no LayerNorm launch on the query stream) with the step-invariant c_kv(ln_2(data)) cached per data tensor."""
    from test_oracle_golden import perceiver_shapes
    c = cases.PERCEIVER_TEXT
    g = load_golden("perceiver_text")
    per = P.perceiver.SimplePerceiver(device=DEV, dtype=dtype, n_data=c["n_data"], width=c["width"], layers=c["layers"],
                                      heads=c["heads"], data_width=c["data_width"])
    per.load_state_dict(det.fill_state_dict(perceiver_shapes(c["width"], c["layers"], c["data_width"]), c["seed"]))
    x, data = cases.perceiver_text_inputs()
    x, data = x.to(DEV), data.to(DEV)
    lib = P._lib.load()
    n0 = lib.pcd_launch_count()
    got = per(x, data)
    first = lib.pcd_launch_count() - n0
    want = torch.from_numpy(g["out"])
    assert rel(got[:, ::c["row_stride"]], want) < tol, describe(got[:, ::c["row_stride"]], want, f"perceiver text {dtype}")
    # second evaluation with the same conditioning tokens (what a sampler does 254 times): the key / value side is
    # reused, the result is identical, and the bf16 path is five launches per block plus one cast of the stream
    n0 = lib.pcd_launch_count()
    again = per(x, data)
    second = lib.pcd_launch_count() - n0
    assert torch.equal(again, got)
    assert second == first - 2 * c["layers"], (first, second)   # ln_2 + c_kv per block dropped
    if dtype == torch.bfloat16:
        assert second == 5 * c["layers"] + 1, second
    # new conditioning values in the same tensor (version bump) must not hit the cache
    data.add_(1.0)
    moved = per(x, data)
    assert not torch.equal(moved, got)


def test_chamfer_golden():
    g = load_golden("ops")
    got = ops.chamfer_distance_xyz(det.uniform((2, 6, 200), 313, 0.3).to(DEV), det.uniform((2, 3, 150), 314, 0.3).to(DEV))
    assert rel(got, g["chamfer"]) < 1e-4, describe(got, g["chamfer"])
