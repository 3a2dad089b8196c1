"""Whole-denoiser parity: pcd_model_forward vs golden outputs of the unmodified reference."""
import pytest
import torch

from conftest import load_golden
from gpu_util import DEV, TOL_BF16, TOL_F32, build_model, describe, rel, to_dev
from oracle import cases

pytestmark = pytest.mark.gpu

SMALL = [c for c in cases.FORWARD_CASES if c.startswith("small")]


@pytest.mark.parametrize("case", SMALL)
@pytest.mark.parametrize("dtype,tol", [(torch.float32, TOL_F32), (torch.bfloat16, TOL_BF16)])
def test_forward_small(case, dtype, tol):
    g = load_golden("forward_" + case)
    model, cfg, _ = build_model(case, dtype)
    x, t, kw = cases.forward_inputs(case)
    with torch.no_grad():
        y = model(x.to(DEV), t.to(DEV), **to_dev(kw))
    torch.cuda.synchronize()
    assert y.shape == g["out"].shape and y.dtype == torch.float32
    assert rel(y, g["out"]) < tol, describe(y, g["out"], f"{case} {dtype}")
    # second call exercises the cached-conditioning path
    with torch.no_grad():
        y2 = model(x.to(DEV), t.to(DEV), **to_dev(kw))
    assert rel(y2, g["out"]) < tol


@pytest.mark.parametrize("case", ["full_imagevec", "full_upsample", "full_base300M"])
@pytest.mark.parametrize("dtype,tol", [(torch.bfloat16, TOL_BF16), (torch.float32, TOL_F32)])
def test_forward_full_size(case, dtype, tol):
    """BASELINE.json shapes: base40M-imagevec (L=1026), upsample (L=4353), base300M (L=1281)."""
    g = load_golden("forward_" + case)
    model, cfg, _ = build_model(case, dtype)
    x, t, kw = cases.forward_inputs(case)
    with torch.no_grad():
        y = model(x.to(DEV), t.to(DEV), **to_dev(kw))
    torch.cuda.synchronize()
    assert rel(y, g["out"]) < tol, describe(y, g["out"], f"{case} {dtype}")


@pytest.mark.parametrize("dtype,tol", [(torch.bfloat16, TOL_BF16), (torch.float32, TOL_F32)])
def test_forward_at_bench_batch(dtype, tol):
    """base40M-imagevec at B = 32 (M = 32 832 rows, 768 grouped attention items: every persistent CTA of the GEMM and
    attention kernels walks several tiles per launch) against the unmodified reference, every sequence on its own."""
    case = "full_imagevec_b32"
    g = load_golden("forward_" + case)
    stride = cases.FORWARD_POINT_STRIDE[case]
    model, cfg, _ = build_model(case, dtype)
    x, t, kw = cases.forward_inputs(case)
    with torch.no_grad():
        y = model(x.to(DEV), t.to(DEV), **to_dev(kw))
    torch.cuda.synchronize()
    got, want = y[:, :, ::stride].cpu(), torch.from_numpy(g["out"])
    assert got.shape == want.shape
    assert rel(got, want) < tol, describe(got, want, f"{case} {dtype}")
    per_seq = (got - want).flatten(1).norm(dim=1) / want.flatten(1).norm(dim=1)
    assert float(per_seq.max()) < tol, f"worst sequence {int(per_seq.argmax())}: {float(per_seq.max()):.3e}"
    if dtype == torch.bfloat16:
        # the sampler's shape: the same 32 clouds as ONE 64-sequence guided evaluation (conditional + unconditional
        # halves sharing x, common timestep) -- the conditional half must equal a plain forward at that timestep
        tt = torch.full((32,), 511, device=DEV)
        with torch.no_grad():
            plain = model(x.to(DEV), tt, **to_dev(kw)).clone()
            e = kw["embeddings"].to(DEV)
            both = model.forward_cfg(x.to(DEV), 511, dict(embeddings=torch.cat([e, torch.zeros_like(e)])), doubled=True)
        torch.cuda.synchronize()
        assert rel(both[:32], plain) < 1e-3, describe(both[:32], plain, "guided batch vs plain forward")


@pytest.mark.parametrize("case", ["full_imagevec", "full_base300M"])
def test_forward_layernorm_folded_vs_separate(case):
    """The LayerNorm-folded bf16 forward (no LayerNorm kernels; residual update, bf16 copy and row
    statistics in the c_proj epilogues) against the same forward with separate LayerNorm kernels
    (per-handle option PCD_MODEL_SEPARATE_LAYERNORM), and both against the reference golden."""
    import pcd_b200 as P
    lib = P._lib.load()
    g = load_golden("forward_" + case)
    model, cfg, _ = build_model(case, torch.bfloat16)
    x, t, kw = cases.forward_inputs(case)
    n0 = lib.pcd_launch_count()
    with torch.no_grad():
        y_fold = model(x.to(DEV), t.to(DEV), **to_dev(kw)).clone()
    n_fold = lib.pcd_launch_count() - n0
    model.fold_layernorm = False   # per-handle option (pcd_model_desc.flags): separate LayerNorm kernels
    n0 = lib.pcd_launch_count()
    with torch.no_grad():
        y_sep = model(x.to(DEV), t.to(DEV), **to_dev(kw)).clone()
    n_sep = lib.pcd_launch_count() - n0
    torch.cuda.synchronize()
    layers = cfg["layers"]
    assert n_sep - n_fold == 2 * layers, (n_sep, n_fold)   # 2 LayerNorm launches per block; embed_tokens emits the first copy
    assert rel(y_fold, g["out"]) < TOL_BF16 and rel(y_sep, g["out"]) < TOL_BF16
    assert rel(y_fold, y_sep) < 1e-2, describe(y_fold, y_sep, case)
    print(f"{case}: folded vs golden {rel(y_fold, g['out']):.2e}, separate vs golden {rel(y_sep, g['out']):.2e}")


def test_forward_width_2048_layernorm_folded():
    """base1B width (2048 = 32 heads, 16 statistics slots per row) on the LayerNorm-folded path against
    the separate-LayerNorm path and the fp32 parity mode (no golden at this width: 2 layers, short sequence)."""
    import pcd_b200 as P
    lib = P._lib.load()
    cfg = dict(P.MODEL_CONFIGS["base40M-uncond"], width=2048, heads=32, layers=2, n_ctx=300)
    torch.manual_seed(3)
    m16 = P.model_from_config(cfg, DEV, dtype=torch.bfloat16)
    with torch.no_grad():
        m16.output_proj.weight.normal_(std=0.02)
    m32 = P.model_from_config(cfg, DEV, dtype=torch.float32)
    m32.load_state_dict(m16.state_dict())
    x = torch.randn(2, cfg["input_channels"], 300, device=DEV)
    t = torch.tensor([900, 17], device=DEV)
    with torch.no_grad():
        y_fold = m16(x, t).clone()
        m16.fold_layernorm = False
        y_sep = m16(x, t).clone()
        y32 = m32(x, t)
    torch.cuda.synchronize()
    assert rel(y_fold, y32) < TOL_BF16 and rel(y_sep, y32) < TOL_BF16, (rel(y_fold, y32), rel(y_sep, y32))
    assert rel(y_fold, y_sep) < 1e-2


@pytest.mark.parametrize("name", ["base40M-imagevec", "base40M-textvec", "base40M-uncond", "base40M", "base300M",
                                  "base1B", "upsample"])
def test_every_registered_config_runs_at_full_size(name):
    """Each entry of the reference's MODEL_CONFIGS (models/configs.py:15-114) at its real shape, B = 1:
    the bf16 tensor-core forward against the fp32 parity-mode forward of the same weights."""
    import pcd_b200 as P
    cfg = P.MODEL_CONFIGS[name]
    torch.manual_seed(11)
    m16 = P.model_from_config(cfg, DEV, dtype=torch.bfloat16)
    with torch.no_grad():
        m16.output_proj.weight.normal_(std=0.02)
    m32 = P.model_from_config(cfg, DEV, dtype=torch.float32)
    m32.load_state_dict(m16.state_dict())
    for m in (m16, m32):
        if hasattr(m, "accept_grid_embeddings"):
            m.accept_grid_embeddings = True
    cls = cfg["name"]
    kw = {}
    if cls == "CLIPImagePointDiffusionTransformer":
        e = torch.randn(1, 768, device=DEV)
        kw["embeddings"] = e / e.norm(dim=1, keepdim=True)
    if "Grid" in cls:
        kw["embeddings"] = torch.randn(1, 1024, 256, device=DEV)
    if "Upsample" in cls:
        lr = torch.rand(1, cfg["input_channels"], cfg["cond_ctx"], device=DEV) - 0.5
        lr[:, 3:] = (lr[:, 3:] + 0.5) * 255.0
        kw["low_res"] = lr
    x = torch.randn(1, cfg["input_channels"], cfg["n_ctx"], device=DEV)
    t = torch.tensor([500], device=DEV)
    with torch.no_grad():
        y16 = m16(x, t, **kw)
        y32 = m32(x, t, **kw)
    torch.cuda.synchronize()
    assert y16.shape == (1, cfg["output_channels"], cfg["n_ctx"]) and torch.isfinite(y16).all()
    assert rel(y16, y32) < TOL_BF16, (name, rel(y16, y32))


def test_forward_cfg_shares_x():
    """2B-sequence CFG forward (cond rows then uncond rows sharing x) == two B-sized calls."""
    model, cfg, _ = build_model("small_imagevec", torch.float32)
    x, t, kw = cases.forward_inputs("small_imagevec")
    B = x.shape[0]
    emb = kw["embeddings"].to(DEV)
    kw2 = dict(embeddings=torch.cat([emb, torch.zeros_like(emb)], 0))
    with torch.no_grad():
        both = model.forward_cfg(x.to(DEV), 511, kw2, doubled=True).clone()
        tt = torch.full((B,), 511, device=DEV)
        c = model(x.to(DEV), tt, embeddings=emb)
        u = model(x.to(DEV), tt, embeddings=torch.zeros_like(emb))
    assert rel(both[:B], c) < 1e-6 and rel(both[B:], u) < 1e-6
    with torch.no_grad():
        eps_only = model.forward_cfg(x.to(DEV), 511, kw2, doubled=True, out_channels=6)
    assert eps_only.shape[1] == 6 and rel(eps_only, both[:, :6]) < 1e-6


def test_reference_state_dict_roundtrip():
    model, cfg, sd = build_model("small_grid", torch.bfloat16)
    out = model.state_dict()
    for k, v in sd.items():
        assert torch.equal(out[k].cpu(), v), k
