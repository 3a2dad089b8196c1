"""The CPU oracle (oracle/) against golden vectors produced by the UNMODIFIED
reference (oracle/make_golden.py).  Pins the checker before it is trusted."""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_l2
from oracle import cases, det
from oracle import denoiser as D
from oracle import sampler as S

TOL = 2e-6  # same fp32 torch ops in the same order; allow for reassociation in BLAS


def shapes_of(cfg):
    """Reference-shaped state dict (keys/shapes only) from the config: the weights
    contract of SURVEY.md 8a."""
    d, L = cfg["width"], cfg["layers"]
    sd = {}

    def lin(name, o, i):
        sd[name + ".weight"] = torch.zeros(o, i)
        sd[name + ".bias"] = torch.zeros(o)

    def ln(name, n):
        sd[name + ".weight"] = torch.zeros(n)
        sd[name + ".bias"] = torch.zeros(n)

    lin("time_embed.c_fc", 4 * d, d)
    lin("time_embed.c_proj", d, 4 * d)
    ln("ln_pre", d)
    ln("ln_post", d)
    for i in range(L):
        p = f"backbone.resblocks.{i}"
        lin(p + ".attn.c_qkv", 3 * d, d)
        lin(p + ".attn.c_proj", d, d)
        ln(p + ".ln_1", d)
        lin(p + ".mlp.c_fc", 4 * d, d)
        lin(p + ".mlp.c_proj", d, 4 * d)
        ln(p + ".ln_2", d)
    lin("input_proj", d, cfg["input_channels"])
    lin("output_proj", cfg["output_channels"], d)
    name = cfg["name"]
    if name == "CLIPImagePointDiffusionTransformer":
        lin("clip_embed", d, 768)
    if "Grid" in name:
        ln("clip_embed.0", 1024)
        lin("clip_embed.1", d, 1024)
    if "Upsample" in name:
        lin("cond_point_proj", d, cfg["input_channels"])
        sd["channel_scales"] = torch.tensor(cfg["channel_scales"], dtype=torch.float32)
        sd["channel_biases"] = torch.tensor(cfg["channel_biases"], dtype=torch.float32)
    return sd


def case_weights(case):
    cfg, B, seed, mode = cases.FORWARD_CASES[case]
    return det.fill_state_dict(shapes_of(cfg), seed, mode=mode, width=cfg["width"]), cfg


def test_ops():
    g = load_golden("ops")
    t_int = torch.tensor([0, 1, 17, 511, 1017, 1023], dtype=torch.long)
    t_flt = torch.tensor([0.5, 250.25, 999.75], dtype=torch.float32)
    for d in (128, 512):
        assert np.array_equal(D.timestep_embedding(t_int, d).numpy(), g[f"temb_int_{d}"])
        assert np.array_equal(D.timestep_embedding(t_flt, d).numpy(), g[f"temb_flt_{d}"])
    assert rel_l2(D.qkv_attention(det.normal((2, 70, 384), 301), 2), g["self_attn"]) < TOL
    assert rel_l2(D.qkv_attention(det.normal((1, 300, 1536), 302, std=2.0), 8), g["self_attn_h8"]) < TOL
    assert rel_l2(D.qkv_cross_attention(det.normal((2, 70, 128), 303), det.normal((2, 77, 256), 304), 2),
                  g["cross_attn"]) < TOL
    coords = det.uniform((2, 70, 3), 307, std=0.5 / 3 ** 0.5)
    assert rel_l2(D.apply_rotary(det.normal((2, 2, 70, 64), 305), coords), g["rope_q"]) < TOL
    assert rel_l2(D.apply_rotary(det.normal((2, 2, 70, 64), 306), coords), g["rope_k"]) < TOL
    sd = det.fill_state_dict({"qkv.weight": torch.zeros(384, 128), "qkv.bias": torch.zeros(384),
                              "out_proj.weight": torch.zeros(128, 128), "out_proj.bias": torch.zeros(128)}, 308)
    assert rel_l2(D.rotary_self_attention(sd, 2, det.normal((2, 70, 128), 309), coords),
                  g["rotary_self_attn"]) < TOL
    assert rel_l2(D.chamfer_distance_xyz(det.uniform((2, 6, 200), 313, 0.3), det.uniform((2, 3, 150), 314, 0.3)),
                  g["chamfer"]) < TOL


def perceiver_shapes(width, layers, data_width):
    sd = {}
    for i in range(layers):
        p = f"resblocks.{i}"
        for n, (o, k) in {"attn.c_q": (width, width), "attn.c_kv": (2 * width, data_width),
                          "attn.c_proj": (width, width), "mlp.c_fc": (4 * width, width),
                          "mlp.c_proj": (width, 4 * width)}.items():
            sd[f"{p}.{n}.weight"] = torch.zeros(o, k)
            sd[f"{p}.{n}.bias"] = torch.zeros(o)
        for n, k in {"ln_1": width, "ln_2": data_width, "ln_3": width}.items():
            sd[f"{p}.{n}.weight"] = torch.zeros(k)
            sd[f"{p}.{n}.bias"] = torch.zeros(k)
    return sd


def test_perceiver():
    g = load_golden("ops")
    sd = det.fill_state_dict(perceiver_shapes(128, 2, 192), 310)
    y = D.perceiver_forward(sd, 2, 2, det.normal((2, 70, 128), 311), det.normal((2, 77, 192), 312))
    assert rel_l2(y, g["perceiver"]) < TOL


@pytest.mark.parametrize("case", [c for c in cases.FORWARD_CASES if c.startswith("small")])
def test_forward_small(case):
    g = load_golden("forward_" + case)
    sd, cfg = case_weights(case)
    x, t, kw = cases.forward_inputs(case)
    y = D.denoiser_forward(sd, cfg, x, t, **kw)
    assert y.shape == g["out"].shape
    assert rel_l2(y, g["out"]) < TOL


@pytest.mark.slow
@pytest.mark.parametrize("case", ["full_imagevec"])
def test_forward_full(case):
    g = load_golden("forward_" + case)
    sd, cfg = case_weights(case)
    x, t, kw = cases.forward_inputs(case)
    with torch.no_grad():
        y = D.denoiser_forward(sd, cfg, x, t, **kw)
    assert rel_l2(y, g["out"]) < 1e-5


def test_perceiver_text_shape():
    """BASELINE config 3 (ii): SimplePerceiver with 1026 query tokens of width 512 over 77 text tokens of width 768."""
    c = cases.PERCEIVER_TEXT
    g = load_golden("perceiver_text")
    sd = det.fill_state_dict(perceiver_shapes(c["width"], c["layers"], c["data_width"]), c["seed"])
    x, data = cases.perceiver_text_inputs()
    with torch.no_grad():
        y = D.perceiver_forward(sd, c["layers"], c["heads"], x, data)
    assert rel_l2(y[:, ::c["row_stride"]], g["out"]) < 1e-5


def test_schedule():
    g = load_golden("schedule")
    for name, dcfg, smax, churn in (("base", "base", 120.0, 3.0), ("upsample", "upsample", 160.0, 0.0)):
        tab = S.Tables(**cases.DIFFUSION_CONFIGS[dcfg])
        assert np.array_equal(tab.alphas_cumprod, g[name + "_alphas_cumprod"])
        sig = S.karras_sigmas(64, 1e-3, smax)
        assert np.array_equal(sig.numpy(), g[name + "_sigmas"])
        s2t = S.SigmaToT(tab)
        t = [s2t(s) for s in g[name + "_eval_sigmas"]]
        assert t == list(g[name + "_eval_t"])


@pytest.mark.parametrize("case", [c for c in cases.SAMPLER_CASES if c.startswith("small")])
def test_sampler_small(case):
    g = load_golden("sampler_" + case)
    sc = cases.SAMPLER_CASES[case]
    sd, cfg = case_weights(sc["model"])
    tab = S.Tables(**cases.DIFFUSION_CONFIGS[sc["diffusion"]])
    kw = cases.sampler_kwargs(case)
    C = cfg["input_channels"]
    cached = (lambda b, k: D.cached_model_kwargs(cfg, b, k)) if cfg["name"].startswith("CLIP") else None
    with torch.no_grad():
        ys = list(S.sample_batch_progressive(
            [S.make_model_fn(sd, cfg)], [cached], [tab], [cfg["n_ctx"]], ["R", "G", "B"][: C - 3],
            sc["B"], kw, guidance_scale=[sc["guidance"]], karras_steps=[sc["steps"]],
            sigma_min=[sc["sigma_min"]], sigma_max=[sc["sigma_max"]], s_churn=[sc["s_churn"]],
            noise_fn=cases.DetNoise(sc["noise_seed"])))
    assert len(ys) == int(g["n_yields"])
    ys = torch.stack(ys)[torch.as_tensor(g["yield_index"])]
    # trajectories are chaotic in the last ulp; compare early steps tightly, all loosely
    assert rel_l2(ys[:3], g["yields"][:3]) < 1e-5
    assert rel_l2(ys, g["yields"]) < 1e-3


@pytest.mark.parametrize("case", list(cases.SOLVER_CASES))
def test_solvers(case):
    """DPM-2, Euler-ancestral and KarrasDenoiser preconditioning (oracle vs the reference)."""
    g = load_golden("solver_" + case)
    sc = cases.SOLVER_CASES[case]
    sd, cfg = case_weights(sc["model"])
    diffusion = S.KarrasScalings(0.5) if sc["diffusion"] == "karras" else S.Tables(**cases.DIFFUSION_CONFIGS["base"])
    shape = (sc["B"], cfg["input_channels"], cfg["n_ctx"])
    with torch.no_grad():
        outs = list(S.karras_progressive(S.make_model_fn(sd, cfg), diffusion, shape, sc["steps"], sampler=sc["sampler"],
                                         sigma_max=sc["sigma_max"], s_churn=sc["s_churn"],
                                         noise_fn=cases.DetNoise(sc["noise_seed"])))
    key = "denoised" if sc["sampler"] == "dpm" else "pred_xstart"
    xs = torch.stack([o["x"] for o in outs])
    preds = torch.stack([o.get(key, o.get("pred_xstart")) for o in outs])
    assert xs.shape == g["x"].shape
    assert rel_l2(xs[:2], g["x"][:2]) < 1e-5 and rel_l2(preds[:2], g["pred"][:2]) < 1e-5
    assert rel_l2(preds, g["pred"]) < 1e-3


def ddpm_tables(dc):
    scales = dict(channel_scales=cases._SCALES, channel_biases=cases._BIASES) if dc["scaled"] else {}
    return S.Tables(schedule=dc["schedule"], timesteps=dc["timesteps"], **scales)


@pytest.mark.parametrize("case", list(cases.DDPM_CASES))
def test_ddpm_ancestral_loop(case):
    """p_mean_variance / p_sample / p_sample_loop_progressive (gaussian_diffusion.py:257-350,407-548): the oracle's
    loop against the reference's, driven through PointCloudSampler(use_karras=[False]) or directly."""
    g = load_golden("ddpm_" + case)
    dc = cases.DDPM_CASES[case]
    sd, cfg = case_weights(dc["model"])
    tab = ddpm_tables(dc)
    shape = (dc["B"], cfg["input_channels"], cfg["n_ctx"])
    fn = S.make_model_fn(sd, cfg)
    kw = cases.ddpm_kwargs(case)
    hooks = dict(zip(("denoised_fn", "cond_fn"), cases.ddpm_hooks())) if dc.get("hooks") else {}
    with torch.no_grad():
        outs = list(S.ddpm_progressive(fn, tab, shape, dc["var_type"], True, kw, cases.DetNoise(dc["noise_seed"]), **hooks))
    preds = torch.stack([o["pred_xstart"] for o in outs])
    assert preds.shape == g["pred"].shape
    assert rel_l2(preds[:2], g["pred"][:2]) < 1e-5 and rel_l2(preds, g["pred"]) < 1e-3
    if "sample" in g:
        assert rel_l2(torch.stack([o["sample"] for o in outs]), g["sample"]) < 1e-3
        x = det.normal(shape, dc["noise_seed"] + 5)
        t = torch.tensor([(dc["timesteps"] - 1, 0, dc["timesteps"] // 2)[i % 3] for i in range(dc["B"])])
        with torch.no_grad():
            r = S.p_mean_variance(tab, fn(x, t, **kw), x, t, dc["var_type"], False, hooks.get("denoised_fn"))
        for k in ("mean", "log_variance", "variance"):
            assert rel_l2(r[k], g["pmv_" + k]) < 1e-5, k
        assert rel_l2(r["pred_xstart"], g["pmv_pred"]) < 1e-5
