"""Ancestral (DDPM) sampling -- GaussianDiffusion.p_mean_variance / p_sample / p_sample_loop_progressive and the
use_karras=False branch of PointCloudSampler (reference gaussian_diffusion.py:257-350,407-548, sampler.py:153-165)
-- on the fused pcd_ddpm_step kernel, against goldens of the unmodified reference and the oracle's formulas."""
import pytest
import torch

import pcd_b200 as P
from conftest import load_golden
from gpu_util import DEV, TOL_F32, build_model, describe, rel, to_dev
from oracle import cases, det
from oracle import sampler as S
from test_oracle_golden import ddpm_tables

pytestmark = pytest.mark.gpu


def make_diffusion(dc):
    scales = dict(channel_scales=cases._SCALES, channel_biases=cases._BIASES) if dc["scaled"] else {}
    return P.GaussianDiffusion(betas=P.get_named_beta_schedule(dc["schedule"], dc["timesteps"]), model_mean_type="epsilon",
                               model_var_type=dc["var_type"], loss_type="mse", **scales)


@pytest.mark.parametrize("var_type", ["fixed_small", "fixed_large", "learned_range", "learned"])
@pytest.mark.parametrize("n_points", [64, 30])  # vectorised and scalar kernels
def test_ddpm_step_kernel_vs_oracle_formulas(var_type, n_points):
    """One fused step with a mixed batch of step indices (t = 0 has no noise), arbitrary model output."""
    B, C, T = 5, 6, 50
    dc = dict(schedule="cosine", timesteps=T, scaled=True, var_type=var_type)
    diffusion, tab = make_diffusion(dc), ddpm_tables(dc)
    x = det.normal((B, C, n_points), 4101)
    out = det.normal((B, C if var_type.startswith("fixed") else 2 * C, n_points), 4102, std=0.7)
    noise = det.normal((B, C, n_points), 4103)
    t = torch.tensor([T - 1, 0, 17, 1, 0])
    model = lambda x_, t_, **kw: out.to(DEV)
    for clip in (True, False):
        want = S.p_mean_variance(tab, out, x, t, var_type, clip)
        nz = (t != 0).float().view(-1, 1, 1)
        want_sample = want["mean"] + nz * torch.exp(0.5 * want["log_variance"]) * noise
        got = diffusion.p_mean_variance(model, x.to(DEV), t.to(DEV), clip_denoised=clip)
        for k in ("mean", "variance", "log_variance", "pred_xstart"):
            assert rel(got[k], want[k]) < 1e-6, describe(got[k], want[k], f"{var_type} clip={clip} {k}")
        s = diffusion.p_sample(model, x.to(DEV), t.to(DEV), clip_denoised=clip, noise=noise.to(DEV))
        assert rel(s["sample"], want_sample) < 1e-6, describe(s["sample"], want_sample, "sample")
        assert rel(s["pred_xstart"], want["pred_xstart"]) < 1e-6
        assert torch.equal(s["sample"][1], got["mean"][1])  # t == 0: the mean itself


@pytest.mark.parametrize("case", list(cases.DDPM_CASES))
def test_ddpm_loop_matches_reference(case):
    """All T ancestral steps with the product denoiser (fp32 parity mode) against the reference's trajectory."""
    g = load_golden("ddpm_" + case)
    dc = cases.DDPM_CASES[case]
    model, cfg, _ = build_model(dc["model"], torch.float32)
    diffusion = make_diffusion(dc)
    C, N, B = cfg["input_channels"], cfg["n_ctx"], dc["B"]
    kw = to_dev(cases.ddpm_kwargs(case))
    noise = cases.DetNoise(dc["noise_seed"])
    draw = lambda shp: noise(shp).to(DEV)
    # denoised_fn / cond_fn cases (:321-326,374-385,433-436): the fused step is split around the user callables
    hooks = dict(zip(("denoised_fn", "cond_fn"), cases.ddpm_hooks())) if dc.get("hooks") else {}
    if dc["via"] == "sampler":
        sampler = P.PointCloudSampler(device=DEV, models=[model], diffusions=[diffusion], num_points=[N],
                                      aux_channels=["R", "G", "B"][: C - 3], guidance_scale=[0.0], use_karras=[False],
                                      karras_steps=[64], sigma_min=[1e-3], sigma_max=[120.0], s_churn=[0.0], noise_fn=draw)
        preds = torch.stack([y.clone() for y in sampler.sample_batch_progressive(B, kw)])
    else:
        outs = list(diffusion.p_sample_loop_progressive(model, (B, C, N), clip_denoised=True, model_kwargs=kw, device=DEV,
                                                        noise_fn=draw, **hooks))
        preds = torch.stack([o["pred_xstart"] for o in outs])
        samples = torch.stack([o["sample"] for o in outs])
        assert rel(samples, g["sample"]) < 1e-3, describe(samples, g["sample"], "samples")
        noise.__init__(dc["noise_seed"])
        final = diffusion.p_sample_loop(model, (B, C, N), clip_denoised=True, model_kwargs=kw, device=DEV, noise_fn=draw,
                                        **hooks)
        assert torch.equal(final, samples[-1])
        x = det.normal((B, C, N), dc["noise_seed"] + 5).to(DEV)
        t = torch.tensor([(dc["timesteps"] - 1, 0, dc["timesteps"] // 2)[i % 3] for i in range(B)], device=DEV)
        r = diffusion.p_mean_variance(model, x, t, clip_denoised=False, model_kwargs=kw,
                                      denoised_fn=hooks.get("denoised_fn"))
        for k, gk in (("mean", "pmv_mean"), ("variance", "pmv_variance"), ("log_variance", "pmv_log_variance"),
                      ("pred_xstart", "pmv_pred")):
            assert rel(r[k], g[gk]) < TOL_F32, describe(r[k], g[gk], k)
    assert preds.shape == g["pred"].shape
    assert rel(preds[:2], g["pred"][:2]) < TOL_F32, describe(preds[:2], g["pred"][:2], "first steps")
    assert rel(preds, g["pred"]) < 1e-3, describe(preds, g["pred"], "all steps")  # chaotic in the last ulp over T steps


def test_ddpm_guided_branch_is_refused_like_the_reference_fails():
    model, cfg, _ = build_model("small_imagevec", torch.float32)
    diffusion = P.diffusion_from_config(P.DIFFUSION_CONFIGS["base40M-imagevec"])
    sampler = P.PointCloudSampler(device=DEV, models=[model], diffusions=[diffusion], num_points=[cfg["n_ctx"]],
                                  aux_channels=["R", "G", "B"], guidance_scale=[3.0], use_karras=[False], karras_steps=[64],
                                  sigma_min=[1e-3], sigma_max=[120.0], s_churn=[0.0])
    with pytest.raises(NotImplementedError):
        next(iter(sampler.sample_batch_progressive(2, to_dev(cases.ddpm_kwargs("learned_range_cosine")))))
