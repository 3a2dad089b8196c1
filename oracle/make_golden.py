"""Generate tests/golden/*.npz by running the UNMODIFIED reference on CPU.

Run in the authoring container only (needs /root/reference):
    python -m oracle.make_golden [--full] [--only NAME]
Inputs/weights come from hash seeds (oracle/cases.py, oracle/det.py); only the
reference's OUTPUTS are stored.  TEST INFRASTRUCTURE (see oracle/__init__.py).
"""
import argparse
import contextlib
import os
import time

import numpy as np
import torch

from . import cases, det
from .ref_harness import ref_modules

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                          "tests", "golden")


def save(name, **arrays):
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    path = os.path.join(GOLDEN_DIR, name + ".npz")
    np.savez_compressed(path, **{k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v))
                                 for k, v in arrays.items()})
    print(f"  wrote {path} ({os.path.getsize(path) / 1024:.1f} KiB)")


def build_reference_model(m, case):
    cfg, B, seed, mode = cases.FORWARD_CASES[case]
    model = m.model_configs.model_from_config(cases.model_ctor_cfg(cfg), torch.device("cpu")).eval()
    sd = det.fill_state_dict(model.state_dict(), seed, mode=mode, width=cfg["width"])
    model.load_state_dict(sd)
    return model, cfg, sd


@contextlib.contextmanager
def patched_noise(noise):
    """Route the sampling loop's draws (k_diffusion.py:139,292) to ``noise``."""
    o_randn, o_like = torch.randn, torch.randn_like
    torch.randn = lambda *shape, **kw: noise(shape[0] if len(shape) == 1 and not isinstance(shape[0], int) else shape)
    torch.randn_like = lambda x, **kw: noise(tuple(x.shape))
    try:
        yield
    finally:
        torch.randn, torch.randn_like = o_randn, o_like


def gen_ops(m):
    print("ops")
    tr, pc, ro, ut = m.transformer, m.perceiver, m.rotary, m.util
    out = {}
    t_int = torch.tensor([0, 1, 17, 511, 1017, 1023], dtype=torch.long)
    t_flt = torch.tensor([0.5, 250.25, 999.75], dtype=torch.float32)
    for d in (128, 512):
        out[f"temb_int_{d}"] = ut.timestep_embedding(t_int, d)
        out[f"temb_flt_{d}"] = ut.timestep_embedding(t_flt, d)
    # self attention, reference QKVMultiheadAttention (transformer.py:65-84)
    qkv = det.normal((2, 70, 2 * 3 * 64), 301)
    att = tr.QKVMultiheadAttention(device="cpu", dtype=torch.float32, heads=2, n_ctx=70)
    out["self_attn"] = att(qkv)
    qkv8 = det.normal((1, 300, 8 * 3 * 64), 302, std=2.0)
    out["self_attn_h8"] = tr.QKVMultiheadAttention(device="cpu", dtype=torch.float32, heads=8, n_ctx=300)(qkv8)
    # cross attention, reference QKVMultiheadCrossAttention (perceiver.py:46-67)
    q = det.normal((2, 70, 128), 303)
    kv = det.normal((2, 77, 256), 304)
    out["cross_attn"] = pc.QKVMultiheadCrossAttention(device="cpu", dtype=torch.float32, heads=2, n_data=77)(q, kv)
    # rotary (rotaryencoderpcd.py:6-27, 58-84)
    rq = det.normal((2, 2, 70, 64), 305)
    rk = det.normal((2, 2, 70, 64), 306)
    coords = det.uniform((2, 70, 3), 307, std=0.5 / 3 ** 0.5)
    a, b = ro.apply_rotary_pos_emb(rq, rk, coords)
    out["rope_q"], out["rope_k"] = a, b
    rsa = ro.RotarySelfAttention(128, heads=2).eval()
    sd = det.fill_state_dict(rsa.state_dict(), 308)
    rsa.load_state_dict(sd)
    with torch.no_grad():
        out["rotary_self_attn"] = rsa(det.normal((2, 70, 128), 309), coords)
    # perceiver (perceiver.py:107-146)
    per = pc.SimplePerceiver(device="cpu", dtype=torch.float32, n_data=77, width=128, layers=2,
                             heads=2, data_width=192).eval()
    sd = det.fill_state_dict(per.state_dict(), 310)
    per.load_state_dict(sd)
    with torch.no_grad():
        out["perceiver"] = per(det.normal((2, 70, 128), 311), det.normal((2, 77, 192), 312))
    # Chamfer parity metric (models/util.py:265-295)
    p1 = det.uniform((2, 6, 200), 313, 0.3)
    p2 = det.uniform((2, 3, 150), 314, 0.3)
    out["chamfer"] = ut.chamfer_distance_xyz(p1, p2)
    save("ops", **out)


def gen_forward(m, case):
    print("forward", case)
    model, cfg, _ = build_reference_model(m, case)
    x, t, kw = cases.forward_inputs(case)
    t0 = time.time()
    with torch.no_grad():
        y = model(x, t, **kw)
    print(f"  reference forward {time.time() - t0:.2f}s  out std {y.std():.4f}")
    stride = cases.FORWARD_POINT_STRIDE.get(case, 1)
    save("forward_" + case, out=y[:, :, ::stride])


def gen_perceiver_text(m):
    """SimplePerceiver at the BASELINE config-3 text-conditioning shape (perceiver.py:107-146)."""
    print("perceiver_text")
    c = cases.PERCEIVER_TEXT
    per = m.perceiver.SimplePerceiver(device="cpu", dtype=torch.float32, n_data=c["n_data"], width=c["width"],
                                      layers=c["layers"], heads=c["heads"], data_width=c["data_width"]).eval()
    per.load_state_dict(det.fill_state_dict(per.state_dict(), c["seed"]))
    x, data = cases.perceiver_text_inputs()
    with torch.no_grad():
        y = per(x, data)
    print("  out std", float(y.std()))
    save("perceiver_text", out=y[:, ::c["row_stride"]])


def gen_schedule(m):
    """Pin sigma schedule + sigma->t truncation (k_diffusion.py:89-103, 225-231)."""
    print("schedule")
    out = {}
    for name, (dcfg, smax, churn) in {"base": ("base", 120.0, 3.0), "upsample": ("upsample", 160.0, 0.0)}.items():
        diffusion = m.diffusion_configs.diffusion_from_config(cases.DIFFUSION_CONFIGS[dcfg])
        sig = m.k_diffusion.get_sigmas_karras(64, 1e-3, smax, 7.0)
        wrap = m.k_diffusion.GaussianToKarrasDenoiser(None, diffusion)
        gamma = min(churn / 64, 2 ** 0.5 - 1)
        evals = []
        for i in range(64):
            sh = sig[i] * (gamma + 1)
            evals.append(sh)
            if sig[i + 1] != 0:
                evals.append(sig[i + 1])
        ev = torch.stack(evals)
        t = torch.tensor([wrap.sigma_to_t(s) for s in ev.numpy()], dtype=torch.long)
        out[name + "_sigmas"] = sig
        out[name + "_eval_sigmas"] = ev
        out[name + "_eval_t"] = t
        out[name + "_alphas_cumprod"] = diffusion.alphas_cumprod
    save("schedule", **out)


def gen_sampler(m, case):
    print("sampler", case)
    sc = cases.SAMPLER_CASES[case]
    model, cfg, _ = build_reference_model(m, sc["model"])
    diffusion = m.diffusion_configs.diffusion_from_config(cases.DIFFUSION_CONFIGS[sc["diffusion"]])
    C = cfg["input_channels"]
    aux = ["R", "G", "B"][: C - 3]
    sampler = m.sampler.PointCloudSampler(
        device=torch.device("cpu"), models=[model], diffusions=[diffusion],
        num_points=[cfg["n_ctx"]], aux_channels=aux, guidance_scale=[sc["guidance"]],
        use_karras=[True], karras_steps=[sc["steps"]], sigma_min=[sc["sigma_min"]],
        sigma_max=[sc["sigma_max"]], s_churn=[sc["s_churn"]])
    kw = cases.sampler_kwargs(case)
    # trace the denoiser evaluations of the unmodified reference
    trace = []
    orig_forward = model.forward

    def traced(x, t, **k):
        y = orig_forward(x, t, **k)
        trace.append((t.clone(), x.clone(), y.clone()))
        return y

    model.forward = traced
    t0 = time.time()
    with patched_noise(cases.DetNoise(sc["noise_seed"])), torch.no_grad():
        yields = [y.clone() for y in sampler.sample_batch_progressive(sc["B"], kw)]
    print(f"  reference sampler {time.time() - t0:.1f}s, {len(yields)} yields, {len(trace)} forwards")
    ys = torch.stack(yields)
    keep = sorted(set([0, 1, 2, len(trace) // 2, len(trace) - 2, len(trace) - 1]))
    arrays = dict(yields=ys if ys.numel() < 2_000_000 else ys[[0, 1, 2, 16, 32, 48, 62, 63, 64]],
                  yield_index=np.arange(len(yields)) if ys.numel() < 2_000_000
                  else np.array([0, 1, 2, 16, 32, 48, 62, 63, 64]),
                  n_yields=len(yields), n_forwards=len(trace),
                  eval_t=torch.stack([tr[0] for tr in trace]),
                  trace_index=np.array(keep))
    for j in keep:
        arrays[f"trace_x_{j}"] = trace[j][1]
        arrays[f"trace_out_{j}"] = trace[j][2]
    save("sampler_" + case, **arrays)


def gen_two_stage(m):
    """Two-stage cascade through the reference PointCloudSampler defaults
    (sampler.py:34-40): base (guided, churned) -> upsampler (low_res chaining)."""
    print("sampler two_stage")
    base, bcfg, _ = build_reference_model(m, "small_imagevec")
    ups_cfg = cases.small_cfg("upsample", cond_ctx=64, n_ctx=128)
    ups = m.model_configs.model_from_config(ups_cfg, torch.device("cpu")).eval()
    ups.load_state_dict(det.fill_state_dict(ups.state_dict(), 109, mode="unit"))
    d1 = m.diffusion_configs.diffusion_from_config(cases.DIFFUSION_CONFIGS["base"])
    d2 = m.diffusion_configs.diffusion_from_config(cases.DIFFUSION_CONFIGS["upsample"])
    sampler = m.sampler.PointCloudSampler(
        device=torch.device("cpu"), models=[base, ups], diffusions=[d1, d2],
        num_points=[64, 128], aux_channels=["R", "G", "B"], guidance_scale=[3.0, 0.0],
        karras_steps=[16, 16], model_kwargs_key_filter=("embeddings", ""))
    e = det.normal((2, 768), 1090)
    kw = dict(embeddings=e / e.norm(dim=1, keepdim=True))
    with patched_noise(cases.DetNoise(9010)), torch.no_grad():
        yields = [y.clone() for y in sampler.sample_batch_progressive(2, kw)]
    print("  ", len(yields), yields[0].shape, yields[-1].shape)
    save("sampler_two_stage", first_stage=torch.stack(yields[:17]), second_stage=torch.stack(yields[17:]))


def gen_solver(m, case):
    """DPM-2 / Euler-ancestral / KarrasDenoiser through the reference's karras_sample_progressive."""
    print("solver", case)
    sc = cases.SOLVER_CASES[case]
    model, cfg, _ = build_reference_model(m, sc["model"])
    if sc["diffusion"] == "karras":
        diffusion = m.k_diffusion.KarrasDenoiser(sigma_data=0.5)
    else:
        diffusion = m.diffusion_configs.diffusion_from_config(cases.DIFFUSION_CONFIGS[sc["diffusion"]])
    shape = (sc["B"], cfg["input_channels"], cfg["n_ctx"])
    with patched_noise(cases.DetNoise(sc["noise_seed"])), torch.no_grad():
        outs = list(m.k_diffusion.karras_sample_progressive(
            diffusion, model, shape, sc["steps"], clip_denoised=True, model_kwargs={}, device=torch.device("cpu"),
            sigma_min=1e-3, sigma_max=sc["sigma_max"], sampler=sc["sampler"], s_churn=sc["s_churn"],
            guidance_scale=0.0))
    key = "denoised" if sc["sampler"] == "dpm" else "pred_xstart"
    xs = torch.stack([o["x"] for o in outs])
    preds = torch.stack([o.get(key, o.get("pred_xstart")) for o in outs])
    print("  ", len(outs), xs.shape, float(preds.std()))
    save("solver_" + case, x=xs, pred=preds)


def gen_ddpm(m, case):
    """Ancestral sampling over every diffusion step through the unmodified reference."""
    print("ddpm", case)
    dc = cases.DDPM_CASES[case]
    model, cfg, _ = build_reference_model(m, dc["model"])
    gd = m.gaussian_diffusion
    scales = dict(channel_scales=np.array(cases._SCALES), channel_biases=np.array(cases._BIASES)) if dc["scaled"] else {}
    diffusion = gd.GaussianDiffusion(betas=gd.get_named_beta_schedule(dc["schedule"], dc["timesteps"]),
                                     model_mean_type="epsilon", model_var_type=dc["var_type"], loss_type="mse", **scales)
    C, N, B = cfg["input_channels"], cfg["n_ctx"], dc["B"]
    kw = cases.ddpm_kwargs(case)
    hooks = dict(zip(("denoised_fn", "cond_fn"), cases.ddpm_hooks())) if dc.get("hooks") else {}
    with patched_noise(cases.DetNoise(dc["noise_seed"])), torch.no_grad():
        if dc["via"] == "sampler":
            sampler = m.sampler.PointCloudSampler(
                device=torch.device("cpu"), models=[model], diffusions=[diffusion], num_points=[N],
                aux_channels=["R", "G", "B"][: C - 3], guidance_scale=[0.0], use_karras=[False], karras_steps=[64],
                sigma_min=[1e-3], sigma_max=[120.0], s_churn=[0.0])
            preds = torch.stack([y.clone() for y in sampler.sample_batch_progressive(B, kw)])
            arrays = dict(pred=preds)
        else:
            outs = list(diffusion.p_sample_loop_progressive(model, (B, C, N), clip_denoised=True, model_kwargs=kw,
                                                            device=torch.device("cpu"), **hooks))
            arrays = dict(pred=torch.stack([o["pred_xstart"] for o in outs]),
                          sample=torch.stack([o["sample"] for o in outs]))
            # one p_mean_variance call (scaled units, no clipping) at a mixed batch of step indices
            x = det.normal((B, C, N), dc["noise_seed"] + 5)
            t = torch.tensor([(dc["timesteps"] - 1, 0, dc["timesteps"] // 2)[i % 3] for i in range(B)])
            pmv = diffusion.p_mean_variance(model, x, t, clip_denoised=False, model_kwargs=kw,
                                            denoised_fn=hooks.get("denoised_fn"))
            arrays.update(pmv_mean=pmv["mean"], pmv_log_variance=pmv["log_variance"].expand_as(x).clone(),
                          pmv_variance=pmv["variance"].expand_as(x).clone(), pmv_pred=pmv["pred_xstart"])
    print("  ", {k: tuple(v.shape) for k, v in arrays.items()}, float(arrays["pred"].std()))
    save("ddpm_" + case, **arrays)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--full", action="store_true", help="also run the full-size (slow) cases")
    ap.add_argument("--only", default=None)
    args = ap.parse_args()
    torch.set_num_threads(os.cpu_count())
    m = ref_modules()
    todo = []
    todo.append(("ops", lambda: gen_ops(m)))
    todo.append(("schedule", lambda: gen_schedule(m)))
    for c in cases.FORWARD_CASES:
        if c.startswith("full") and not args.full:
            continue
        todo.append(("forward_" + c, lambda c=c: gen_forward(m, c)))
    for c in cases.SAMPLER_CASES:
        if c.startswith("full") and not args.full:
            continue
        todo.append(("sampler_" + c, lambda c=c: gen_sampler(m, c)))
    todo.append(("sampler_two_stage", lambda: gen_two_stage(m)))
    todo.append(("perceiver_text", lambda: gen_perceiver_text(m)))
    for c in cases.SOLVER_CASES:
        todo.append(("solver_" + c, lambda c=c: gen_solver(m, c)))
    for c in cases.DDPM_CASES:
        todo.append(("ddpm_" + c, lambda c=c: gen_ddpm(m, c)))
    for name, fn in todo:
        if args.only and args.only != name:
            continue
        fn()


if __name__ == "__main__":
    main()
