"""Sampler parity: fused Heun kernels + denoiser vs the reference's own sampling loop."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from gpu_util import DEV, build_model, describe, rel, to_dev
from oracle import cases
from oracle import denoiser as D
from oracle import sampler as S

import pcd_b200 as P

pytestmark = pytest.mark.gpu


def make_sampler(case, dtype, graph):
    sc = cases.SAMPLER_CASES[case]
    model, cfg, sd = build_model(sc["model"], dtype)
    dname = "upsample" if sc["diffusion"] == "upsample" else "base40M"
    diffusion = P.diffusion_from_config(P.DIFFUSION_CONFIGS[dname])
    C = cfg["input_channels"]
    noise = cases.DetNoise(sc["noise_seed"])
    sampler = P.PointCloudSampler(
        device=DEV, models=[model], diffusions=[diffusion], num_points=[cfg["n_ctx"]],
        aux_channels=["R", "G", "B"][: C - 3], guidance_scale=[sc["guidance"]], use_karras=[True],
        karras_steps=[sc["steps"]], sigma_min=[sc["sigma_min"]], sigma_max=[sc["sigma_max"]],
        s_churn=[sc["s_churn"]], use_cuda_graph=graph, noise_fn=lambda shp: noise(shp).to(DEV))
    return sampler, sc, cfg, sd


def test_sampler_kernels_one_step_vs_oracle_math():
    """One guided Heun step of the fused kernels against the reference formulas on the CPU."""
    B, Cc, N, Co = 3, 6, 64, 12
    from oracle import det
    diffusion = P.diffusion_from_config(P.DIFFUSION_CONFIGS["base40M"])
    plan = P.HeunPlan(diffusion, 8, 1e-3, 120.0, 7.0, 3.0)
    HeunState = P.k_diffusion.HeunState
    st = HeunState(diffusion, plan, (B, Cc, N), DEV, 3.0, True)
    xT = det.normal((B, Cc, N), 801) * 120.0
    n0, n1 = det.normal((B, Cc, N), 802), det.normal((B, Cc, N), 803)
    o1, o2 = det.normal((2 * B, Co, N), 804), det.normal((2 * B, Co, N), 805)
    st.x.copy_(xT.to(DEV))
    st.begin(n0.to(DEV))
    s0 = plan.steps[0]
    x_hat = xT + n0 * np.float32(s0.noise_scale)
    assert rel(st.x, x_hat) < 1e-7 and rel(st.model_in, x_hat * np.float32(s0.first.c_in)) < 1e-7
    pred = torch.empty(B, Cc, N, device=DEV)
    st.predictor(0, o1.to(DEV), pred)

    tab = S.Tables(**cases.DIFFUSION_CONFIGS["base"])

    def denoise(x, ev, out):
        x_in = x * np.float32(ev.c_in)
        t = torch.full((B,), ev.t, dtype=torch.long)
        c = tab.pred_xstart(out[:B], x_in, t)
        u = tab.pred_xstart(out[B:], x_in, t)
        return u + 3.0 * (c - u)

    den = denoise(x_hat, s0.first, o1)
    d = (x_hat - den) / np.float32(s0.sigma_hat)
    x2 = x_hat + d * np.float32(s0.dt)
    assert rel(pred, tab.unscale(den)) < 1e-6, describe(pred, tab.unscale(den), "pred_xstart")
    assert rel(st.d, d) < 1e-6 and rel(st.model_in, x2 * np.float32(s0.second.c_in)) < 1e-6
    st.corrector(0, o2.to(DEV), n1.to(DEV))
    den2 = denoise(x2, s0.second, o2)
    d2 = (x2 - den2) / np.float32(s0.second.sigma)
    xn = x_hat + (d + d2) / 2 * np.float32(s0.dt)
    s1 = plan.steps[1]
    xn = xn + n1 * np.float32(s1.noise_scale)
    assert rel(st.x, xn) < 1e-6, describe(st.x, xn, "x after corrector")
    assert rel(st.model_in, xn * np.float32(s1.first.c_in)) < 1e-6


@pytest.mark.parametrize("case", [c for c in cases.SAMPLER_CASES if c.startswith("small")])
@pytest.mark.parametrize("graph", [False, True])
def test_sampler_small_fp32_matches_reference(case, graph):
    g = load_golden("sampler_" + case)
    sampler, sc, cfg, _ = make_sampler(case, torch.float32, graph)
    kw = to_dev(cases.sampler_kwargs(case))
    ys = [y.clone() for y in sampler.sample_batch_progressive(sc["B"], kw)]
    assert len(ys) == int(g["n_yields"])
    ys = torch.stack(ys).cpu()[torch.as_tensor(g["yield_index"])]
    want = torch.from_numpy(g["yields"])
    assert rel(ys[0], want[0]) < 1e-4, describe(ys[0], want[0], "first yield")
    # trajectories amplify rounding differences; the final cloud is what matters
    cd = D.chamfer_distance_xyz(ys[-1], want[-1])
    assert float(cd.max()) < 1e-3, f"final-cloud chamfer {cd.tolist()}"
    assert rel(ys[-1], want[-1]) < 5e-2, describe(ys[-1], want[-1], "final yield")


@pytest.mark.parametrize("case", ["small_imagevec_guided", "small_upsample_unguided"])
def test_sampler_small_bf16_chamfer(case):
    g = load_golden("sampler_" + case)
    sampler, sc, cfg, _ = make_sampler(case, torch.bfloat16, True)
    kw = to_dev(cases.sampler_kwargs(case))
    final = sampler.sample_batch(sc["B"], kw).cpu()
    want = torch.from_numpy(g["yields"])[-1]
    cd = D.chamfer_distance_xyz(final, want)
    print("bf16 final chamfer", cd.tolist(), "rel", rel(final, want))
    assert float(cd.max()) < 1e-3, f"final-cloud chamfer {cd.tolist()}"


def test_graph_equals_eager_bitwise():
    a, sc, _, _ = make_sampler("small_imagevec_guided", torch.bfloat16, False)
    b, _, _, _ = make_sampler("small_imagevec_guided", torch.bfloat16, True)
    kw = to_dev(cases.sampler_kwargs("small_imagevec_guided"))
    ya = a.sample_batch(sc["B"], kw)
    yb = b.sample_batch(sc["B"], kw)
    yb2 = b.sample_batch(sc["B"], kw)  # replay of the captured graph with fresh deterministic noise
    assert torch.equal(ya, yb), describe(yb, ya, "graph vs eager")


def test_graph_replay_follows_new_conditioning_and_new_weights():
    """A captured stage must not read stale addresses (round-1 advisor finding): (a) new `embeddings` tensors
    (token_cond=False -> the conditioning vector is added to every token) and (b) re-loaded weights both have to
    show up in the next sample_batch of a CUDA-graph sampler, bit for bit like a fresh eager sampler."""
    from oracle import det
    from test_oracle_golden import shapes_of
    case = "small_imagevec_guided"
    graph, sc, cfg, sd = make_sampler(case, torch.bfloat16, True)
    B = sc["B"]
    kw1 = to_dev(cases.sampler_kwargs(case))
    y1 = graph.sample_batch(B, kw1).clone()

    # (a) fresh conditioning tensor at a different address, old one released
    e2 = det.normal(tuple(kw1["embeddings"].shape), 4321)
    kw2 = dict(embeddings=(e2 / e2.norm(dim=1, keepdim=True)).to(DEV))
    del kw1
    junk = [torch.full((B, 768), 7.0, device=DEV) for _ in range(8)]  # recycle the freed blocks with other data
    def restart_noise():
        fresh = cases.DetNoise(sc["noise_seed"])
        graph.noise_fn = lambda shp: fresh(shp).to(DEV)

    restart_noise()
    y2 = graph.sample_batch(B, kw2).clone()
    eager, _, _, _ = make_sampler(case, torch.bfloat16, False)
    want2 = eager.sample_batch(B, kw2)
    assert torch.equal(y2, want2), describe(y2, want2, "graph replay with new embeddings")
    assert not torch.equal(y2, y1)

    # (b) new weights loaded into the same module: the packed copies move, the graph has to be re-captured
    cfgm = cases.FORWARD_CASES[sc["model"]][0]
    sd3 = det.fill_state_dict(shapes_of(cfgm), 977, mode="unit", width=cfgm["width"])
    graph.models[0].load_state_dict(sd3)
    restart_noise()
    y3 = graph.sample_batch(B, kw2).clone()
    eager3, _, _, _ = make_sampler(case, torch.bfloat16, False)
    eager3.models[0].load_state_dict(sd3)
    want3 = eager3.sample_batch(B, kw2)
    assert torch.equal(y3, want3), describe(y3, want3, "graph replay after load_state_dict")
    del junk


def test_two_stage_cascade():
    g = load_golden("sampler_two_stage")
    base, bcfg, _ = build_model("small_imagevec", torch.float32)
    from oracle import det
    from test_oracle_golden import shapes_of
    ups_cfg = cases.small_cfg("upsample", cond_ctx=64, n_ctx=128)
    ups = P.model_from_config(ups_cfg, DEV, dtype=torch.float32)
    ups.load_state_dict(det.fill_state_dict(shapes_of(ups_cfg), 109, mode="unit"))
    d1 = P.diffusion_from_config(P.DIFFUSION_CONFIGS["base40M"])
    d2 = P.diffusion_from_config(P.DIFFUSION_CONFIGS["upsample"])
    noise = cases.DetNoise(9010)
    sampler = P.PointCloudSampler(
        device=DEV, models=[base, ups], diffusions=[d1, d2], num_points=[64, 128],
        aux_channels=["R", "G", "B"], guidance_scale=[3.0, 0.0], karras_steps=[16, 16],
        model_kwargs_key_filter=("embeddings", ""), noise_fn=lambda shp: noise(shp).to(DEV))
    e = det.normal((2, 768), 1090)
    ys = [y.clone().cpu() for y in sampler.sample_batch_progressive(2, dict(embeddings=(e / e.norm(dim=1, keepdim=True)).to(DEV)))]
    assert len(ys) == 34 and ys[0].shape == (2, 6, 64) and ys[-1].shape == (2, 6, 192)
    first, second = torch.stack(ys[:17]), torch.stack(ys[17:])
    assert rel(first[0], g["first_stage"][0]) < 1e-4
    assert float(D.chamfer_distance_xyz(first[-1], torch.from_numpy(g["first_stage"][-1])).max()) < 1e-3
    assert float(D.chamfer_distance_xyz(second[-1], torch.from_numpy(g["second_stage"][-1])).max()) < 1e-3


@pytest.mark.parametrize("dtype", [torch.bfloat16])
def test_full_size_sampler_matches_reference(dtype):
    """North-star config 1/2 shape: base40M-imagevec, B=1, 1024 pts, 64 Heun steps, guidance 3,
    s_churn 3 -- per-step denoiser output <= 2e-2 rel-L2 and final-cloud Chamfer <= 1e-3."""
    g = load_golden("sampler_full_imagevec_guided")
    sampler, sc, cfg, _ = make_sampler("full_imagevec_guided", dtype, True)
    model = sampler.models[0]
    # per-step denoiser parity on the reference's own traced inputs
    for j in g["trace_index"]:
        x = torch.from_numpy(g[f"trace_x_{j}"]).to(DEV)
        want = g[f"trace_out_{j}"]
        t = torch.from_numpy(g["eval_t"][j]).to(DEV)
        kwj = cases.sampler_kwargs("full_imagevec_guided")
        emb = kwj["embeddings"].to(DEV)
        if j % 2 == 1:  # the reference alternates cond / uncond forwards (k_diffusion.py:194-203)
            emb = torch.zeros_like(emb)
        with torch.no_grad():
            y = model(x, t, embeddings=emb)
        assert rel(y, want) < 2e-2, describe(y, want, f"denoiser eval {j}")
    kw = to_dev(cases.sampler_kwargs("full_imagevec_guided"))
    ys = [y.clone().cpu() for y in sampler.sample_batch_progressive(sc["B"], kw)]
    assert len(ys) == 65
    ys = torch.stack(ys)[torch.as_tensor(g["yield_index"])]
    want = torch.from_numpy(g["yields"])
    cd = D.chamfer_distance_xyz(ys[-1], want[-1])
    print("full-size final chamfer", cd.tolist(), "first-yield rel", rel(ys[0], want[0]))
    assert rel(ys[0], want[0]) < 2e-2
    assert float(cd.max()) < 1e-3


@pytest.mark.parametrize("case", list(cases.SOLVER_CASES))
def test_other_solvers_match_reference(case):
    """karras_sample_progressive(sampler="dpm" | "ancestral") and the KarrasDenoiser (EDM)
    preconditioning on the fused kernels vs the reference's own loops (fp32 mode)."""
    g = load_golden("solver_" + case)
    sc = cases.SOLVER_CASES[case]
    model, cfg, _ = build_model(sc["model"], torch.float32)
    if sc["diffusion"] == "karras":
        diffusion = P.k_diffusion.KarrasDenoiser(sigma_data=0.5)
    else:
        diffusion = P.diffusion_from_config(P.DIFFUSION_CONFIGS["base40M"])
    noise = cases.DetNoise(sc["noise_seed"])
    shape = (sc["B"], cfg["input_channels"], cfg["n_ctx"])
    outs = list(P.karras_sample_progressive(
        diffusion, model, shape, sc["steps"], clip_denoised=True, model_kwargs={}, device=DEV, sigma_min=1e-3,
        sigma_max=sc["sigma_max"], sampler=sc["sampler"], s_churn=sc["s_churn"], guidance_scale=0.0,
        noise_fn=lambda shp: noise(shp).to(DEV)))
    key = "denoised" if sc["sampler"] == "dpm" else "pred_xstart"
    assert len(outs) == g["x"].shape[0]
    assert all(key in o for o in outs[:-1]) and "pred_xstart" in outs[-1]
    xs = torch.stack([o["x"] for o in outs]).cpu()
    preds = torch.stack([o.get(key, o.get("pred_xstart")) for o in outs]).cpu()
    assert rel(xs[:2], g["x"][:2]) < 1e-4, describe(xs[:2], g["x"][:2], "x first steps")
    assert rel(preds[:2], g["pred"][:2]) < 1e-4, describe(preds[:2], g["pred"][:2], "pred first steps")
    assert rel(preds[-1], g["pred"][-1]) < 2e-2, describe(preds[-1], g["pred"][-1], "final")


def test_full_size_two_stage_pipeline_to_point_clouds():
    """The reference's image2pointcloud pipeline at its real shapes (examples / run.py:114-161): base40M-imagevec
    (1024 points, guidance 3) -> upsample (4096 points), both stages replayed from CUDA graphs, then
    output_to_point_clouds -> PointCloud -> PLY.  Semantic checks: stage chaining (sampler.py:166-171: the low-res
    cloud is kept in front of the upsampled points), colour quantisation, and determinism under a fixed seed."""
    import io
    torch.manual_seed(5)
    base = P.model_from_config(P.MODEL_CONFIGS["base40M-imagevec"], DEV)
    up = P.model_from_config(P.MODEL_CONFIGS["upsample"], DEV)
    for m in (base, up):
        with torch.no_grad():
            m.output_proj.weight.normal_(std=0.02)
    sampler = P.PointCloudSampler(
        device=DEV, models=[base, up],
        diffusions=[P.diffusion_from_config(P.DIFFUSION_CONFIGS["base40M-imagevec"]),
                    P.diffusion_from_config(P.DIFFUSION_CONFIGS["upsample"])],
        num_points=[1024, 4096 - 1024], aux_channels=["R", "G", "B"], guidance_scale=[3.0, 0.0],
        model_kwargs_key_filter=("embeddings", ""), karras_steps=[16, 16], use_cuda_graph=True)
    e = torch.randn(2, 768, device=DEV)
    e = e / e.norm(dim=1, keepdim=True)

    def run():
        torch.manual_seed(99)
        outs = list(sampler.sample_batch_progressive(batch_size=2, model_kwargs=dict(embeddings=e)))
        return outs

    outs = run()
    assert len(outs) == 2 * 17  # 16 steps + final yield per stage
    first_stage_final, final = outs[16], outs[-1]
    assert first_stage_final.shape == (2, 6, 1024) and final.shape == (2, 6, 4096)
    assert torch.isfinite(final).all()
    assert torch.equal(final[:, :, :1024], first_stage_final), "stage 2 keeps the low-res cloud in front"
    again = run()[-1]
    assert torch.equal(again, final), "same seed, same graph replay -> bit-identical clouds"
    pcs = sampler.output_to_point_clouds(final)
    assert len(pcs) == 2 and pcs[0].coords.shape == (4096, 3) and set(pcs[0].channels) == {"R", "G", "B"}
    for ch in pcs[0].channels.values():
        assert ch.min() >= 0.0 and ch.max() <= 1.0 and np.allclose(ch * 255.0, np.round(ch * 255.0), atol=1e-4)
    f = io.BytesIO()
    pcs[0].write_ply(f)
    raw = f.getvalue()
    assert raw.startswith(b"ply\nformat binary_little_endian 1.0\nelement vertex 4096\n")
    assert len(raw) == raw.index(b"end_header\n") + 11 + 4096 * 15
    sub = pcs[0].farthest_point_sample(1024, init_idx=0)
    assert sub.coords.shape == (1024, 3) and len(np.unique(sub.coords, axis=0)) > 1000
