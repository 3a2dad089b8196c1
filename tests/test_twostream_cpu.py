"""TwoStreamDenoiser (reference models/model.py:437-547, models/modules.py): the oracle restatement against
golden outputs of the unmodified reference."""
import pytest
import torch

from conftest import load_golden
from oracle import twostream as OT
from oracle.make_golden_twostream import CASES, fill, inputs


def golden_state(name):
    g = load_golden("twostream_" + name)
    shapes = {k: tuple(int(x) for x in v.split(",")) if v else () for k, v in zip(g["shapes_keys"], g["shapes_vals"])}
    sd = fill(shapes, CASES[name]["seed"])
    return g, sd


@pytest.mark.parametrize("name", ["small", "config"])
def test_oracle_matches_reference(name):
    c = CASES[name]
    g, sd = golden_state(name)
    x, t, labels, views, prev = inputs(c)
    with torch.no_grad():
        y0, z0 = OT.twostream_forward(sd, c, x, t, labels, views)
        y1, z1 = OT.twostream_forward(sd, c, x, t, labels, views, prev_latent=prev)
        y2, z2 = OT.twostream_forward(sd, c, x, t, torch.zeros_like(labels), None, prev_latent=torch.from_numpy(g["z0"]))
    rel = lambda a, b: float((a - torch.from_numpy(b)).norm() / torch.from_numpy(b).norm())
    assert rel(y0, g["y0"]) < 2e-6 and rel(z0, g["z0"]) < 2e-6
    assert rel(y1, g["y1"]) < 2e-6 and rel(z1[:, ::8], g["z1"]) < 2e-6
    assert rel(y2, g["y2"]) < 2e-6 and rel(z2[:, ::8], g["z2"]) < 2e-6


def test_oracle_sampler_threads_the_latent_like_the_reference():
    """Guided Heun sampling with latent self-conditioning (k_diffusion.py:171-203): the oracle's loop against the
    reference's own PointCloudSampler run on the small TwoStream model."""
    from oracle import cases
    from oracle import sampler as S
    c = CASES["small"]
    g, sd = golden_state("small")
    want = torch.from_numpy(load_golden("twostream_sampler_small")["yields"])
    _, _, labels, views, _ = inputs(c)
    tab = S.Tables(schedule="linear", timesteps=1000)
    fn = lambda xx, tt, **k: OT.twostream_forward(sd, c, xx, tt, k.get("class_labels"), k.get("viewpoints"), k.get("prev_latent"))
    with torch.no_grad():
        ys = list(S.sample_batch_progressive([fn], [lambda b, k: k], [tab], [c["num_points"]], [], c["B"],
                                             dict(class_labels=labels, viewpoints=views), guidance_scale=[3.0], karras_steps=[6],
                                             sigma_min=[1e-3], sigma_max=[120], s_churn=[0.0], noise_fn=cases.DetNoise(777)))
    got = torch.stack(ys)
    assert got.shape == want.shape
    assert float((got - want).norm() / want.norm()) < 1e-4


def test_oracle_unguided_sampler_does_not_thread_the_latent():
    """guidance_scale 0: the reference's plain `denoiser` drops the returned latent (k_diffusion.py:150-166), so every
    evaluation starts from prev_latent=None -- pinned by the reference's own PointCloudSampler run."""
    from oracle import cases
    from oracle import sampler as S
    c = CASES["small"]
    g, sd = golden_state("small")
    want = torch.from_numpy(load_golden("twostream_sampler_small_unguided")["yields"])
    _, _, labels, views, _ = inputs(c)
    tab = S.Tables(schedule="linear", timesteps=1000)
    fn = lambda xx, tt, **k: OT.twostream_forward(sd, c, xx, tt, k.get("class_labels"), k.get("viewpoints"), k.get("prev_latent"))
    with torch.no_grad():
        ys = list(S.sample_batch_progressive([fn], [lambda b, k: k], [tab], [c["num_points"]], [], c["B"],
                                             dict(class_labels=labels, viewpoints=views), guidance_scale=[0.0], karras_steps=[6],
                                             sigma_min=[1e-3], sigma_max=[120], s_churn=[0.0], noise_fn=cases.DetNoise(777)))
    got = torch.stack(ys)
    assert got.shape == want.shape
    assert float((got - want).norm() / want.norm()) < 1e-4


def test_oracle_partial_cloud_and_depth_encoders_match_reference():
    """PartialPointCloudEncoder / DepthMapEncoder (model.py:262-434: nn.TransformerEncoder / Decoder stacks) on their
    own, inside the full four-modality forward, and with the depth map dropped."""
    from oracle.make_golden_twostream import extra_inputs
    c = CASES["full4"]
    g, sd = golden_state("full4")
    x, t, labels, views, prev = inputs(c)
    pcd, depth = extra_inputs(c)
    rel = lambda a, b: float((a - torch.from_numpy(b)).norm() / torch.from_numpy(b).norm())
    with torch.no_grad():
        assert rel(OT.partial_pcd_encoder(sd, pcd), g["tok_p"]) < 5e-6
        assert rel(OT.depth_encoder(sd, depth), g["tok_d"]) < 5e-6
        y0, z0 = OT.twostream_forward(sd, c, x, t, labels, views, prev_latent=prev, partial_pcd=pcd, depth_maps=depth)
        y1, z1 = OT.twostream_forward(sd, c, x, t, labels, views, partial_pcd=pcd, depth_maps=torch.zeros_like(depth))
    assert rel(y0, g["y0"]) < 5e-6 and rel(z0[:, ::4], g["z0"]) < 5e-6
    assert rel(y1, g["y1"]) < 5e-6 and rel(z1[:, ::4], g["z1"]) < 5e-6
    assert rel(OT.sincos_2d(16, 16, c["latent_dim"])[::5, ::3], g["pos_embed"]) < 1e-6


@pytest.mark.parametrize("name", list(CASES))
def test_product_state_dict_has_the_reference_keys_and_shapes(name):
    """Drop-in weights: the product's parameter containers (no GPU needed to build them) expose exactly the keys and
    shapes of the reference's TwoStreamDenoiser.state_dict() recorded in the goldens."""
    import pcd_b200 as P
    from oracle.make_golden_twostream import ctor_kwargs
    g = load_golden("twostream_" + name)
    want = {k: tuple(int(x) for x in v.split(",")) if v else () for k, v in zip(g["shapes_keys"], g["shapes_vals"])}
    model = P.TwoStreamDenoiser(**ctor_kwargs(CASES[name]))
    got = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    assert got == want
    if "depth" in CASES[name]["active_modalities"]:  # the fixed positional table is built like the reference's
        pe = model.encoders["depth"].pos_embed
        assert float((pe[::5, ::3] - torch.from_numpy(g["pos_embed"])).abs().max()) < 1e-6
