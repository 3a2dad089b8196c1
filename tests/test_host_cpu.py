"""CPU-only checks of the host logic and of the C-ABI surface (no compute calls)."""
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT, load_golden
from test_oracle_golden import shapes_of
from oracle import cases

import pcd_b200 as P


def test_library_loads_and_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "pcd_b200.h")).read()
    declared = set(re.findall(r"PCD_API\s+[\w\s\*]+?\b(pcd_\w+)\s*\(", header))
    assert len(declared) >= 19
    lib = P._lib.load()
    assert declared == set(P._lib.EXPORTED_SYMBOLS), declared ^ set(P._lib.EXPORTED_SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.pcd_abi_version() == P._lib.ABI_VERSION == 4


def test_product_never_imports_the_oracle():
    pkg_dir = os.path.dirname(P.__file__)
    for root, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(root, f)).read()
                assert "oracle" not in src.replace("# oracle", ""), f"{f} mentions the oracle"


def test_no_cpu_fallback():
    model = P.model_from_config(dict(P.MODEL_CONFIGS["base40M-uncond"], width=128, layers=1, heads=2, n_ctx=16),
                                torch.device("cpu"))
    if torch.cuda.is_available():
        pytest.skip("CPU-only assertion")
    with pytest.raises(P._lib.PcdError):
        model(torch.zeros(1, 6, 16), torch.zeros(1))
    with pytest.raises(P._lib.PcdError):
        P.ops.layernorm(torch.zeros(4, 128), torch.ones(128), torch.zeros(128))


@pytest.mark.parametrize("name", ["base", "upsample"])
def test_heun_plan_matches_reference_schedule(name):
    g = load_golden("schedule")
    smax, churn = (120.0, 3.0) if name == "base" else (160.0, 0.0)
    diffusion = P.diffusion_from_config(P.DIFFUSION_CONFIGS["base40M" if name == "base" else "upsample"])
    assert np.array_equal(diffusion.alphas_cumprod, g[name + "_alphas_cumprod"])
    plan = P.HeunPlan(diffusion, 64, 1e-3, smax, 7.0, churn)
    assert np.array_equal(plan.sigmas.numpy(), g[name + "_sigmas"])
    sig = []
    for s in plan.steps:
        sig.append(s.first.sigma)
        if s.second is not None:
            sig.append(s.second.sigma)
    assert np.array_equal(np.array(sig, dtype=np.float32), g[name + "_eval_sigmas"])
    assert plan.eval_timesteps() == list(g[name + "_eval_t"])
    assert plan.num_evals == 127
    # scalar identities of the reference arithmetic
    s0 = plan.steps[0]
    assert s0.noise_scale > 0 if churn else s0.noise_scale == 0
    assert plan.steps[-1].second is None and plan.steps[-1].dt == -plan.steps[-1].sigma_hat


@pytest.mark.parametrize("cfg_name", list(P.MODEL_CONFIGS) + ["upsample-plain"])
def test_state_dict_contract(cfg_name):
    if cfg_name == "base1B":
        pytest.skip("1B parameters: shape logic identical to base300M")
    cfg = cases.small_cfg(cfg_name) if cfg_name in cases.MODEL_CONFIGS else None
    assert cfg is not None
    want = shapes_of(cfg)
    model = P.model_from_config(cases.model_ctor_cfg(cfg), torch.device("cpu"))
    have = model.state_dict()
    assert set(have) == set(want), set(have) ^ set(want)
    for k in want:
        assert tuple(have[k].shape) == tuple(want[k].shape), k
    assert float(model.output_proj.weight.abs().sum()) == 0.0  # zero-init like the reference


def test_sampler_constructor_semantics():
    dev = torch.device("cpu")
    m = [object(), object()]
    d = [None, None]
    s = P.PointCloudSampler(dev, m, d, [1024, 3072], ["R", "G", "B"], guidance_scale=[3.0],
                            karras_steps=[64], sigma_min=[1e-3], sigma_max=[120], s_churn=[3],
                            use_karras=[True], model_kwargs_key_filter=["*"])
    assert list(s.guidance_scale) == [3.0, 1.0]  # "don't guide the upsamplers by default"
    assert list(s.karras_steps) == [64, 64] and list(s.s_churn) == [3, 3]
    assert s.num_stages == 2
    with pytest.raises(AssertionError):  # reference asserts per-stage list lengths (sampler.py:63-69)
        P.PointCloudSampler(dev, [object()], [None], [1024], [])
    out = torch.tensor([[[0.1, 0.2], [0.3, 0.4], [0.5, 0.6], [300.0, -5.0], [127.4, 127.6], [0.0, 255.0]]])
    pos, aux = s.split_model_output(out, rescale_colors=True)
    assert pos.shape == (1, 3, 2)
    assert torch.allclose(aux["R"], torch.tensor([[1.0, 0.0]]))
    assert torch.allclose(aux["G"], torch.tensor([[127.0 / 255, 128.0 / 255]]))
    c = P.PointCloudSampler.combine(s, s)
    assert c.num_stages == 4
    assert list(c.num_points) == [1024, 3072, 1024, 3072] and list(c.guidance_scale) == [3.0, 1.0, 3.0, 1.0]
    # like the reference, with_options forwards guidance_scale as given: callers pass per-stage lists
    w = s.with_options([2.0, 1.0], False, use_karras=[True, True], karras_steps=[8, 8], sigma_min=[1e-3, 1e-3],
                       sigma_max=[120, 160], s_churn=[0, 0])
    assert list(w.karras_steps) == [8, 8] and w.clip_denoised is False and w.models is s.models
    assert list(w.model_kwargs_key_filter) == list(s.model_kwargs_key_filter)
    clouds = s.output_to_point_clouds(out.repeat(2, 1, 1))
    assert len(clouds) == 2 and clouds[0].coords.shape == (2, 3) and set(clouds[0].channels) == {"R", "G", "B"}
    assert abs(float(clouds[1].channels["G"][1]) - 128.0 / 255) < 1e-6


def test_configs_mirror_reference_registry():
    for name, cfg in P.MODEL_CONFIGS.items():
        assert cfg["width"] == cfg["heads"] * 64
    assert P.MODEL_CONFIGS["upsample"]["n_ctx"] == 3072 and P.MODEL_CONFIGS["upsample"]["cond_ctx"] == 1024
    assert P.DIFFUSION_CONFIGS["upsample"]["schedule"] == "linear"
    assert P.DIFFUSION_CONFIGS["base40M"]["schedule"] == "cosine"
    d = P.diffusion_from_config(P.DIFFUSION_CONFIGS["base40M"])
    x = torch.randn(2, 6, 5)
    assert torch.allclose(d.unscale_channels(d.scale_channels(x)), x, atol=1e-5)


def test_layernorm_fold_algebra():
    """Host-side weight folding used by the LayerNorm-folded forward: rstd (x W'^T - mu colsum) + const
    equals LayerNorm -> Linear on the same (bf16-rounded) rows."""
    from importlib import import_module
    T = import_module(P.__name__ + ".transformer")
    g = torch.Generator().manual_seed(5)
    d, n, m = 256, 96, 40
    w = (torch.randn(n, d, generator=g) / d ** 0.5).bfloat16().float()
    b = torch.randn(n, generator=g) * 0.1
    gamma = 1 + 0.3 * torch.randn(d, generator=g)
    beta = 0.2 * torch.randn(d, generator=g)
    x = (torch.randn(m, d, generator=g) * 2 + torch.randn(m, 1, generator=g)).bfloat16().double()
    wf, colsum, const = T.fold_layernorm_into_linear(w, b, gamma, beta)
    assert wf.dtype == torch.bfloat16 and colsum.dtype == torch.float32 and const.dtype == torch.float32
    mu = x.mean(dim=1, keepdim=True)
    rstd = 1.0 / torch.sqrt(x.var(dim=1, unbiased=False, keepdim=True) + 1e-5)
    got = rstd * (x @ wf.double().t() - mu * colsum.double()[None, :]) + const.double()[None, :]
    want = torch.nn.functional.layer_norm(x, (d,), gamma.double(), beta.double(), 1e-5) @ w.double().t() + b.double()
    # the only difference is the bf16 rounding of gamma o W
    assert float((got - want).norm() / want.norm()) < 3e-3
    exact = rstd * (x @ (w.double() * gamma.double()[None, :]).t() - mu * (w.double() * gamma.double()[None, :]).sum(1)[None, :]) + const.double()
    assert float((exact - want).norm() / want.norm()) < 1e-6


def test_checkpoint_lookup_is_local_only(tmp_path):
    """Reference checkpoint names resolve to the files its downloader would have cached; nothing is fetched."""
    with pytest.raises(ValueError):
        P.download.load_checkpoint("nope", torch.device("cpu"), cache_dir=str(tmp_path))
    with pytest.raises(FileNotFoundError):
        P.download.load_checkpoint("base40M-uncond", torch.device("cpu"), cache_dir=str(tmp_path))
    cfg = dict(P.MODEL_CONFIGS["base40M-uncond"], width=128, layers=1, heads=2, n_ctx=16)
    model = P.model_from_config(cfg, torch.device("cpu"))
    torch.save(model.state_dict(), tmp_path / "base_40m_uncond.pt")
    sd = P.download.load_checkpoint("base40M-uncond", torch.device("cpu"), cache_dir=str(tmp_path))
    model2 = P.model_from_config(cfg, torch.device("cpu"))
    model2.load_state_dict(sd)
    assert all(torch.equal(a, b) for a, b in zip(model.state_dict().values(), model2.state_dict().values()))
