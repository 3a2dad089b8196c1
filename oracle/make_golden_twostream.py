"""Golden vectors of the unmodified reference's TwoStreamDenoiser (models/model.py, models/modules.py):

    python -m oracle.make_golden_twostream     # -> tests/golden/twostream_{small,config}.npz

Weights and inputs are regenerated from hash seeds (oracle/det.py); only the outputs are stored."""
import importlib
import os

import numpy as np
import torch

from oracle import det
from oracle.ref_harness import load_reference

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# "config": the shapes of the reference's config.yaml:24-38 restricted to the class / view modalities
CASES = {
    "small": dict(num_points=96, num_latents=24, latent_dim=64, x_dim=64, num_blocks=2, num_compute_layers=2,
                  num_heads=2, num_classes=16, active_modalities=["class", "view"], B=3, seed=1501),
    "config": dict(num_points=1024, num_latents=256, latent_dim=256, x_dim=256, num_blocks=6, num_compute_layers=4,
                   num_heads=8, num_classes=16, active_modalities=["class", "view"], B=2, seed=1502),
    # all four modalities: the partial-cloud and depth-map encoders always have 8 + 4 + 4 layers of 8 heads
    # (model.py:264-266,343-344), so latent_dim = 256 keeps their head dim at the config's 32; the backbone is shallow
    "full4": dict(num_points=128, num_latents=32, latent_dim=256, x_dim=256, num_blocks=1, num_compute_layers=1,
                  num_heads=8, num_classes=16, num_tokens_ppcd=16, num_tokens_depth=8,
                  active_modalities=["class", "view", "partial_pcd", "depth"], B=2, seed=1503, partial_points=80),
}


def ctor_kwargs(c):
    return {k: v for k, v in c.items() if k not in ("B", "seed", "partial_points")}


def n_cond(c):
    sizes = {"class": 1, "view": 1, "partial_pcd": c.get("num_tokens_ppcd", 64), "depth": c.get("num_tokens_depth", 32)}
    return sum(sizes[m] for m in c["active_modalities"])


def extra_inputs(c):
    """partial cloud [B, P, 3] in the dataset's range and depth maps [B, 1, 512, 512] (model.py:352: the positional
    table is built for 512 x 512 inputs)."""
    B, s = c["B"], c["seed"]
    return det.uniform((B, c["partial_points"], 3), s + 4, 0.5), det.uniform((B, 1, 512, 512), s + 5, 1.0).abs()


def inputs(c):
    B, N, s = c["B"], c["num_points"], c["seed"]
    x = det.normal((B, 3, N), s + 1)
    t = torch.tensor([(37 * i + 5) % 1000 for i in range(B)], dtype=torch.long)
    labels = torch.tensor([(3 * i + 1) % c["num_classes"] for i in range(B)], dtype=torch.long)
    views = det.uniform((B, 3), s + 2, 1.0)
    n_lat = c["num_latents"] + n_cond(c) + 1
    prev = det.normal((B, n_lat, c["latent_dim"]), s + 3, std=0.5)
    return x, t, labels, views, prev


def fill(shapes, seed):
    """Deterministic weights: N(0, 0.25/sqrt(fan_in)) matrices, small biases, LayerNorm near identity -- including
    ln_latent (zero-initialised in the reference, which would hide the self-conditioning path)."""
    sd = {}
    for i, (k, shp) in enumerate(shapes.items()):
        if k == "token_types_template":
            continue
        if k.endswith("weight") and len(shp) == 1:      # LayerNorm gain
            sd[k] = 1.0 + 0.1 * det.normal(shp, seed + 7 * i)
        elif k.endswith("bias"):
            sd[k] = 0.05 * det.normal(shp, seed + 7 * i)
        elif len(shp) >= 2:
            sd[k] = det.normal(shp, seed + 7 * i) * (1.0 / (shp[-1] ** 0.5))
        else:
            sd[k] = 0.02 * det.normal(shp, seed + 7 * i)
    return sd


def main():
    load_reference()
    model_mod = importlib.import_module("point_e.models.model")
    for name, c in CASES.items():
        torch.manual_seed(0)
        ref = model_mod.TwoStreamDenoiser(**ctor_kwargs(c)).eval()
        shapes = {k: tuple(v.shape) for k, v in ref.state_dict().items()}
        sd = fill(shapes, c["seed"])
        sd["token_types_template"] = ref.state_dict()["token_types_template"]
        ref.load_state_dict(sd)
        x, t, labels, views, prev = inputs(c)
        if "partial_pcd" in c["active_modalities"]:
            gen_full(ref, name, c, shapes)
            continue
        with torch.no_grad():
            y0, z0 = ref(x, t, class_labels=labels, viewpoints=views)
            y1, z1 = ref(x, t, class_labels=labels, viewpoints=views, prev_latent=prev)
            y2, z2 = ref(x, t, class_labels=torch.zeros_like(labels), viewpoints=None, prev_latent=z0)  # unconditional branch
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", f"twostream_{name}.npz"),
                            y0=y0.numpy(), z0=z0.numpy(), y1=y1.numpy(), y2=y2.numpy(),
                            # later latents: every 8th token row is enough to pin them (file size)
                            z1=z1[:, ::8].numpy(), z2=z2[:, ::8].numpy(),
                            shapes_keys=np.array(list(shapes.keys())),
                            shapes_vals=np.array([",".join(map(str, v)) for v in shapes.values()]))
        print(name, y0.shape, z0.shape, float(y0.abs().mean()), float(z1.abs().mean()), len(shapes), "tensors")


def gen_full(ref, name, c, shapes):
    """All four modalities: the two transformer encoders on their own, the full forward, and the forward with the
    depth map dropped (zero tokens + masked type embedding, model.py:503-507,527-535)."""
    x, t, labels, views, prev = inputs(c)
    pcd, depth = extra_inputs(c)
    with torch.no_grad():
        tok_p = ref.encoders["partial_pcd"](pcd)
        tok_d = ref.encoders["depth"](depth)
        y0, z0 = ref(x, t, class_labels=labels, viewpoints=views, partial_pcd=pcd, depth_maps=depth, prev_latent=prev)
        y1, z1 = ref(x, t, class_labels=labels, viewpoints=views, partial_pcd=pcd, depth_maps=torch.zeros_like(depth))
    torch.manual_seed(0)
    fresh = type(ref.encoders["depth"])(in_channels=1, embed_dim=c["latent_dim"], num_tokens=c["num_tokens_depth"])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", f"twostream_{name}.npz"),
                        tok_p=tok_p.numpy(), tok_d=tok_d.numpy(), y0=y0.numpy(), z0=z0[:, ::4].numpy(),
                        y1=y1.numpy(), z1=z1[:, ::4].numpy(), pos_embed=fresh.pos_embed[::5, ::3].numpy(),
                        shapes_keys=np.array(list(shapes.keys())),
                        shapes_vals=np.array([",".join(map(str, v)) for v in shapes.values()]))
    print(name, tok_p.shape, tok_d.shape, y0.shape, z0.shape, float(tok_p.abs().mean()), float(tok_d.abs().mean()),
          float(y0.abs().mean()), len(shapes), "tensors")


def gen_sampler(guidance=3.0, out_name="twostream_sampler_small.npz"):
    """The reference's own sampling setup for this model (run.py:119-141, config.yaml:40-58): guided Heun with the
    latent self-conditioning threaded by guided_denoiser, on the small case, deterministic noise.  ``guidance=0``
    pins the UNGUIDED path, where the reference's plain ``denoiser`` never threads the latent (k_diffusion.py:150-166)."""
    from oracle import cases
    from oracle.make_golden import patched_noise
    load_reference()
    model_mod = importlib.import_module("point_e.models.model")
    gd = importlib.import_module("point_e.diffusion.gaussian_diffusion")
    smp = importlib.import_module("point_e.diffusion.sampler")
    c = CASES["small"]
    ref = model_mod.TwoStreamDenoiser(**ctor_kwargs(c)).eval()
    shapes = {k: tuple(v.shape) for k, v in ref.state_dict().items()}
    sd = fill(shapes, c["seed"])
    sd["token_types_template"] = ref.state_dict()["token_types_template"]
    ref.load_state_dict(sd)
    diffusion = gd.GaussianDiffusion(betas=gd.get_named_beta_schedule("linear", 1000), model_mean_type="epsilon",
                                     model_var_type="fixed_small", loss_type="mse")
    sampler = smp.PointCloudSampler(device=torch.device("cpu"), models=[ref], diffusions=[diffusion], num_points=[c["num_points"]],
                                    aux_channels=[], guidance_scale=[guidance], clip_denoised=True, use_karras=[True],
                                    karras_steps=[6], sigma_min=[1e-3], sigma_max=[120], s_churn=[0.0])
    _, _, labels, views, _ = inputs(c)
    with patched_noise(cases.DetNoise(777)), torch.no_grad():
        ys = [y.clone() for y in sampler.sample_batch_progressive(c["B"], dict(class_labels=labels, viewpoints=views))]
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", out_name), yields=torch.stack(ys).numpy())
    print("sampler", len(ys), float(ys[-1].abs().mean()))


if __name__ == "__main__":
    main()
    gen_sampler()
    gen_sampler(0.0, "twostream_sampler_small_unguided.npz")
