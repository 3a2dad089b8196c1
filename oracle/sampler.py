"""Oracle restatement of the Karras/Heun sampling stack (CPU, fp32).

Follows reference diffusion/gaussian_diffusion.py, diffusion/k_diffusion.py and
diffusion/sampler.py.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
"""
import math
from typing import Any, Callable, Dict, Iterator, Optional, Sequence

import numpy as np
import torch

from . import denoiser as D


# ---------------------------------------------------------------------------
# Gaussian diffusion tables (reference diffusion/gaussian_diffusion.py)
# ---------------------------------------------------------------------------
def named_beta_schedule(name: str, n: int) -> np.ndarray:
    """reference gaussian_diffusion.py:26-72 (float64)."""
    if name == "linear":
        scale = 1000 / n
        return np.linspace(scale * 0.0001, scale * 0.02, n, dtype=np.float64)
    if name == "cosine":
        f = lambda t: math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2
        betas = [min(1 - f((i + 1) / n) / f(i / n), 0.999) for i in range(n)]
        return np.array(betas)
    raise NotImplementedError(name)


class Tables:
    """The pieces of GaussianDiffusion.__init__ the Karras path touches
    (reference gaussian_diffusion.py:144-196) plus the channel scaling
    (:938-965)."""

    def __init__(self, schedule="cosine", timesteps=1024, channel_scales=None,
                 channel_biases=None, mean_type="epsilon", **_):
        betas = np.array(named_beta_schedule(schedule, timesteps), dtype=np.float64)
        self.betas = betas
        self.num_timesteps = int(betas.shape[0])
        self.alphas_cumprod = np.cumprod(1.0 - betas, axis=0)
        self.sqrt_recip_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod)
        self.sqrt_recipm1_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod - 1)
        # posterior q(x_{t-1} | x_t, x_0), reference gaussian_diffusion.py:183-196
        prev = np.append(1.0, self.alphas_cumprod[:-1])
        self.posterior_variance = betas * (1.0 - prev) / (1.0 - self.alphas_cumprod)
        self.posterior_log_variance_clipped = np.log(np.append(self.posterior_variance[1], self.posterior_variance[1:]))
        self.posterior_mean_coef1 = betas * np.sqrt(prev) / (1.0 - self.alphas_cumprod)
        self.posterior_mean_coef2 = (1.0 - prev) * np.sqrt(1.0 - betas) / (1.0 - self.alphas_cumprod)
        self.channel_scales = None if channel_scales is None else np.array(channel_scales)
        self.channel_biases = None if channel_biases is None else np.array(channel_biases)
        assert mean_type == "epsilon"

    def unscale(self, x: torch.Tensor) -> torch.Tensor:
        """reference gaussian_diffusion.py:949-958."""
        shape = [1, -1] + [1] * (x.dim() - 2)
        if self.channel_biases is not None:
            x = x - torch.from_numpy(self.channel_biases).to(x).reshape(shape)
        if self.channel_scales is not None:
            x = x / torch.from_numpy(self.channel_scales).to(x).reshape(shape)
        return x

    def pred_xstart(self, model_out: torch.Tensor, x_in: torch.Tensor, t: torch.Tensor,
                    clip: bool = True) -> torch.Tensor:
        """reference gaussian_diffusion.py:285-338,352-357 with learned_range output:
        eps = first C channels; x0 = a_t*x - b_t*eps; clamp to [-1, 1]."""
        C = x_in.shape[1]
        eps = model_out[:, :C]
        a = torch.from_numpy(self.sqrt_recip_alphas_cumprod)[t.cpu()].float().to(x_in.device)
        b = torch.from_numpy(self.sqrt_recipm1_alphas_cumprod)[t.cpu()].float().to(x_in.device)
        while a.dim() < x_in.dim():
            a, b = a[..., None], b[..., None]
        x0 = a * x_in - b * eps
        return x0.clamp(-1, 1) if clip else x0


class SigmaToT:
    """reference k_diffusion.py:79-96 (scipy interp1d over the float64 alpha-bar table;
    the caller truncates to int64, :99-103)."""

    def __init__(self, tables: Tables):
        from scipy import interpolate

        self.ac = tables.alphas_cumprod
        self.T = tables.num_timesteps
        self.f = interpolate.interp1d(self.ac, np.arange(0, self.T))

    def __call__(self, sigma) -> int:
        sigma = np.float32(sigma)
        alpha_cumprod = 1.0 / (sigma ** 2 + 1)
        if alpha_cumprod > self.ac[0]:
            return 0
        if alpha_cumprod <= self.ac[-1]:
            return self.T - 1
        return int(float(self.f(alpha_cumprod)))


def karras_sigmas(n: int, sigma_min: float, sigma_max: float, rho: float = 7.0) -> torch.Tensor:
    """reference k_diffusion.py:225-231 (+ append_zero :362)."""
    ramp = torch.linspace(0, 1, n)
    min_inv_rho = sigma_min ** (1 / rho)
    max_inv_rho = sigma_max ** (1 / rho)
    sigmas = (max_inv_rho + ramp * (min_inv_rho - max_inv_rho)) ** rho
    return torch.cat([sigmas, sigmas.new_zeros([1])])


def heun_progressive(
    model_fn: Callable[..., torch.Tensor],
    tables: Tables,
    shape: Sequence[int],
    steps: int = 64,
    sigma_min: float = 1e-3,
    sigma_max: float = 120.0,
    rho: float = 7.0,
    s_churn: float = 0.0,
    s_tmin: float = 0.0,
    s_tmax: float = float("inf"),
    s_noise: float = 1.0,
    guidance_scale: float = 0.0,
    clip_denoised: bool = True,
    model_kwargs: Optional[Dict[str, torch.Tensor]] = None,
    generator: Optional[torch.Generator] = None,
    trace: Optional[list] = None,
    noise_fn: Optional[Callable] = None,
) -> Iterator[Dict[str, Any]]:
    """reference k_diffusion.py:118-222 (karras_sample_progressive with a
    GaussianDiffusion, sampler="heun") + :270-310 (sample_heun) + :79-108
    (GaussianToKarrasDenoiser) + :182-207 (guided_denoiser).

    ``model_fn(x, t, **kwargs)`` returns [B, 2C, N] (or a tuple whose first item
    is that).  With guidance the kwargs hold 2B rows: [:B] conditional, [B:]
    unconditional, evaluated as two separate B-sized calls.  Yields the
    *unscaled* dicts exactly like the reference.  ``trace`` (optional list)
    collects per-eval (t, x_in, model_out) tuples for per-step parity checks.
    """
    model_kwargs = model_kwargs or {}
    sigmas = karras_sigmas(steps, sigma_min, sigma_max, rho)
    if noise_fn is None:
        noise_fn = lambda shp: torch.randn(*shp, generator=generator)
    x = noise_fn(tuple(shape)) * sigma_max
    s2t = SigmaToT(tables)
    B = shape[0]

    # latent self-conditioning of tuple-returning models (TwoStreamDenoiser): the guided denoiser keeps one
    # latent per branch and feeds it back as prev_latent at the next evaluation (k_diffusion.py:171-203);
    # the unguided path passes model_kwargs through unchanged and drops the latent (:160-164,
    # gaussian_diffusion.py:285-289)
    latents = {"cond": None, "uncond": None}

    def denoise(x_t, sigma, kwargs, branch=None):
        t = torch.tensor([s2t(s) for s in sigma.cpu().numpy()], dtype=torch.long, device=x_t.device)
        c_in = (1.0 / (sigma ** 2 + 1) ** 0.5)[(...,) + (None,) * (x_t.dim() - 1)]
        x_in = x_t * c_in
        kwargs = dict(kwargs)
        if branch is not None and latents[branch] is not None:
            kwargs["prev_latent"] = latents[branch]
        out = model_fn(x_in, t, **kwargs)
        if isinstance(out, tuple):
            out, extra = out
            if branch is not None:
                latents[branch] = extra
        if trace is not None:
            trace.append((t.clone(), x_in.clone(), out.clone()))
        return tables.pred_xstart(out, x_in, t, clip_denoised)

    if guidance_scale != 0 and guidance_scale != 1:
        def denoiser(x_t, sigma):
            cond = denoise(x_t, sigma, {k: v[:B] for k, v in model_kwargs.items() if k != "prev_latent"}, "cond")
            uncond = denoise(x_t, sigma, {k: v[B:] for k, v in model_kwargs.items() if k != "prev_latent"}, "uncond")
            return uncond + guidance_scale * (cond - uncond)
    else:
        def denoiser(x_t, sigma):
            return denoise(x_t, sigma, model_kwargs)

    def to_d(x, sigma, denoised):
        return (x - denoised) / sigma[(...,) + (None,) * (x.dim() - sigma.dim())].to(x.device)

    s_in = x.new_ones([B])
    denoised = None
    for i in range(len(sigmas) - 1):
        gamma = (min(s_churn / (len(sigmas) - 1), 2 ** 0.5 - 1)
                 if s_tmin <= sigmas[i] <= s_tmax else 0.0)
        eps = noise_fn(tuple(x.shape)) * s_noise  # always drawn (:292)
        sigma_hat = sigmas[i] * (gamma + 1)
        if gamma > 0:
            x = x + eps * (sigma_hat ** 2 - sigmas[i] ** 2) ** 0.5
        denoised = denoiser(x, sigma_hat * s_in)
        d = to_d(x, sigma_hat, denoised)
        yield {"x": tables.unscale(x), "i": i, "pred_xstart": tables.unscale(denoised)}
        dt = sigmas[i + 1] - sigma_hat
        if sigmas[i + 1] == 0:
            x = x + d * dt
        else:
            x_2 = x + d * dt
            denoised_2 = denoiser(x_2, sigmas[i + 1] * s_in)
            d_2 = to_d(x_2, sigmas[i + 1], denoised_2)
            x = x + (d + d_2) / 2 * dt
    yield {"x": tables.unscale(x), "pred_xstart": tables.unscale(denoised)}


def sample_batch_progressive(
    model_fns: Sequence[Callable],
    cached_kwargs_fns: Sequence[Optional[Callable]],
    tables: Sequence[Tables],
    num_points: Sequence[int],
    aux_channels: Sequence[str],
    batch_size: int,
    model_kwargs: Dict[str, Any],
    guidance_scale: Sequence[float] = (3.0, 3.0),
    karras_steps: Sequence[int] = (64, 64),
    sigma_min: Sequence[float] = (1e-3, 1e-3),
    sigma_max: Sequence[float] = (120, 160),
    s_churn: Sequence[float] = (3, 0),
    key_filter: Sequence[str] = ("*",),
    clip_denoised: bool = True,
    generator: Optional[torch.Generator] = None,
    noise_fn: Optional[Callable] = None,
) -> Iterator[torch.Tensor]:
    """reference diffusion/sampler.py:96-171 (Karras branch)."""
    n = len(model_fns)
    if len(key_filter) == 1:
        key_filter = list(key_filter) * n
    samples = None
    for s in range(n):
        kw = dict(model_kwargs)
        if key_filter[s] != "*":
            keep = set(key_filter[s].split(","))
            kw = {k: v for k, v in kw.items() if k in keep}
        if samples is not None:
            kw["low_res"] = samples
        if cached_kwargs_fns[s] is not None:
            kw = cached_kwargs_fns[s](batch_size, kw)
        shape = (batch_size, 3 + len(aux_channels), num_points[s])
        g = guidance_scale[s]
        if g != 1 and g != 0:
            kw = {k: torch.cat([v, torch.zeros_like(v)], dim=0) for k, v in kw.items()}
        for out in heun_progressive(
            model_fns[s], tables[s], shape, steps=karras_steps[s], sigma_min=sigma_min[s],
            sigma_max=sigma_max[s], s_churn=s_churn[s], guidance_scale=g,
            clip_denoised=clip_denoised, model_kwargs=kw, generator=generator, noise_fn=noise_fn,
        ):
            samples = out["pred_xstart"][:batch_size]
            if "low_res" in kw:
                samples = torch.cat([kw["low_res"][: len(samples)], samples], dim=-1)
            yield samples


def make_model_fn(sd, cfg):
    """Bind an oracle denoiser to a state dict: model_fn(x, t, **kwargs)."""
    def fn(x, t, **kw):
        return D.denoiser_forward(sd, cfg, x, t, **kw)
    return fn


# ---------------------------------------------------------------------------
# Other solvers / preconditioning of reference k_diffusion.py (rows a7 and f2 of SURVEY 8)
# ---------------------------------------------------------------------------
class KarrasScalings:
    """reference k_diffusion.py:31-45,71-76 (KarrasDenoiser.get_scalings / denoise)."""

    def __init__(self, sigma_data: float = 0.5):
        self.sigma_data = sigma_data

    def denoise(self, model_fn, x_t, sigmas, **kw):
        sd = self.sigma_data
        c_skip = sd ** 2 / (sigmas ** 2 + sd ** 2)
        c_out = sigmas * sd / (sigmas ** 2 + sd ** 2) ** 0.5
        c_in = 1 / (sigmas ** 2 + sd ** 2) ** 0.5
        ex = (...,) + (None,) * (x_t.dim() - 1)
        rescaled_t = 1000 * 0.25 * torch.log(sigmas + 1e-44)
        out = model_fn(c_in[ex] * x_t, rescaled_t, **kw)
        return c_out[ex] * out + c_skip[ex] * x_t


def karras_progressive(model_fn, diffusion, shape, steps, sampler="heun", sigma_min=1e-3, sigma_max=120.0,
                       rho=7.0, s_churn=0.0, clip_denoised=True, model_kwargs=None, noise_fn=None):
    """reference k_diffusion.py:118-222 without guidance, for sampler in {heun, dpm, ancestral} and
    diffusion = Tables (GaussianDiffusion) or KarrasScalings (KarrasDenoiser).  Yields like the reference."""
    model_kwargs = model_kwargs or {}
    sigmas = karras_sigmas(steps, sigma_min, sigma_max, rho)
    x = noise_fn(tuple(shape)) * sigma_max
    B = shape[0]
    gaussian = isinstance(diffusion, Tables)
    if gaussian:
        s2t = SigmaToT(diffusion)

        def denoiser(x_t, sigma):
            t = torch.tensor([s2t(s) for s in sigma.numpy()], dtype=torch.long)
            c_in = (1.0 / (sigma ** 2 + 1) ** 0.5)[(...,) + (None,) * (x_t.dim() - 1)]
            x_in = x_t * c_in
            out = model_fn(x_in, t, **model_kwargs)
            return diffusion.pred_xstart(out, x_in, t, clip_denoised)
        fin = diffusion.unscale
    else:
        def denoiser(x_t, sigma):
            d = diffusion.denoise(model_fn, x_t, sigma, **model_kwargs)
            return d.clamp(-1, 1) if clip_denoised else d
        fin = lambda v: v

    def to_d(x, sigma, den):
        return (x - den) / sigma[(...,) + (None,) * (x.dim() - sigma.dim())]

    s_in = x.new_ones([B])
    n = len(sigmas) - 1
    if sampler == "ancestral":  # :248-266
        for i in range(n):
            den = denoiser(x, sigmas[i] * s_in)
            s_from, s_to = sigmas[i], sigmas[i + 1]
            s_up = (s_to ** 2 * (s_from ** 2 - s_to ** 2) / s_from ** 2) ** 0.5
            s_down = (s_to ** 2 - s_up ** 2) ** 0.5
            yield {"x": fin(x), "i": i, "pred_xstart": fin(den)}
            d = to_d(x, sigmas[i], den)
            x = x + d * (s_down - sigmas[i])
            x = x + noise_fn(tuple(x.shape)) * s_up
        yield {"x": fin(x), "pred_xstart": fin(x)}
        return
    den = None
    for i in range(n):
        gamma = min(s_churn / n, 2 ** 0.5 - 1)
        eps = noise_fn(tuple(x.shape))
        sigma_hat = sigmas[i] * (gamma + 1)
        if gamma > 0:
            x = x + eps * (sigma_hat ** 2 - sigmas[i] ** 2) ** 0.5
        den = denoiser(x, sigma_hat * s_in)
        d = to_d(x, sigma_hat, den)
        yield {"x": fin(x), "i": i, ("denoised" if sampler == "dpm" else "pred_xstart"): fin(den)}
        if sampler == "dpm":  # :313-351
            sigma_mid = ((sigma_hat ** (1 / 3) + sigmas[i + 1] ** (1 / 3)) / 2) ** 3
            x_2 = x + d * (sigma_mid - sigma_hat)
            den_2 = denoiser(x_2, sigma_mid * s_in)
            d_2 = to_d(x_2, sigma_mid, den_2)
            x = x + d_2 * (sigmas[i + 1] - sigma_hat)
        else:  # heun :299-309
            dt = sigmas[i + 1] - sigma_hat
            if sigmas[i + 1] == 0:
                x = x + d * dt
            else:
                x_2 = x + d * dt
                den_2 = denoiser(x_2, sigmas[i + 1] * s_in)
                d_2 = to_d(x_2, sigmas[i + 1], den_2)
                x = x + (d + d_2) / 2 * dt
    yield {"x": fin(x), "pred_xstart": fin(den)}


# ---------------------------------------------------------------------------
# Ancestral (DDPM) sampling of reference gaussian_diffusion.py (row f4 of SURVEY 8)
# ---------------------------------------------------------------------------
def p_mean_variance(tables: Tables, model_out: torch.Tensor, x: torch.Tensor, t: torch.Tensor, var_type: str,
                    clip_denoised: bool, denoised_fn: Optional[Callable] = None):
    """reference gaussian_diffusion.py:257-350 for epsilon-prediction models, given the model output; ``denoised_fn``
    acts on the x_0 prediction before the clamp (:321-326)."""
    ex = lambda arr: torch.from_numpy(arr)[t.cpu()].float().to(x.device)[(...,) + (None,) * (x.dim() - 1)]
    C = x.shape[1]
    if var_type in ("learned", "learned_range"):
        eps, v = model_out[:, :C], model_out[:, C:2 * C]
        if var_type == "learned":
            log_var = v
        else:
            frac = (v + 1) / 2
            log_var = frac * ex(np.log(tables.betas)) + (1 - frac) * ex(tables.posterior_log_variance_clipped)
        var = torch.exp(log_var)
    else:
        eps = model_out
        # fixed variances come straight from the tables (:305-318): "variance" is NOT exp(log_variance) at t = 0 of
        # fixed_small (posterior variance 0, log clipped to the t = 1 value)
        if var_type == "fixed_large":
            v_tab = np.append(tables.posterior_variance[1], tables.betas[1:])
            var, log_var = ex(v_tab).expand_as(x), ex(np.log(v_tab)).expand_as(x)
        else:
            var = ex(tables.posterior_variance).expand_as(x)
            log_var = ex(tables.posterior_log_variance_clipped).expand_as(x)
    x0 = ex(tables.sqrt_recip_alphas_cumprod) * x - ex(tables.sqrt_recipm1_alphas_cumprod) * eps
    if denoised_fn is not None:
        x0 = denoised_fn(x0)
    if clip_denoised:
        x0 = x0.clamp(-1, 1)
    mean = ex(tables.posterior_mean_coef1) * x0 + ex(tables.posterior_mean_coef2) * x
    return dict(mean=mean, log_variance=log_var, variance=var, pred_xstart=x0)


def ddpm_progressive(model_fn, tables: Tables, shape, var_type: str, clip_denoised: bool = True, model_kwargs=None,
                     noise_fn: Optional[Callable] = None, denoised_fn: Optional[Callable] = None,
                     cond_fn: Optional[Callable] = None):
    """reference gaussian_diffusion.py:407-449 (p_sample) inside :499-548 (p_sample_loop_progressive): x_T ~ N(0, I);
    for t = T-1 .. 0: x <- mean + [t != 0] exp(log_var / 2) noise, with the mean shifted by variance * cond_fn(x, t)
    when a ``cond_fn`` is given (condition_mean, :374-385).  Yields the unscaled dicts like the reference."""
    model_kwargs = model_kwargs or {}
    if noise_fn is None:
        noise_fn = lambda shp: torch.randn(*shp)
    x = noise_fn(tuple(shape))
    for i in reversed(range(tables.num_timesteps)):
        t = torch.full((shape[0],), i, dtype=torch.long)
        out = model_fn(x, t, **model_kwargs)
        if isinstance(out, tuple):
            out = out[0]
        r = p_mean_variance(tables, out, x, t, var_type, clip_denoised, denoised_fn)
        if cond_fn is not None:
            r["mean"] = r["mean"] + r["variance"] * cond_fn(x, t, **model_kwargs)
        noise = noise_fn(tuple(shape))
        sample = r["mean"] + (0.0 if i == 0 else 1.0) * torch.exp(0.5 * r["log_variance"]) * noise
        yield {"sample": tables.unscale(sample), "pred_xstart": tables.unscale(r["pred_xstart"])}
        x = sample
