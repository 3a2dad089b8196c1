"""Karras schedule + Heun solver driven by the fused CUDA sampler kernels.

Drop-in for the reference's ``diffusion/k_diffusion.py`` entry points on the
GaussianDiffusion + "heun" path (karras_sample_progressive :118-222, sample_heun
:270-310, GaussianToKarrasDenoiser :79-108, guided_denoiser :182-207).

Every per-step scalar is computed once on the host with the same fp32 / float64
operations the reference performs per evaluation (including the scipy interp1d
sigma->t lookup and its integer truncation), then passed by value to the kernels.
That removes the reference's three host syncs per step and makes the whole loop
CUDA-graph capturable.
"""
import ctypes as C
from dataclasses import dataclass
from typing import Any, Callable, Dict, Iterator, List, Optional

import numpy as np
import torch as th

from . import _lib
from ._lib import check, ptr, require_cuda, stream_ptr
from .gaussian_diffusion import GaussianDiffusion
from .ops import step_scalars


def append_zero(x):
    return th.cat([x, x.new_zeros([1])])


def append_dims(x, target_dims):
    dims_to_append = target_dims - x.ndim
    if dims_to_append < 0:
        raise ValueError(f"input has {x.ndim} dims but target_dims is {target_dims}, which is less")
    return x[(...,) + (None,) * dims_to_append]


def get_sigmas_karras(n, sigma_min, sigma_max, rho=7.0, device="cpu"):
    """Noise schedule of Karras et al. (2022); computed on the CPU like the reference
    (k_diffusion.py:225-231) so the fp32 values are identical, then moved."""
    ramp = th.linspace(0, 1, n)
    min_inv_rho = sigma_min ** (1 / rho)
    max_inv_rho = sigma_max ** (1 / rho)
    sigmas = (max_inv_rho + ramp * (min_inv_rho - max_inv_rho)) ** rho
    return append_zero(sigmas).to(device)


class KarrasDenoiser:
    """EDM preconditioning (reference k_diffusion.py:31-76).  ``get_scalings`` feeds the fused
    sampler kernels (coef_x = c_skip/c_in, coef_eps = -c_out); ``denoise`` is the plain call."""

    def __init__(self, sigma_data: float = 0.5):
        self.sigma_data = sigma_data

    def get_snr(self, sigmas):
        return sigmas ** -2

    def get_sigmas(self, sigmas):
        return sigmas

    def get_scalings(self, sigma):
        c_skip = self.sigma_data ** 2 / (sigma ** 2 + self.sigma_data ** 2)
        c_out = sigma * self.sigma_data / (sigma ** 2 + self.sigma_data ** 2) ** 0.5
        c_in = 1 / (sigma ** 2 + self.sigma_data ** 2) ** 0.5
        return c_skip, c_out, c_in

    def denoise(self, model, x_t, sigmas, **model_kwargs):
        c_skip, c_out, c_in = [append_dims(x, x_t.ndim) for x in self.get_scalings(sigmas)]
        rescaled_t = 1000 * 0.25 * th.log(sigmas + 1e-44)
        model_output = model(c_in * x_t, rescaled_t, **model_kwargs)
        denoised = c_out * model_output + c_skip * x_t
        return model_output, denoised


class GaussianToKarrasDenoiser:
    """sigma -> integer timestep lookup (reference k_diffusion.py:79-103)."""

    def __init__(self, model, diffusion: GaussianDiffusion):
        from scipy import interpolate

        self.model = model
        self.diffusion = diffusion
        self.alpha_cumprod_to_t = interpolate.interp1d(diffusion.alphas_cumprod,
                                                       np.arange(0, diffusion.num_timesteps))

    def sigma_to_t(self, sigma):
        alpha_cumprod = 1.0 / (sigma ** 2 + 1)
        if alpha_cumprod > self.diffusion.alphas_cumprod[0]:
            return 0
        elif alpha_cumprod <= self.diffusion.alphas_cumprod[-1]:
            return self.diffusion.num_timesteps - 1
        else:
            return float(self.alpha_cumprod_to_t(alpha_cumprod))

    def sigma_to_int_t(self, sigma) -> int:
        # th.tensor([...], dtype=th.long) truncates toward zero (k_diffusion.py:99-103)
        return int(self.sigma_to_t(np.float32(sigma)))


@dataclass
class HeunEval:
    sigma: float      # sigma of this denoiser evaluation (fp32 value)
    t: float          # truncated integer timestep (GaussianDiffusion) or 250*ln(sigma) (KarrasDenoiser)
    c_in: float
    coef_x: float
    coef_eps: float


@dataclass
class HeunStep:
    sigma: float          # sigma_i
    sigma_hat: float      # sigma_i * (1 + gamma)
    noise_scale: float    # sqrt(sigma_hat^2 - sigma_i^2), 0 when gamma == 0
    dt: float             # sigma_{i+1} - sigma_hat
    first: HeunEval
    second: Optional[HeunEval]   # None on the last (Euler) step


def get_ancestral_step(sigma_from, sigma_to):
    """reference k_diffusion.py:239-245."""
    sigma_up = (sigma_to ** 2 * (sigma_from ** 2 - sigma_to ** 2) / sigma_from ** 2) ** 0.5
    sigma_down = (sigma_to ** 2 - sigma_up ** 2) ** 0.5
    return sigma_down, sigma_up


class HeunPlan:
    """All host-side scalars of one stage for ``sampler`` in {"heun", "dpm", "ancestral"}
    (reference sample_heun :270-310, sample_dpm :313-351, sample_euler_ancestral :248-266)."""

    def __init__(self, diffusion, steps: int, sigma_min: float, sigma_max: float,
                 rho: float = 7.0, s_churn: float = 0.0, s_tmin: float = 0.0,
                 s_tmax: float = float("inf"), s_noise: float = 1.0, sampler: str = "heun"):
        if s_noise != 1.0:
            raise NotImplementedError("s_noise != 1 (the reference never changes it)")
        if sampler not in ("heun", "dpm", "ancestral"):
            raise KeyError(sampler)
        self.sampler = sampler
        self.sigmas = get_sigmas_karras(steps, sigma_min, sigma_max, rho)  # CPU fp32 [steps+1]
        self.sigma_max = sigma_max
        self.karras = isinstance(diffusion, KarrasDenoiser)
        wrap = None if self.karras else GaussianToKarrasDenoiser(None, diffusion)
        sig = self.sigmas
        n = len(sig) - 1

        def make_eval(s: th.Tensor) -> HeunEval:
            if self.karras:
                # x0 = c_out*F + c_skip*x = (c_skip/c_in)*(x*c_in) - (-c_out)*F   (k_diffusion.py:71-76)
                c_skip, c_out, c_in = diffusion.get_scalings(s)
                t = float(1000 * 0.25 * th.log(s + 1e-44))
                return HeunEval(float(s), t, float(c_in), float(c_skip / c_in), float(-c_out))
            t = wrap.sigma_to_int_t(s.numpy())
            c_in = 1.0 / (s ** 2 + 1) ** 0.5
            a = np.float32(diffusion.sqrt_recip_alphas_cumprod[t])
            b = np.float32(diffusion.sqrt_recipm1_alphas_cumprod[t])
            return HeunEval(float(s), t, float(c_in), float(a), float(b))

        self.steps: List[HeunStep] = []
        for i in range(n):
            if sampler == "ancestral":
                sigma_down, sigma_up = get_ancestral_step(sig[i], sig[i + 1])
                st = HeunStep(float(sig[i]), float(sig[i]), 0.0, float(sigma_down - sig[i]), make_eval(sig[i]), None)
                st.sigma_up = float(sigma_up)
                self.steps.append(st)
                continue
            gamma = min(s_churn / n, 2 ** 0.5 - 1) if s_tmin <= sig[i] <= s_tmax else 0.0
            sigma_hat = sig[i] * (gamma + 1)
            noise = float((sigma_hat ** 2 - sig[i] ** 2) ** 0.5) if gamma > 0 else 0.0
            if sampler == "dpm":
                # midpoint chosen on a rho=3 Karras schedule; the second evaluation always happens
                sigma_mid = ((sigma_hat ** (1 / 3) + sig[i + 1] ** (1 / 3)) / 2) ** 3
                st = HeunStep(float(sig[i]), float(sigma_hat), noise, float(sigma_mid - sigma_hat),
                              make_eval(sigma_hat), make_eval(sigma_mid))
                st.dt2 = float(sig[i + 1] - sigma_hat)
                self.steps.append(st)
                continue
            dt = sig[i + 1] - sigma_hat
            second = None if sig[i + 1] == 0 else make_eval(sig[i + 1])
            self.steps.append(HeunStep(float(sig[i]), float(sigma_hat), noise, float(dt),
                                       make_eval(sigma_hat), second))

    @property
    def num_evals(self) -> int:
        return sum(1 + (s.second is not None) for s in self.steps)

    def eval_timesteps(self) -> List[float]:
        out = []
        for s in self.steps:
            out.append(s.first.t)
            if s.second is not None:
                out.append(s.second.t)
        return out


class HeunState:
    """Device buffers + kernel launches for one Heun trajectory (state fp32 [B, C, N])."""

    def __init__(self, diffusion: GaussianDiffusion, plan: HeunPlan, shape, device, guidance_scale: float,
                 clip_denoised: bool):
        require_cuda()
        self.lib = _lib.load()
        self.plan = plan
        self.B, self.Cc, self.N = shape
        self.guided = guidance_scale != 0 and guidance_scale != 1
        self.guidance = float(guidance_scale)
        self.clip = 1.0 if clip_denoised else 0.0
        self.x = th.empty(shape, device=device, dtype=th.float32)
        self.d = th.empty(shape, device=device, dtype=th.float32)
        self.model_in = th.empty(shape, device=device, dtype=th.float32)
        f32 = dict(device=device, dtype=th.float32)
        scales = getattr(diffusion, "channel_scales", None)
        biases = getattr(diffusion, "channel_biases", None)
        self.ch_scale = None if scales is None else th.tensor(scales, **f32)
        self.ch_bias = None if biases is None else th.tensor(biases, **f32)

    def _scal(self, ev: Optional[HeunEval], dt: float, nxt: Optional[HeunEval], next_noise: float,
              dt2: float = 0.0, mode: float = 0.0):
        return step_scalars(c_in=ev.c_in if ev else 0.0, coef_x=ev.coef_x if ev else 0.0,
                            coef_eps=ev.coef_eps if ev else 0.0, sigma=ev.sigma if ev else 0.0, dt=dt,
                            guidance=self.guidance, clip=self.clip,
                            next_c_in=nxt.c_in if nxt else 0.0, next_noise=next_noise, dt2=dt2, mode=mode)

    def begin(self, noise0: th.Tensor):
        """x <- x_T (+ churn of step 0); model_in <- x * c_in(sigma_hat_0)."""
        st = self.plan.steps[0]
        s = self._scal(None, 0.0, st.first, st.noise_scale)
        check(self.lib.pcd_sampler_begin(ptr(self.x), ptr(noise0), ptr(self.model_in), C.byref(s),
                                         self.x.numel(), stream_ptr()), "sampler_begin")

    def renoise(self, noise: th.Tensor, sigma_up: float, nxt: Optional[HeunEval]):
        """Euler-ancestral: x <- x + noise*sigma_up; model_in <- x*c_in(next sigma)."""
        s = self._scal(None, 0.0, nxt, sigma_up)
        if nxt is None and sigma_up == 0.0:
            return
        check(self.lib.pcd_sampler_begin(ptr(self.x), ptr(noise), ptr(self.model_in), C.byref(s),
                                         self.x.numel(), stream_ptr()), "sampler_begin")

    def predictor(self, i: int, model_out: th.Tensor, pred_out: th.Tensor):
        st = self.plan.steps[i]
        last = st.second is None
        s = self._scal(st.first, st.dt, st.second, 0.0)
        assert model_out.shape[0] == self.B * (2 if self.guided else 1) and model_out.is_contiguous()
        check(self.lib.pcd_sampler_predictor(ptr(self.x), ptr(model_out), model_out.shape[1], int(self.guided),
                                             ptr(self.d), ptr(self.model_in), ptr(pred_out), ptr(self.ch_scale),
                                             ptr(self.ch_bias), C.byref(s), self.B, self.Cc, self.N, int(last),
                                             stream_ptr()), "sampler_predictor")

    def corrector(self, i: int, model_out: th.Tensor, next_noise: Optional[th.Tensor]):
        st = self.plan.steps[i]
        nxt = self.plan.steps[i + 1] if i + 1 < len(self.plan.steps) else None
        dpm = self.plan.sampler == "dpm"
        s = self._scal(st.second, st.dt, nxt.first if nxt else None, nxt.noise_scale if nxt else 0.0,
                       dt2=getattr(st, "dt2", 0.0), mode=1.0 if dpm else 0.0)
        assert model_out.is_contiguous()
        check(self.lib.pcd_sampler_corrector(ptr(self.x), ptr(model_out), model_out.shape[1], int(self.guided),
                                             ptr(self.d), ptr(next_noise), ptr(self.model_in), C.byref(s),
                                             self.B, self.Cc, self.N, stream_ptr()), "sampler_corrector")


def _native(model) -> bool:
    return bool(getattr(model, "pcd_native", False))


def make_denoiser_eval(model, model_kwargs: Dict[str, Any], B: int, guided: bool, device,
                       eps_channels: Optional[int] = None, karras: bool = False) -> Callable:
    """Returns eval(model_in [B,C,N], t:int) -> model output [B or 2B, C_out, N] (contiguous fp32).

    Native modules evaluate the conditional and unconditional halves as ONE 2B-sequence
    forward sharing x (per-sample arithmetic is independent, SURVEY.md 7.2 "CFG");
    arbitrary callables get the reference's two B-sized calls (k_diffusion.py:182-207),
    with ``prev_latent`` threaded per branch when the model returns a tuple."""
    model_kwargs = model_kwargs or {}
    if _native(model):
        if hasattr(model, "begin_trajectory"):  # models that carry a latent between evaluations (TwoStreamDenoiser)
            seqs = 2 * B if guided else B
            if hasattr(model, "stage_prev_latent"):  # only an unguided run keeps the caller's prev_latent
                model.stage_prev_latent(seqs, None if guided else model_kwargs.get("prev_latent"))
            model.begin_trajectory(seqs)

        def eval_native(model_in, t):
            return model.forward_cfg(model_in, t, model_kwargs, doubled=guided, out_channels=eps_channels)
        return eval_native

    latents = {"cond": None, "uncond": None}

    def call(model_in, tt, kwargs, key):
        kw = dict(kwargs)
        if latents[key] is not None:
            kw["prev_latent"] = latents[key]
        out = model(model_in, tt, **kw)
        if isinstance(out, tuple):
            out, latents[key] = out
        return out.float()

    def eval_generic(model_in, t):
        tt = (th.full((B,), int(t), dtype=th.long, device=device) if float(t).is_integer() and not karras
              else th.full((B,), float(t), dtype=th.float32, device=device))
        if not guided:
            # the reference's unguided `denoiser` passes model_kwargs through unchanged on every call and drops the
            # returned latent (k_diffusion.py:150-166): only guided_denoiser threads prev_latent (:171-203)
            out = model(model_in, tt, **model_kwargs)
            if isinstance(out, tuple):
                out = out[0]
            return out.float().contiguous()
        cond = {k: v[:B] for k, v in model_kwargs.items() if k != "prev_latent"}
        uncond = {k: v[B:] for k, v in model_kwargs.items() if k != "prev_latent"}
        return th.cat([call(model_in, tt, cond, "cond"), call(model_in, tt, uncond, "uncond")], dim=0).contiguous()

    return eval_generic


def karras_sample_progressive(
    diffusion,
    model,
    shape,
    steps,
    clip_denoised=True,
    progress=False,
    model_kwargs=None,
    device=None,
    sigma_min=0.002,
    sigma_max=80,
    rho=7.0,
    sampler="heun",
    s_churn=0.0,
    s_tmin=0.0,
    s_tmax=float("inf"),
    s_noise=1.0,
    guidance_scale=0.0,
    x_target=None,
    noise_fn: Optional[Callable] = None,
) -> Iterator[Dict[str, Any]]:
    """Same signature and yields as the reference (k_diffusion.py:118-222).  ``noise_fn(shape)``
    optionally replaces the torch RNG draws (draw #0 for x_T, then one per step, in the
    reference's order, k_diffusion.py:139,292)."""
    if sampler not in ("heun", "dpm", "ancestral"):
        raise KeyError(sampler)
    karras = isinstance(diffusion, KarrasDenoiser)
    if not karras and not isinstance(diffusion, GaussianDiffusion):
        raise NotImplementedError
    guided = guidance_scale != 0 and guidance_scale != 1
    if karras and guided:
        raise NotImplementedError("classifier-free guidance needs a GaussianDiffusion-wrapped model "
                                  "(the reference's guided path calls model.denoise, k_diffusion.py:194)")
    device = th.device(device if device is not None else "cuda")
    require_cuda()
    if noise_fn is None:
        noise_fn = lambda shp: th.randn(*shp, device=device)
    if sampler == "ancestral":
        plan = HeunPlan(diffusion, steps, sigma_min, sigma_max, rho, sampler=sampler)
    else:
        plan = HeunPlan(diffusion, steps, sigma_min, sigma_max, rho, s_churn, s_tmin, s_tmax, s_noise, sampler)
    B = shape[0]
    state = HeunState(diffusion, plan, tuple(shape), device, guidance_scale, clip_denoised)
    eps_ch = shape[1] if (karras or diffusion.eps_channels_doubled) else None
    evaluate = make_denoiser_eval(model, model_kwargs, B, state.guided, device, eps_ch, karras)
    unscale = (lambda v: v) if karras else diffusion.unscale_channels
    f32 = lambda v: th.tensor(v)

    with th.no_grad():
        state.x.copy_(noise_fn(tuple(shape)).to(device) * sigma_max)
        indices = range(len(plan.steps))
        if progress:
            from tqdm.auto import tqdm
            indices = tqdm(indices)
        pred = None
        if sampler == "ancestral":
            # reference k_diffusion.py:248-266: no churn, one evaluation per step, fresh noise after it
            state.renoise(state.x, 0.0, plan.steps[0].first)  # model_in <- x * c_in(sigma_0)
            for i in indices:
                st = plan.steps[i]
                out = evaluate(state.model_in, st.first.t)
                x_before = state.x.clone()
                pred = th.empty_like(state.x)
                last_flag = st.second  # always None: the predictor applies the Euler step in place
                state.predictor(i, out, pred)
                yield {"x": unscale(x_before), "i": i, "sigma": f32(st.sigma), "sigma_hat": f32(st.sigma),
                       "pred_xstart": pred}
                nxt = plan.steps[i + 1].first if i + 1 < len(plan.steps) else None
                state.renoise(noise_fn(tuple(shape)).to(device), st.sigma_up, nxt)
            yield {"x": unscale(state.x.clone()), "pred_xstart": unscale(state.x.clone())}
            return
        eps = noise_fn(tuple(shape)).to(device)  # always drawn, even when gamma == 0
        state.begin(eps)
        key = "denoised" if sampler == "dpm" else "pred_xstart"
        for i in indices:
            st = plan.steps[i]
            out = evaluate(state.model_in, st.first.t)
            x_hat = state.x.clone()  # the reference yields x after churn, before the update
            pred = th.empty_like(state.x)
            state.predictor(i, out, pred)
            yield {"x": unscale(x_hat), "i": i, "sigma": f32(st.sigma), "sigma_hat": f32(st.sigma_hat), key: pred}
            if st.second is not None:
                out2 = evaluate(state.model_in, st.second.t)
                nxt = noise_fn(tuple(shape)).to(device) if i + 1 < len(plan.steps) else None
                state.corrector(i, out2, nxt)
        yield {"x": unscale(state.x.clone()), "pred_xstart": pred}


def karras_sample(*args, **kwargs):
    last = None
    for x in karras_sample_progressive(*args, **kwargs):
        last = x["x"]
    return last
