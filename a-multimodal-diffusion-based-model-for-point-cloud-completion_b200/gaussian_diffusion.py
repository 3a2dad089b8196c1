"""Host-side Gaussian diffusion tables (float64, numpy) for the Karras sampling path, and the ancestral
(DDPM) sampling loop on the fused ``pcd_ddpm_step`` kernel.

Mirrors the pieces of the reference's ``GaussianDiffusion`` that sampling touches
(diffusion/gaussian_diffusion.py:26-72 schedules, :144-196 tables, :257-350 p_mean_variance, :407-548
p_sample / p_sample_loop / p_sample_loop_progressive, :938-965 channel scaling).  Training losses and DDIM are
out of scope.
"""
import ctypes as C
import math
from typing import Any, Callable, Dict, Optional, Sequence, Union

import numpy as np
import torch as th

from . import _lib


def get_beta_schedule(beta_schedule, *, beta_start, beta_end, num_diffusion_timesteps):
    if beta_schedule != "linear":
        raise NotImplementedError(beta_schedule)
    return np.linspace(beta_start, beta_end, num_diffusion_timesteps, dtype=np.float64)


def betas_for_alpha_bar(num_diffusion_timesteps, alpha_bar, max_beta=0.999):
    betas = []
    for i in range(num_diffusion_timesteps):
        t1 = i / num_diffusion_timesteps
        t2 = (i + 1) / num_diffusion_timesteps
        betas.append(min(1 - alpha_bar(t2) / alpha_bar(t1), max_beta))
    return np.array(betas)


def get_named_beta_schedule(schedule_name: str, num_diffusion_timesteps: int) -> np.ndarray:
    if schedule_name == "linear":
        scale = 1000 / num_diffusion_timesteps
        return get_beta_schedule("linear", beta_start=scale * 0.0001, beta_end=scale * 0.02,
                                 num_diffusion_timesteps=num_diffusion_timesteps)
    if schedule_name == "cosine":
        return betas_for_alpha_bar(num_diffusion_timesteps,
                                   lambda t: math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2)
    raise NotImplementedError(f"unknown beta schedule: {schedule_name}")


class GaussianDiffusion:
    """Same constructor keywords as the reference class (gaussian_diffusion.py:144-153)."""

    def __init__(self, *, betas: Sequence[float], model_mean_type: str = "epsilon",
                 model_var_type: str = "learned_range", loss_type: str = "mse",
                 discretized_t0: bool = False, channel_scales: Optional[np.ndarray] = None,
                 channel_biases: Optional[np.ndarray] = None):
        if model_mean_type != "epsilon":
            raise NotImplementedError("the fused sampler kernels implement epsilon prediction "
                                      "(every registered config, diffusion/configs.py:16-38)")
        self.model_mean_type = model_mean_type
        self.model_var_type = model_var_type
        self.loss_type = loss_type
        self.discretized_t0 = discretized_t0
        self.channel_scales = None if channel_scales is None else np.asarray(channel_scales, dtype=np.float64)
        self.channel_biases = None if channel_biases is None else np.asarray(channel_biases, dtype=np.float64)
        betas = np.array(betas, dtype=np.float64)
        assert betas.ndim == 1 and (betas > 0).all() and (betas <= 1).all()
        self.betas = betas
        self.num_timesteps = int(betas.shape[0])
        alphas = 1.0 - betas
        self.alphas_cumprod = np.cumprod(alphas, axis=0)
        self.alphas_cumprod_prev = np.append(1.0, self.alphas_cumprod[:-1])
        self.sqrt_alphas_cumprod = np.sqrt(self.alphas_cumprod)
        self.sqrt_one_minus_alphas_cumprod = np.sqrt(1.0 - self.alphas_cumprod)
        self.sqrt_recip_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod)
        self.sqrt_recipm1_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod - 1)
        # posterior q(x_{t-1} | x_t, x_0) (gaussian_diffusion.py:183-196); its variance is 0 at t = 0, hence the clip
        self.posterior_variance = betas * (1.0 - self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_log_variance_clipped = (np.log(np.append(self.posterior_variance[1], self.posterior_variance[1:]))
                                               if len(betas) > 1 else np.zeros(1))
        self.posterior_mean_coef1 = betas * np.sqrt(self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_mean_coef2 = (1.0 - self.alphas_cumprod_prev) * np.sqrt(alphas) / (1.0 - self.alphas_cumprod)
        self._ddpm_tables: Dict[Any, th.Tensor] = {}

    @property
    def eps_channels_doubled(self) -> bool:
        """learned / learned_range models emit 2C channels; only the first C (epsilon)
        feed the Karras path (gaussian_diffusion.py:291-303)."""
        return self.model_var_type in ("learned", "learned_range")

    def _chan(self, arr, x):
        return th.from_numpy(arr).to(x).reshape([1, -1, *([1] * (x.dim() - 2))])

    def scale_channels(self, x: th.Tensor) -> th.Tensor:
        if self.channel_scales is not None:
            x = x * self._chan(self.channel_scales, x)
        if self.channel_biases is not None:
            x = x + self._chan(self.channel_biases, x)
        return x

    def unscale_channels(self, x: th.Tensor) -> th.Tensor:
        if self.channel_biases is not None:
            x = x - self._chan(self.channel_biases, x)
        if self.channel_scales is not None:
            x = x / self._chan(self.channel_scales, x)
        return x

    def unscale_out_dict(self, out: Dict[str, Union[th.Tensor, Any]]) -> Dict[str, Union[th.Tensor, Any]]:
        return {k: (self.unscale_channels(v) if isinstance(v, th.Tensor) and v.dim() >= 2 else v)
                for k, v in out.items()}

    # ------------------------------------------------------------------------------------------------
    # Ancestral (DDPM) sampling: p_mean_variance / p_sample / p_sample_loop(_progressive)
    # (gaussian_diffusion.py:257-350, 407-548) on the fused pcd_ddpm_step kernel.
    # ------------------------------------------------------------------------------------------------
    def _var_mode(self) -> int:
        modes = {"fixed_small": _lib.VAR_FIXED, "fixed_large": _lib.VAR_FIXED, "learned_range": _lib.VAR_LEARNED_RANGE,
                 "learned": _lib.VAR_LEARNED}
        if self.model_var_type not in modes:
            raise KeyError(self.model_var_type)
        return modes[self.model_var_type]

    def _ddpm_table(self, device) -> th.Tensor:
        """[T, 8] float32 schedule rows for pcd_ddpm_step (columns PCD_DDPM_*, include/pcd_b200.h)."""
        key = str(device)
        if key not in self._ddpm_tables:
            if self.model_var_type == "fixed_large":  # gaussian_diffusion.py:305-311
                fixed = np.log(np.append(self.posterior_variance[1], self.betas[1:]))
            else:
                fixed = self.posterior_log_variance_clipped
            tab = np.zeros((self.num_timesteps, _lib.DDPM_COLS), dtype=np.float64)
            for col, arr in enumerate((self.sqrt_recip_alphas_cumprod, self.sqrt_recipm1_alphas_cumprod,
                                       self.posterior_mean_coef1, self.posterior_mean_coef2,
                                       self.posterior_log_variance_clipped, np.log(self.betas), fixed)):
                tab[:, col] = arr
            self._ddpm_tables[key] = th.from_numpy(tab).float().to(device).contiguous()
        return self._ddpm_tables[key]

    def _channel_tensors(self, device):
        """channel_scales / channel_biases on the device (cached; None when the diffusion has none)."""
        key = ("channels", str(device))
        if key not in self._ddpm_tables:
            f32 = dict(device=device, dtype=th.float32)
            self._ddpm_tables[key] = (None if self.channel_scales is None else th.tensor(self.channel_scales, **f32),
                                      None if self.channel_biases is None else th.tensor(self.channel_biases, **f32))
        return self._ddpm_tables[key]

    def _ddpm_step(self, model, x, t, clip_denoised, model_kwargs, noise, outputs):
        """One fused p_mean_variance (+ p_sample when ``noise`` is given).  ``outputs``: names of the optional result
        tensors to materialise among x_next / pred_xstart / sample_unscaled / mean / log_variance."""
        _lib.require_cuda(x)
        B, Cc = x.shape[:2]
        assert t.shape == (B,)
        assert x.dtype == th.float32, "the sampler state is fp32"
        out = model(x, t, **(model_kwargs or {}))
        extra = None
        if isinstance(out, tuple):
            out, extra = out
        mode = self._var_mode()
        need = Cc * (1 if mode == _lib.VAR_FIXED else 2)
        assert out.shape[0] == B and out.shape[1] == need and out.shape[2:] == x.shape[2:], \
            f"model output {tuple(out.shape)} does not match state {tuple(x.shape)} for model_var_type={self.model_var_type}"
        x = x.contiguous()
        out = out.float().contiguous()
        n_points = int(np.prod(x.shape[2:]))
        res = {k: th.empty_like(x) for k in outputs}
        a = _lib.DdpmArgs()
        a.x, a.model_out, a.noise = _lib.ptr(x), _lib.ptr(out), _lib.ptr(noise.contiguous() if noise is not None else None)
        if not t.is_cuda and bool(((t < 0) | (t >= self.num_timesteps)).any()):
            raise IndexError("timestep out of range")  # (device-side indices are trusted: checking them would sync)
        tt = t.to(device=x.device, dtype=th.int64).contiguous()
        a.t, a.table = _lib.ptr(tt), _lib.ptr(self._ddpm_table(x.device))
        sc, bi = self._channel_tensors(x.device)
        a.ch_scale, a.ch_bias = _lib.ptr(sc), _lib.ptr(bi)
        for k in ("x_next", "pred_xstart", "sample_unscaled", "mean", "log_variance"):
            setattr(a, k, _lib.ptr(res.get(k)))
        a.batch, a.channels, a.n_points, a.out_channels = B, Cc, n_points, out.shape[1]
        a.var_mode, a.clip_denoised, a.unscale = mode, int(bool(clip_denoised)), int("sample_unscaled" in outputs)
        a.num_timesteps = self.num_timesteps
        _lib.check(_lib.load().pcd_ddpm_step(C.byref(a), _lib.stream_ptr()), "ddpm_step")
        res["extra"] = extra
        return res

    def _variance_of(self, r, x, t):
        if self._var_mode() == _lib.VAR_FIXED:
            # table lookup like the reference (:305-318): at t = 0 of fixed_small this is 0, not exp(clipped log)
            tab = (np.append(self.posterior_variance[1], self.betas[1:]) if self.model_var_type == "fixed_large"
                   else self.posterior_variance)
            row = th.from_numpy(tab).float().to(x.device)[t.to(x.device).long()]
            return row.view(-1, *([1] * (x.dim() - 1))).expand_as(x).contiguous()
        return th.exp(r["log_variance"])

    def p_mean_variance(self, model, x, t, clip_denoised=False, denoised_fn=None, model_kwargs=None):
        """Same contract as the reference (:257-350): dict with mean / variance / log_variance / pred_xstart / extra.
        ``denoised_fn`` (an arbitrary callable on the x_0 prediction, applied before the clamp, :321-326) splits the
        fused step at the hook: the kernel produces the raw x_0 and the variance, the posterior mean of the hooked x_0
        (:240-243) is two device-side multiply-adds."""
        if denoised_fn is None:
            r = self._ddpm_step(model, x, t, clip_denoised, model_kwargs, None, ("mean", "log_variance", "pred_xstart"))
            mean, x0 = r["mean"], r["pred_xstart"]
        else:
            r = self._ddpm_step(model, x, t, False, model_kwargs, None, ("log_variance", "pred_xstart"))
            x0 = denoised_fn(r["pred_xstart"])
            if clip_denoised:
                x0 = x0.clamp(-1, 1)
            rows = self._ddpm_table(x.device)[t.to(x.device).long()]
            col = lambda c: rows[:, c].view(-1, *([1] * (x.dim() - 1)))
            mean = col(_lib.DDPM_MEAN_X0) * x0 + col(_lib.DDPM_MEAN_XT) * x
        return {"mean": mean, "variance": self._variance_of(r, x, t), "log_variance": r["log_variance"],
                "pred_xstart": x0, "extra": r["extra"]}

    def condition_mean(self, cond_fn, p_mean_var, x, t, model_kwargs=None):
        """mean + variance * grad log p(y | x) (:374-385, Sohl-Dickstein et al. conditioning)."""
        gradient = cond_fn(x, t, **(model_kwargs or {}))
        return p_mean_var["mean"].float() + p_mean_var["variance"] * gradient.float()

    def p_sample(self, model, x, t, clip_denoised=False, denoised_fn=None, cond_fn=None, model_kwargs=None,
                 noise: Optional[th.Tensor] = None):
        """x_{t-1} ~ p(. | x_t) (:407-449).  ``noise`` optionally replaces the torch draw (tests).  Without hooks this is
        ONE fused kernel; with ``denoised_fn`` / ``cond_fn`` the step is split around the callables."""
        if noise is None:
            noise = th.randn_like(x)
        if denoised_fn is None and cond_fn is None:
            r = self._ddpm_step(model, x, t, clip_denoised, model_kwargs, noise.to(x), ("x_next", "pred_xstart"))
            return {"sample": r["x_next"], "pred_xstart": r["pred_xstart"]}
        out = self.p_mean_variance(model, x, t, clip_denoised=clip_denoised, denoised_fn=denoised_fn,
                                   model_kwargs=model_kwargs)
        if cond_fn is not None:
            out["mean"] = self.condition_mean(cond_fn, out, x, t, model_kwargs=model_kwargs)
        nonzero = (t.to(x.device) != 0).float().view(-1, *([1] * (x.dim() - 1)))  # no noise when t == 0
        sample = out["mean"] + nonzero * th.exp(0.5 * out["log_variance"]) * noise.to(x)
        return {"sample": sample, "pred_xstart": out["pred_xstart"]}

    def p_sample_loop_progressive(self, model, shape, noise=None, clip_denoised=False, denoised_fn=None, cond_fn=None,
                                  model_kwargs=None, device=None, progress=False, temp=1.0,
                                  noise_fn: Optional[Callable] = None):
        """Generator over the T ancestral steps (:499-548); every yield is the *unscaled* {"sample", "pred_xstart"}
        dict like the reference's ``unscale_out_dict(out)``.  ``noise_fn(shape)`` optionally replaces the torch draws
        (x_T first, then one per step, the reference's order)."""
        if device is None:
            device = next(model.parameters()).device
        assert isinstance(shape, (tuple, list))
        draw = noise_fn if noise_fn is not None else (lambda shp: th.randn(*shp, device=device))
        img = noise if noise is not None else draw(tuple(shape)).to(device) * temp
        img = img.to(device=device, dtype=th.float32).contiguous()
        indices = list(range(self.num_timesteps))[::-1]
        if progress:
            from tqdm.auto import tqdm
            indices = tqdm(indices)
        hooked = denoised_fn is not None or cond_fn is not None
        with th.no_grad():
            for i in indices:
                t = th.full((shape[0],), i, device=device, dtype=th.int64)
                if hooked:
                    out = self.p_sample(model, img, t, clip_denoised=clip_denoised, denoised_fn=denoised_fn,
                                        cond_fn=cond_fn, model_kwargs=model_kwargs, noise=draw(tuple(shape)).to(img))
                    yield self.unscale_out_dict(out)
                    img = out["sample"].contiguous()
                    continue
                r = self._ddpm_step(model, img, t, clip_denoised, model_kwargs, draw(tuple(shape)).to(img),
                                    ("x_next", "pred_xstart", "sample_unscaled"))
                yield {"sample": r["sample_unscaled"], "pred_xstart": r["pred_xstart"]}
                img = r["x_next"]

    def p_sample_loop(self, model, shape, **kwargs):
        final = None
        for sample in self.p_sample_loop_progressive(model, shape, **kwargs):
            final = sample
        return final["sample"]
