"""Host-side Gaussian diffusion tables (float64, numpy) for the Karras sampling path.

Mirrors the pieces of the reference's ``GaussianDiffusion`` that the sampler touches
(diffusion/gaussian_diffusion.py:26-72 schedules, :144-196 tables, :938-965 channel
scaling).  Training losses, DDIM and the ancestral p_sample loop are out of scope
(SURVEY.md 8a row a8).
"""
import math
from typing import Any, Dict, Optional, Sequence, Union

import numpy as np
import torch as th


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
