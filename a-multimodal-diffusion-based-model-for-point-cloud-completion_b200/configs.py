"""Model / diffusion registries (reference models/configs.py:15-134, diffusion/configs.py:16-64)."""
from typing import Any, Dict

import numpy as np
import torch
import torch.nn as nn

from .gaussian_diffusion import GaussianDiffusion, get_named_beta_schedule
from .transformer import (CLIPImageGridPointDiffusionTransformer,
                          CLIPImageGridUpsamplePointDiffusionTransformer,
                          CLIPImagePointDiffusionTransformer, PointDiffusionTransformer,
                          UpsamplePointDiffusionTransformer)

_POINT_E = dict(heads=8, init_scale=0.25, input_channels=6, layers=12, n_ctx=1024, output_channels=12,
                time_token_cond=True, width=512)
_CH_BIASES = [0.0, 0.0, 0.0, -1.0, -1.0, -1.0]
_CH_SCALES = [2.0, 2.0, 2.0, 0.007843137255, 0.007843137255, 0.007843137255]

MODEL_CONFIGS: Dict[str, Dict[str, Any]] = {
    "base40M-imagevec": dict(_POINT_E, name="CLIPImagePointDiffusionTransformer", cond_drop_prob=0.1,
                             token_cond=True),
    "base40M-textvec": dict(_POINT_E, name="CLIPImagePointDiffusionTransformer", cond_drop_prob=0.1,
                            token_cond=True),
    "base40M-uncond": dict(_POINT_E, name="PointDiffusionTransformer"),
    "base40M": dict(_POINT_E, name="CLIPImageGridPointDiffusionTransformer", cond_drop_prob=0.1),
    "base300M": dict(_POINT_E, name="CLIPImageGridPointDiffusionTransformer", cond_drop_prob=0.1,
                     heads=16, layers=24, width=1024),
    "base1B": dict(_POINT_E, name="CLIPImageGridPointDiffusionTransformer", cond_drop_prob=0.1,
                   heads=32, layers=24, width=2048),
    "upsample": dict(_POINT_E, name="CLIPImageGridUpsamplePointDiffusionTransformer", cond_drop_prob=0.1,
                     cond_ctx=1024, n_ctx=3072, channel_biases=_CH_BIASES, channel_scales=_CH_SCALES),
}

_CLASSES = {c.__name__: c for c in (
    PointDiffusionTransformer, CLIPImagePointDiffusionTransformer, CLIPImageGridPointDiffusionTransformer,
    UpsamplePointDiffusionTransformer, CLIPImageGridUpsamplePointDiffusionTransformer)}


def model_from_config(config: Dict[str, Any], device: torch.device,
                      dtype: torch.dtype = torch.bfloat16) -> nn.Module:
    """Reference signature plus ``dtype``: torch.bfloat16 (tensor-core mode, default) or
    torch.float32 (CUDA-core 1e-4 parity mode).  The reference hard-wires fp32
    (models/configs.py:121-133)."""
    config = config.copy()
    name = config.pop("name")
    if name not in _CLASSES:
        raise ValueError(f"unknown model name: {name}")
    return _CLASSES[name](device=device, dtype=dtype, **config)


BASE_DIFFUSION_CONFIG = {
    "channel_biases": _CH_BIASES,
    "channel_scales": _CH_SCALES,
    "mean_type": "epsilon",
    "schedule": "cosine",
    "timesteps": 1024,
}

DIFFUSION_CONFIGS = {
    "base40M-imagevec": BASE_DIFFUSION_CONFIG,
    "base40M-textvec": BASE_DIFFUSION_CONFIG,
    "base40M-uncond": BASE_DIFFUSION_CONFIG,
    "base40M": BASE_DIFFUSION_CONFIG,
    "base300M": BASE_DIFFUSION_CONFIG,
    "base1B": BASE_DIFFUSION_CONFIG,
    "upsample": dict(BASE_DIFFUSION_CONFIG, schedule="linear"),
}


def diffusion_from_config(config: Dict[str, Any]) -> GaussianDiffusion:
    schedule = config["schedule"]
    steps = config["timesteps"]
    if config.get("respacing", None) is not None:
        raise NotImplementedError("respaced (SpacedDiffusion) schedules are not used by the sampler path")
    betas = get_named_beta_schedule(schedule, steps)
    channel_scales = config.get("channel_scales", None)
    channel_biases = config.get("channel_biases", None)
    return GaussianDiffusion(
        betas=betas,
        model_mean_type=config.get("mean_type", "epsilon"),
        model_var_type="learned_range",
        loss_type="mse",
        channel_scales=None if channel_scales is None else np.array(channel_scales),
        channel_biases=None if channel_biases is None else np.array(channel_biases),
    )
