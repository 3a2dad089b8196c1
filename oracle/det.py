"""Platform-independent deterministic tensors for fixtures (TEST INFRASTRUCTURE).

A splitmix64 hash of (seed, index) in exact uint64 arithmetic -> uniform in
[0,1) with 53 bits -> either a scaled uniform (weights) or a Box-Muller normal
(noise / inputs).  No dependence on torch's RNG, so the same tensors can be
regenerated on any box without storing them in the fixtures.
"""
import re

import numpy as np
import torch

_M = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M
        z = x
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M
        return z ^ (z >> np.uint64(31))


def _uniform01(n: int, seed: int, stream: int = 0) -> np.ndarray:
    with np.errstate(over="ignore"):
        base = _splitmix64(np.array([seed * 2 + stream], dtype=np.uint64))[0]
        idx = np.arange(n, dtype=np.uint64) * np.uint64(0xD1342543DE82EF95) + base
    h = _splitmix64(idx)
    return (h >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def uniform(shape, seed: int, std: float = 1.0) -> torch.Tensor:
    """Zero-mean uniform with the given standard deviation, float32."""
    n = int(np.prod(shape)) if len(shape) else 1
    u = _uniform01(n, seed)
    a = std * np.sqrt(3.0)
    return torch.from_numpy(((2.0 * u - 1.0) * a).astype(np.float32).reshape(shape))


def normal(shape, seed: int, std: float = 1.0) -> torch.Tensor:
    """Box-Muller standard normal (float64 math, rounded to float32)."""
    n = int(np.prod(shape)) if len(shape) else 1
    u1 = _uniform01(n, seed, 0)
    u2 = _uniform01(n, seed, 1)
    z = np.sqrt(-2.0 * np.log(1.0 - u1)) * np.cos(2.0 * np.pi * u2)
    return torch.from_numpy((z * std).astype(np.float32).reshape(shape))


def fill_state_dict(reference_sd, seed: int, mode: str = "unit", width: int = 0):
    """Deterministic weights for every floating tensor of a reference-shaped
    state dict (keys sorted, one hash stream per tensor).

    mode="unit":      matrices ~ U with std 1/sqrt(fan_in) (O(1) activations and
                      attention logits with std ~1: a demanding parity case).
    mode="reference": the reference's own init scales -- std 0.25/sqrt(width) for the
                      backbone and time-MLP matrices (transformer.py:17-20,134,175),
                      1/sqrt(3 fan_in) (nn.Linear default) for the other projections;
                      the zero-initialised output_proj (transformer.py:191-193) gets
                      std 0.02.
    LayerNorm gains ~ 1 + 0.1 U, LayerNorm biases ~ 0.05 U, other biases ~ 0.02 U so
    that every term of the forward is exercised; non-float buffers and the
    channel_scales / channel_biases buffers are kept."""
    out = {}
    for i, (k, v) in enumerate(sorted(reference_sd.items())):
        if (not torch.is_floating_point(v)) or k in ("channel_scales", "channel_biases"):
            out[k] = v.clone()
            continue
        s = seed * 1000 + i
        shape = tuple(v.shape)
        is_ln = bool(re.search(r"(^|\.)(ln_\w+|norm\w*|clip_embed\.0)\.", k))
        if v.dim() >= 2:
            fan_in = shape[-1]
            if mode == "unit":
                std = 1.0 / np.sqrt(fan_in)
            elif k.startswith("output_proj"):
                std = 0.02
            elif k.startswith("backbone.") or k.startswith("time_embed.") or k.startswith("resblocks."):
                std = 0.25 / np.sqrt(width)
            else:
                std = 1.0 / np.sqrt(3.0 * fan_in)
            out[k] = uniform(shape, s, std)
        elif is_ln:
            out[k] = (1.0 + uniform(shape, s, 0.1)) if k.endswith("weight") else uniform(shape, s, 0.05)
        else:
            out[k] = uniform(shape, s, 0.02)
    return out
