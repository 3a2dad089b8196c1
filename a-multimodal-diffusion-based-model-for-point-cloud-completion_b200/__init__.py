"""B200-native drop-in for the denoising-sampler hot path of
entheeb/A-Multimodal-Diffusion-Based-Model-for-Point-Cloud-Completion.

Public surface mirrors the reference modules on that path:

    configs.MODEL_CONFIGS / DIFFUSION_CONFIGS / model_from_config / diffusion_from_config
    transformer.*PointDiffusionTransformer            (models/transformer.py)
    perceiver.SimplePerceiver                         (models/perceiver.py)
    rotaryencoderpcd.RotarySelfAttention              (models/rotaryencoderpcd.py)
    k_diffusion.karras_sample_progressive, ...        (diffusion/k_diffusion.py)
    sampler.PointCloudSampler                         (diffusion/sampler.py)
    twostream.TwoStreamDenoiser                       (models/model.py, models/modules.py)
    point_cloud.PointCloud, ply_util.write_ply        (util/point_cloud.py, util/ply_util.py)
    ops.*                                             (kernel-level entry points)

All compute goes through ``libpcd_b200.so`` (hand-written sm_100a CUDA behind the C ABI
in ``include/pcd_b200.h``); there is no CPU or PyTorch fallback.
"""
from . import _lib  # noqa: F401
from .configs import (DIFFUSION_CONFIGS, MODEL_CONFIGS, diffusion_from_config,  # noqa: F401
                      model_from_config)
from .gaussian_diffusion import GaussianDiffusion, get_named_beta_schedule  # noqa: F401
from .k_diffusion import (HeunPlan, get_sigmas_karras, karras_sample,  # noqa: F401
                          karras_sample_progressive)
from .sampler import PointCloudSampler  # noqa: F401
from . import dist, download, ops, perceiver, ply_util, point_cloud, rotaryencoderpcd, transformer, twostream  # noqa: F401
from .twostream import TwoStreamDenoiser  # noqa: F401
from .point_cloud import PointCloud  # noqa: F401

__all__ = ["MODEL_CONFIGS", "DIFFUSION_CONFIGS", "model_from_config", "diffusion_from_config",
           "GaussianDiffusion", "PointCloudSampler", "karras_sample_progressive", "karras_sample",
           "get_sigmas_karras", "HeunPlan", "PointCloud", "ops", "perceiver", "ply_util", "point_cloud",
           "rotaryencoderpcd", "transformer", "twostream", "TwoStreamDenoiser"]
