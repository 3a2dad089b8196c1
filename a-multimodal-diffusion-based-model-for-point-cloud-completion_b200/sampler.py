"""Multi-stage point-cloud sampler on the fused Heun kernels.

Same constructor, attributes and methods as the reference's
``diffusion/sampler.py`` ``PointCloudSampler`` (:16-291).  Two execution modes:

* eager (default): ``sample_batch_progressive`` yields after every Heun step, like
  the reference;
* ``use_cuda_graph=True``: every stage's 64-step loop (all denoiser evaluations and
  sampler updates, ~10^4 kernel launches) is captured once per (stage, batch) into a
  CUDA graph and replayed; the 65 yields of a stage become available when the replay
  finishes.  Noise is drawn by torch in the reference's order before the replay, so a
  fixed seed gives the same trajectory as the eager mode.
"""
import dataclasses
from typing import Any, Callable, Dict, Iterator, List, Optional, Sequence, Tuple

import torch

from .point_cloud import PointCloud
import torch.nn as nn

from .gaussian_diffusion import GaussianDiffusion
from .k_diffusion import HeunPlan, HeunState, karras_sample_progressive, make_denoiser_eval
from ._lib import require_cuda


class _GraphedStage:
    """One stage's whole Heun loop as a CUDA graph over persistent buffers."""

    def __init__(self, model, diffusion: GaussianDiffusion, plan: HeunPlan, shape, device, guidance: float,
                 clip_denoised: bool):
        self.model, self.diffusion, self.plan = model, diffusion, plan
        self.shape, self.device = tuple(shape), device
        self.state = HeunState(diffusion, plan, self.shape, device, guidance, clip_denoised)
        n = len(plan.steps)
        self.noise = torch.empty((n + 1,) + self.shape, device=device, dtype=torch.float32)
        self.preds = torch.empty((n,) + self.shape, device=device, dtype=torch.float32)
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.kwargs: Dict[str, Any] = {}
        self.eps_channels = self.shape[1] if diffusion.eps_channels_doubled else None
        self.launches_per_replay = 0

    def _enqueue(self, max_steps: Optional[int] = None):
        st, plan = self.state, self.plan
        B = self.shape[0]
        torch.mul(self.noise[0], plan.sigma_max, out=st.x)
        if hasattr(self.model, "begin_trajectory"):
            self.model.begin_trajectory(B * (2 if st.guided else 1))  # an unguided run's prev_latent was staged by run()
        st.begin(self.noise[1])
        for i, step in enumerate(plan.steps[:max_steps]):
            out = self.model.forward_cfg(st.model_in, step.first.t, self.kwargs, st.guided, self.eps_channels)
            st.predictor(i, out, self.preds[i])
            if step.second is not None:
                out2 = self.model.forward_cfg(st.model_in, step.second.t, self.kwargs, st.guided,
                                              self.eps_channels)
                st.corrector(i, out2, self.noise[i + 2])

    def run(self, kwargs: Dict[str, Any], noise_fn: Optional[Callable]) -> torch.Tensor:
        """Returns preds [steps, B, C, N] (unscaled pred_xstart of every step)."""
        seqs = self.shape[0] * (2 if self.state.guided else 1)
        # noise in the reference's draw order: x_T, then one eps per step (k_diffusion.py:139,292)
        for k in range(self.noise.shape[0]):
            if noise_fn is None:
                self.noise[k].normal_()
            else:
                self.noise[k].copy_(noise_fn(self.shape))
        self.kwargs = {k: v for k, v in kwargs.items() if k != "prev_latent"}
        if hasattr(self.model, "stage_prev_latent"):  # only an unguided run keeps the caller's prev_latent
            self.model.stage_prev_latent(seqs, None if self.state.guided else kwargs.get("prev_latent"))
        if getattr(self.model, "cfg_halves", False):
            self.model.prepare_cond(seqs, self.kwargs, self.state.guided)
        else:
            self.model.prepare_cond(seqs, self.kwargs)
        self.model.prepare_time_tokens(self.plan.eval_timesteps())
        if self.graph is None:
            # warm-up on a side stream, then capture.  ONE Heun step is enough to populate every cache the capture
            # relies on (workspace, conditioning tokens; the time tokens of the whole schedule were prepared above)
            from . import _lib
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                self._enqueue(max_steps=1)
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            n0 = _lib.load().pcd_launch_count()
            with torch.cuda.graph(self.graph):
                self._enqueue()
            self.launches_per_replay = int(_lib.load().pcd_launch_count() - n0)  # kernels of this library in the graph
        self.graph.replay()
        return self.preds


# Per-stage options of the sampler, in constructor order, with the value a stage after the first gets when the
# caller passed a single entry (reference diffusion/sampler.py:45-62: everything repeats, except that upsampler
# stages are not guided unless asked).
_STAGE_OPTIONS = ("model_kwargs_key_filter", "guidance_scale", "use_karras", "karras_steps", "sigma_min", "sigma_max",
                  "s_churn")
_COLOUR_CHANNELS = frozenset("RGBA")


@dataclasses.dataclass
class _Stage:
    index: int
    model: Any
    diffusion: Any
    num_points: int
    key_filter: str
    guidance: float
    karras: bool
    steps: int
    sigma_min: float
    sigma_max: float
    churn: float

    @property
    def guided(self) -> bool:
        return self.guidance not in (0, 1)


class PointCloudSampler:
    """Runs a cascade of point-cloud diffusion models (e.g. base 1024 points -> upsampler 4096) and yields the clouds.

    Drop-in for the reference's ``diffusion/sampler.py`` ``PointCloudSampler`` (:16-291): same constructor keywords and
    defaults, the same public attributes (``models``, ``diffusions``, ``num_points``, ``guidance_scale``, ...) and methods
    (``sample_batch``, ``sample_batch_progressive``, ``combine``, ``with_options``, ``split_model_output``,
    ``output_to_point_clouds``, ``num_stages``).  Extras: ``use_cuda_graph`` and ``noise_fn`` (deterministic draws)."""

    def __init__(self, device: torch.device, models: Sequence[nn.Module], diffusions: Sequence[GaussianDiffusion],
                 num_points: Sequence[int], aux_channels: Sequence[str],
                 model_kwargs_key_filter: Sequence[str] = ("*",), guidance_scale: Sequence[float] = (3.0, 3.0),
                 clip_denoised: bool = True, use_karras: Sequence[bool] = (True, True),
                 karras_steps: Sequence[int] = (64, 64), sigma_min: Sequence[float] = (1e-3, 1e-3),
                 sigma_max: Sequence[float] = (120, 160), s_churn: Sequence[float] = (3, 0),
                 use_cuda_graph: bool = False, noise_fn: Optional[Callable] = None):
        stages = len(models)
        assert stages > 0
        given = dict(model_kwargs_key_filter=model_kwargs_key_filter, guidance_scale=guidance_scale, use_karras=use_karras,
                     karras_steps=karras_steps, sigma_min=sigma_min, sigma_max=sigma_max, s_churn=s_churn)
        for name in _STAGE_OPTIONS:
            value = given[name]
            if stages > 1 and len(value) == 1:
                # one entry for a cascade: later stages repeat it -- but are unguided (scale 1) by default
                rest = [1.0] * (stages - 1) if name == "guidance_scale" else list(value) * (stages - 1)
                value = list(value) + rest
            if name == "model_kwargs_key_filter" and len(value) == 0:
                value = ["*"] * stages
            assert len(value) == stages, f"{name}: expected one entry per stage ({stages}), got {len(value)}"
            setattr(self, name, value)
        self.device, self.models, self.diffusions = device, models, diffusions
        self.num_points, self.aux_channels, self.clip_denoised = num_points, aux_channels, clip_denoised
        self.use_cuda_graph, self.noise_fn = use_cuda_graph, noise_fn
        self._graphs: Dict[Tuple, _GraphedStage] = {}

    @property
    def num_stages(self) -> int:
        return len(self.models)

    def _stages(self) -> Iterator[_Stage]:
        for i in range(self.num_stages):
            yield _Stage(i, self.models[i], self.diffusions[i], self.num_points[i], self.model_kwargs_key_filter[i],
                         self.guidance_scale[i], self.use_karras[i], self.karras_steps[i], self.sigma_min[i],
                         self.sigma_max[i], self.s_churn[i])

    def sample_batch(self, batch_size: int, model_kwargs: Dict[str, Any]) -> torch.Tensor:
        """The final clouds only: the last yield of ``sample_batch_progressive``."""
        last = None
        for last in self.sample_batch_progressive(batch_size, model_kwargs):
            pass
        return last

    def _graphed_stage(self, st: _Stage, shape) -> _GraphedStage:
        base = (st.index, tuple(shape), st.steps, st.sigma_min, st.sigma_max, st.churn, st.guidance, self.clip_denoised)
        # a graph bakes in the addresses of the packed weights / time tokens: re-capture when the model re-packs them
        # (load_state_dict, an optimiser step) and drop the stale capture
        key = base + (st.model.graph_key() if hasattr(st.model, "graph_key") else None,)
        for old in [k for k in self._graphs if k[:len(base)] == base and k != key]:
            del self._graphs[old]
        if key not in self._graphs:
            plan = HeunPlan(st.diffusion, st.steps, st.sigma_min, st.sigma_max, 7.0, st.churn)
            self._graphs[key] = _GraphedStage(st.model, st.diffusion, plan, shape, self.device, st.guidance,
                                              self.clip_denoised)
        return self._graphs[key]

    def _stage_conditioning(self, st: _Stage, batch_size: int, model_kwargs: Dict[str, Any], previous) -> Dict[str, Any]:
        """kwargs of one stage (reference sampler.py:121-136): key filter, the previous stage's clouds as ``low_res``,
        the model's ``cached_model_kwargs`` hook, then -- for classifier-free guidance -- every tensor doubled with an
        all-zero unconditional half ([:B] conditional, [B:] unconditional)."""
        kw = dict(model_kwargs)
        if st.key_filter != "*":
            wanted = set(st.key_filter.split(","))
            kw = {name: v for name, v in kw.items() if name in wanted}
        if previous is not None:
            kw["low_res"] = previous
        if hasattr(st.model, "cached_model_kwargs"):
            kw = st.model.cached_model_kwargs(batch_size, kw)
        if st.guided:
            kw = {name: (v if name == "prev_latent" else torch.cat([v, torch.zeros_like(v)], dim=0))
                  for name, v in kw.items()}
        return kw

    def _stage_predictions(self, st: _Stage, shape, kw: Dict[str, Any], x_target) -> Iterator[torch.Tensor]:
        """The x0 predictions a stage yields, one per step (+ the final one repeated), [B or 2B, C, N]."""
        if not st.karras:
            # ancestral DDPM loop over all diffusion steps (reference sampler.py:153-165).  With a guidance scale the
            # reference wraps the model in _uncond_guide_model, whose (x_t, ts, model_kwargs) signature does not
            # match how p_mean_variance calls models (**model_kwargs): that call raises TypeError there, so there
            # is no behaviour to reproduce.
            if st.guidance:
                raise NotImplementedError(
                    "use_karras=False with a guidance scale: the reference's own branch fails with a TypeError "
                    "(diffusion/sampler.py:194-233 vs gaussian_diffusion.py:285); use guidance_scale=0 or use_karras=True")
            for out in st.diffusion.p_sample_loop_progressive(st.model, shape=shape, model_kwargs=kw, device=self.device,
                                                              clip_denoised=self.clip_denoised, noise_fn=self.noise_fn):
                yield out["pred_xstart"]
        elif self.use_cuda_graph and getattr(st.model, "pcd_native", False):
            # fresh tensors like the reference (the stage buffers are overwritten by the next replay)
            preds = self._graphed_stage(st, shape).run(kw, self.noise_fn).clone()
            yield from preds
            yield preds[-1]
        else:
            for out in karras_sample_progressive(diffusion=st.diffusion, model=st.model, shape=shape, steps=st.steps,
                                                 clip_denoised=self.clip_denoised, model_kwargs=kw, device=self.device,
                                                 sigma_min=st.sigma_min, sigma_max=st.sigma_max, s_churn=st.churn,
                                                 guidance_scale=st.guidance, x_target=x_target, noise_fn=self.noise_fn):
                yield out["pred_xstart"]

    def sample_batch_progressive(self, batch_size: int, model_kwargs: Dict[str, Any],
                                 x_target: torch.Tensor = None) -> Iterator[torch.Tensor]:
        """Yields [batch_size, 3 + len(aux_channels), N] after every sampling step of every stage; an upsampling
        stage's yields carry its conditioning cloud in front of the new points (reference sampler.py:166-171)."""
        require_cuda()
        clouds = None
        for st in self._stages():
            kw = self._stage_conditioning(st, batch_size, model_kwargs, clouds)
            shape = (batch_size, 3 + len(self.aux_channels), st.num_points)
            low_res = kw.get("low_res")
            for pred in self._stage_predictions(st, shape, kw, x_target):
                clouds = pred[:batch_size]
                if low_res is not None:
                    clouds = torch.cat([low_res[:len(clouds)], clouds], dim=-1)
                yield clouds

    def _options(self) -> Dict[str, Any]:
        return dict({name: getattr(self, name) for name in _STAGE_OPTIONS}, device=self.device, models=self.models,
                    diffusions=self.diffusions, num_points=self.num_points, aux_channels=self.aux_channels,
                    clip_denoised=self.clip_denoised, use_cuda_graph=self.use_cuda_graph, noise_fn=self.noise_fn)

    @classmethod
    def combine(cls, *samplers: "PointCloudSampler") -> "PointCloudSampler":
        """One cascade out of several samplers, stages in argument order (reference sampler.py:173-192)."""
        first = samplers[0]
        for other in samplers[1:]:
            assert other.device == first.device
            assert other.aux_channels == first.aux_channels
            assert other.clip_denoised == first.clip_denoised
        merged = first._options()
        for name in ("models", "diffusions", "num_points") + _STAGE_OPTIONS:
            merged[name] = [entry for smp in samplers for entry in getattr(smp, name)]
        return cls(**merged)

    def with_options(self, guidance_scale: float, clip_denoised: bool, use_karras: Sequence[bool] = (True, True),
                     karras_steps: Sequence[int] = (64, 64), sigma_min: Sequence[float] = (1e-3, 1e-3),
                     sigma_max: Sequence[float] = (120, 160), s_churn: Sequence[float] = (3, 0)) -> "PointCloudSampler":
        """A sampler over the same models with other sampling options (reference sampler.py:267-291)."""
        opts = self._options()
        opts.update(guidance_scale=guidance_scale, clip_denoised=clip_denoised, use_karras=use_karras,
                    karras_steps=karras_steps, sigma_min=sigma_min, sigma_max=sigma_max, s_churn=s_churn)
        return PointCloudSampler(**opts)

    def split_model_output(self, output: torch.Tensor,
                           rescale_colors: bool = False) -> Tuple[torch.Tensor, Dict[str, torch.Tensor]]:
        """[B, 3 + aux, N] -> (xyz [B, 3, N], {channel name: [B, N]}); colour channels are clamped to [0, 255] and
        rounded, optionally rescaled to [0, 1] (reference sampler.py:235-253)."""
        assert output.shape[1] == 3 + len(self.aux_channels), "there must be three spatial channels before aux"
        named = {}
        for offset, name in enumerate(self.aux_channels):
            values = output[:, 3 + offset]
            if name in _COLOUR_CHANNELS:
                values = values.clamp(0, 255).round()
                values = values / 255.0 if rescale_colors else values
            named[name] = values
        return output[:, :3], named

    def output_to_point_clouds(self, output: torch.Tensor) -> List[PointCloud]:
        """One ``PointCloud`` per sample, colours rescaled to [0, 1] (reference sampler.py:255-265)."""
        xyz, named = self.split_model_output(output, rescale_colors=True)
        xyz = xyz.permute(0, 2, 1).cpu().numpy()
        named = {name: v.cpu().numpy() for name, v in named.items()}
        return [PointCloud(coords=xyz[b], channels={name: v[b] for name, v in named.items()})
                for b in range(output.shape[0])]
