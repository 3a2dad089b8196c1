#!/usr/bin/env python
"""Headline benchmark: completed clouds/s of the full 64-step Karras/Heun sampler.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference|eager-gpu] [--workload NAME]

A "step" is one complete sampling pass over one batch: 64 Heun steps = 127 denoiser
evaluations (each a 2B-sequence classifier-free-guidance forward) + the fused sampler
updates, replayed from one CUDA graph.  Workload at N=1 is BASELINE.json configs[1]:
base40M-imagevec (width 512, 12 layers, L = 1026), 1024 points, batch 64 per GPU, bf16,
guidance 3, s_churn 3, synthetic unit-norm CLIP embeddings, reference-init weights.
Multi-GPU (torchrun): one rank per GPU, batch sharded through the product API
(``pcd_b200.dist.sample_sharded``), no per-step collective, one NCCL all-gather of the
finished clouds per step.

The headline line also carries ``other_workloads``: BASELINE.json configs[2..4] at their
stated batches (text-vector model, upsampler with global batch 128 strong-scaled over the
ranks, 300M image + partial-cloud completion at 64 clouds per GPU) and the perceiver
cross-attention block at the text-conditioning shape, each measured in the same run
(one warm pass + one timed pass).
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "completed clouds/sec (64-step Heun, 1024 pts)"
UNIT = "clouds/s"
HEADLINE = "base40M-imagevec-1024pt-b64"
WORKLOADS = {
    # BASELINE.json configs[1] -- the configuration the headline metric is quoted on: 64 clouds per GPU (weak scaling)
    "base40M-imagevec-1024pt-b64": dict(model="base40M-imagevec", diffusion="base40M-imagevec", batch=64, scaling="weak",
                                        sigma_max=120.0, churn=3.0, guidance=3.0, baseline_config=1),
    # configs[2] (i): text-conditioned = the same network fed a CLIP text vector; batch 256 across 8 GPUs = 32 per GPU
    "base40M-textvec-1024pt-b32": dict(model="base40M-textvec", diffusion="base40M-textvec", batch=32, scaling="weak",
                                       sigma_max=120.0, churn=3.0, guidance=3.0, baseline_config=2),
    # configs[3]: upsampler 1024 -> 4096 points (L = 4353), GLOBAL batch 128 at 1/2/4/8 GPUs (strong scaling),
    # unguided like the stage-2 default of the reference's text2pc flow (127 evaluations)
    "upsample-4096pt-b128": dict(model="upsample", diffusion="upsample", batch=128, scaling="strong",
                                 sigma_max=160.0, churn=0.0, guidance=0.0, baseline_config=3),
    # configs[4]: 300M denoiser (width 1024, 24 layers) with image grid + partial-cloud conditioning, 4096 points,
    # batch 512 across 8 GPUs = 64 per GPU, guided
    "base300M-upsample-4096pt-b64": dict(model="base300M-upsample", diffusion="upsample", batch=64, scaling="weak",
                                         sigma_max=160.0, churn=0.0, guidance=3.0, baseline_config=4),
    # the TwoStreamDenoiser of the reference's config.yaml (all four modalities, 57.5 M parameters)
    # under its own sampling settings (config.yaml:40-58: 32 samples, guidance 3, 64 steps, s_churn 0)
    "twostream-config-1024pt-b32": dict(model="twostream", diffusion="linear-1000", batch=32, scaling="weak",
                                        sigma_max=120.0, churn=0.0, guidance=3.0, baseline_config=None),
}
OTHER_WORKLOADS = ["base40M-textvec-1024pt-b32", "upsample-4096pt-b128", "base300M-upsample-4096pt-b64"]
TWOSTREAM_CONFIG = dict(num_points=1024, num_latents=256, input_channels=3, output_channels=3, latent_dim=256, x_dim=256,
                        num_blocks=6, num_compute_layers=4, num_heads=8, num_classes=10, num_tokens_ppcd=256,
                        num_tokens_depth=128, active_modalities=["class", "view", "partial_pcd", "depth"])
# BASELINE config 3 (ii): cross-attention block at the text-conditioning shape
PERCEIVER_TEXT = dict(n_q=1026, n_data=77, width=512, heads=8, data_width=768, layers=1)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback (B200_PROFILING.md)")


def model_cfg(P_or_cases_configs, name):
    """Constructor config of a workload's model (grid-upsample class with base300M dims)."""
    if name == "base300M-upsample":
        return dict(P_or_cases_configs["upsample"], width=1024, layers=24, heads=16)
    return dict(P_or_cases_configs[name])


def local_batch(w, world, rank):
    """Clouds this rank samples: weak scaling keeps the per-GPU batch, strong scaling shards the global batch."""
    if w["scaling"] == "weak":
        return w["batch"], world * w["batch"]
    base, rem = divmod(w["batch"], world)
    return base + (1 if rank < rem else 0), w["batch"]


def workload_config(name, world):
    """The `config` object of a bench line; identical for the b200 and the reference arm."""
    w = WORKLOADS[name]
    guided = w["guidance"] not in (0.0, 1.0)
    per_gpu, glob = local_batch(w, world, 0)
    return {"workload": name, "baseline_config": w["baseline_config"], "per_gpu_batch": per_gpu, "global_batch": glob,
            "scaling": w["scaling"], "heun_steps": 64, "denoiser_evals_per_step": 127,
            "forwards_per_eval": 2 if guided else 1, "guidance": w["guidance"], "s_churn": w["churn"],
            "sigma_max": w["sigma_max"], "parallelism": f"batch-sharded x{world}, all-gather of finished clouds"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.lines, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------
# CPU baseline: the oracle port of the reference's path on the host cores
# ---------------------------------------------------------------------------
def cpu_conditioning(cfg, B, guided):
    """Synthetic conditioning of the oracle model (deterministic), doubled for guidance like sampler.py:133-136."""
    import torch

    from oracle import det
    kw = {}
    cls = cfg["name"]
    if cls == "CLIPImagePointDiffusionTransformer":
        e = det.normal((B, 768), 77)
        kw["embeddings"] = e / e.norm(dim=1, keepdim=True)
    if "Grid" in cls:
        kw["embeddings"] = det.normal((B, 1024, 256), 78)
    if "Upsample" in cls:
        lr = det.uniform((B, cfg["input_channels"], cfg["cond_ctx"]), 79, std=0.5 / 3 ** 0.5)
        lr[:, 3:] = (lr[:, 3:] + 0.5) * 255.0
        kw["low_res"] = lr
    if guided:
        kw = {k: torch.cat([v, torch.zeros_like(v)], 0) for k, v in kw.items()}
    return kw


def cpu_sample(workload, heun_steps=2, batch=None, threads=None):
    """Time `heun_steps` Heun steps of the ORACLE sampler (reference algorithm, fp32 torch CPU, every host thread)
    for a small batch of the workload and scale to clouds/s of the 64-step sampler.  The first evaluation is
    untimed (thread-pool / allocator warm-up).  Returns (cpu_baseline dict, timed seconds)."""
    import torch

    from oracle import cases, det
    from oracle import sampler as S
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_oracle_golden import shapes_of

    w = WORKLOADS[workload]
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    cfg = model_cfg(cases.MODEL_CONFIGS, w["model"])
    heavy = cfg["width"] > 512 or cfg["n_ctx"] > 1024
    B = batch or (1 if heavy else 8)
    sd = det.fill_state_dict(shapes_of(cfg), 201, mode="reference", width=cfg["width"])
    tab = S.Tables(**cases.DIFFUSION_CONFIGS["upsample" if "upsample" in w["diffusion"] else "base"])
    guided = w["guidance"] not in (0.0, 1.0)
    kw = cpu_conditioning(cfg, B, guided)
    gen = S.heun_progressive(S.make_model_fn(sd, cfg), tab, (B, cfg["input_channels"], cfg["n_ctx"]), steps=64,
                             sigma_min=1e-3, sigma_max=w["sigma_max"], s_churn=w["churn"], guidance_scale=w["guidance"],
                             model_kwargs=kw, noise_fn=cases.DetNoise(5))
    fwd_per_step = 2 * (2 if guided else 1)
    with torch.no_grad():
        next(gen)  # first yield comes after the first evaluation of step 0: start the clock there
        t0 = time.perf_counter()
        for _ in range(heun_steps):
            next(gen)
        dt = time.perf_counter() - t0
    per_batch = dt / heun_steps * 64.0
    return dict(value=B / per_batch, unit=UNIT, cores=threads, kind="port",
                sample=f"oracle (reference algorithm, fp32 torch-CPU, {threads} threads) on {B} cloud(s) of {workload}: "
                       f"{heun_steps} of 64 Heun steps ({fwd_per_step * heun_steps} B={B} denoiser forwards) in {dt:.1f}s, "
                       f"scaled x{64 / heun_steps:g}"), dt


def run_reference(args):
    """Reference arm: the reference's own CPU implementation of the path (oracle port: the reference is pure Python
    and cannot travel to the GPU box) on the host cores; each step = a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    times, base = [], None
    for i in range(args.warmup + args.steps):
        base, dt = cpu_sample(args.workload, heun_steps=1, batch=None if args.workload != HEADLINE else 4)
        if i >= args.warmup:
            times.append(dt)
    B = 4 if args.workload == HEADLINE else int(base["sample"].split(" on ")[1].split(" ")[0])
    v = B / ((sum(times) / len(times)) * 64.0)
    base["value"] = v
    base["sample"] += f"; mean of {len(times)} such samples"
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times),
            "higher_is_better": True, "scaling": WORKLOADS[args.workload]["scaling"], "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(args.workload, world),
            "note": "each step = bounded CPU sample (1 of 64 Heun steps of a small batch); value scaled to the full sampler",
            "cpu_baseline": base,
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ---------------------------------------------------------------------------
# Same-box comparators: the reference algorithm (oracle port, plain PyTorch ops) on the GPU -- as
# shipped (fp32, [B,H,L,L] attention materialised), under bf16 autocast, and the LIBRARY bar: bf16 autocast with
# F.scaled_dot_product_attention (cuDNN / flash backends) and cuBLAS GEMMs.  None of this repo's kernels run here.
# ---------------------------------------------------------------------------
def _sdpa_qkv_attention(qkv, heads):
    """oracle.denoiser.qkv_attention (transformer.py:65-84) through F.scaled_dot_product_attention."""
    import torch
    import torch.nn.functional as F
    bs, n_ctx, width = qkv.shape
    hd = width // heads // 3
    q, k, v = qkv.view(bs, n_ctx, heads, 3, hd).permute(3, 0, 2, 1, 4)   # [3][B, H, L, hd]
    out = F.scaled_dot_product_attention(q, k, v, scale=1.0 / math.sqrt(hd))
    return out.transpose(1, 2).reshape(bs, n_ctx, heads * hd)


def run_eager_gpu(args):
    import torch

    from oracle import cases, det
    from oracle import denoiser as D
    from oracle import sampler as S
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_oracle_golden import shapes_of

    if int(os.environ.get("RANK", "0")) != 0:
        return
    w = WORKLOADS[args.workload]
    if w["model"] == "twostream":
        return run_eager_gpu_twostream(args)
    B = args.batch or w["batch"]
    dev = torch.device("cuda:0")
    torch.backends.cuda.matmul.allow_tf32 = False
    cfg = model_cfg(cases.MODEL_CONFIGS, w["model"])
    sd = {k: v.to(dev) for k, v in det.fill_state_dict(shapes_of(cfg), 201, mode="reference", width=cfg["width"]).items()}
    tab = S.Tables(**cases.DIFFUSION_CONFIGS["upsample" if "upsample" in w["diffusion"] else "base"])
    guided = w["guidance"] not in (0.0, 1.0)
    kw = {k: v.to(dev) for k, v in cpu_conditioning(cfg, B, guided).items()}
    g = torch.Generator(device=dev).manual_seed(1234)
    heun_steps = max(1, args.steps)
    out = {}
    plain_attention = D.qkv_attention
    modes = ("fp32", "bf16-autocast", "bf16-autocast+sdpa", "bf16-autocast+sdpa+compile")
    for mode in modes:
        base_fn = S.make_model_fn(sd, cfg)
        D.qkv_attention = _sdpa_qkv_attention if "sdpa" in mode else plain_attention
        if "compile" in mode:
            try:
                base_fn = torch.compile(base_fn, dynamic=False)
            except Exception as ex:
                out[mode] = f"unavailable: {ex!r}"[:120]
                continue
        if mode == "fp32":
            fn = base_fn
        else:
            def fn(x, t, _f=base_fn, **k):
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    return _f(x, t, **k).float()
        try:
            gen = S.heun_progressive(fn, tab, (B, cfg["input_channels"], cfg["n_ctx"]), steps=64, sigma_min=1e-3,
                                     sigma_max=w["sigma_max"], s_churn=w["churn"], guidance_scale=w["guidance"],
                                     model_kwargs=kw, noise_fn=lambda shp: torch.randn(*shp, device=dev, generator=g))
            with torch.no_grad():
                for _ in range(1 + args.warmup):
                    next(gen)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for _ in range(heun_steps):
                    next(gen)
                torch.cuda.synchronize()
                dt = time.perf_counter() - t0
            out[mode] = B / (dt / heun_steps * 64.0)
            del gen
        except Exception as ex:  # e.g. torch.compile without a working inductor toolchain on the box
            out[mode] = f"unavailable: {ex!r}"[:160]
        finally:
            D.qkv_attention = plain_attention
        torch.cuda.empty_cache()
    best = max(v for v in out.values() if isinstance(v, float))
    print(json.dumps({"impl": "eager-gpu", "metric": METRIC, "unit": UNIT, "value": best, "modes": out, "n_gpus": 1,
                      "config": {"workload": args.workload, "batch": B,
                                 "note": f"reference algorithm as plain PyTorch ops on cuda:0 (oracle port; one kernel per op); "
                                         f"'sdpa' = F.scaled_dot_product_attention (library flash / cuDNN attention) + cuBLAS bf16 "
                                         f"GEMMs; {heun_steps} of 64 Heun steps timed by wall clock around synchronize, scaled to 64"}}))


def run_eager_gpu_twostream(args):
    """The reference's TwoStreamDenoiser algorithm (oracle restatement, plain PyTorch ops, condition encoders re-run
    in every forward like model.py:498-509) under the oracle's guided Heun loop on cuda:0, fp32 and bf16 autocast."""
    import torch

    import pcd_b200 as P
    from oracle import sampler as S
    from oracle import twostream as OT
    w = WORKLOADS[args.workload]
    B = args.batch or w["batch"]
    dev = torch.device("cuda:0")
    torch.backends.cuda.matmul.allow_tf32 = False
    c = TWOSTREAM_CONFIG
    torch.manual_seed(1234)
    shell = P.TwoStreamDenoiser(**c, device=dev, dtype=torch.float32)  # parameter container only: reference-shaped state_dict
    sd = {k: v.detach() for k, v in shell.state_dict().items()}
    tab = S.Tables(schedule="linear", timesteps=1000)
    kw = dict(class_labels=torch.randint(1, c["num_classes"], (B,), device=dev), viewpoints=torch.rand(B, 3, device=dev),
              partial_pcd=torch.rand(B, 1024, 3, device=dev) - 0.5, depth_maps=torch.rand(B, 1, 512, 512, device=dev))
    kw = {k: torch.cat([v, torch.zeros_like(v)]) for k, v in kw.items()}
    g = torch.Generator(device=dev).manual_seed(1)
    heun_steps = max(1, args.steps)
    out = {}
    for mode in ("fp32", "bf16-autocast"):
        def fn(x, t, _m=mode, **k):
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=(_m != "fp32")):
                y, z = OT.twostream_forward(sd, c, x, t, k.get("class_labels"), k.get("viewpoints"), k.get("prev_latent"),
                                            partial_pcd=k.get("partial_pcd"), depth_maps=k.get("depth_maps"))
            return y.float(), z.float()
        gen = S.heun_progressive(fn, tab, (B, 3, c["num_points"]), steps=64, sigma_min=1e-3, sigma_max=w["sigma_max"],
                                 s_churn=w["churn"], guidance_scale=w["guidance"], model_kwargs=kw,
                                 noise_fn=lambda shp: torch.randn(*shp, device=dev, generator=g))
        with torch.no_grad():
            for _ in range(1 + args.warmup):
                next(gen)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(heun_steps):
                next(gen)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        out[mode] = B / (dt / heun_steps * 64.0)
        del gen
        torch.cuda.empty_cache()
    print(json.dumps({"impl": "eager-gpu", "metric": METRIC, "unit": UNIT, "value": out["bf16-autocast"],
                      "fp32": out["fp32"], "bf16_autocast": out["bf16-autocast"], "n_gpus": 1,
                      "config": {"workload": args.workload, "batch": B,
                                 "note": f"reference algorithm as plain PyTorch ops on cuda:0 (oracle port, condition encoders "
                                         f"re-run every forward as in the reference); {heun_steps} of 64 Heun steps timed by "
                                         f"wall clock around synchronize, scaled to 64"}}))


# ---------------------------------------------------------------------------
# TwoStreamDenoiser (config.yaml shapes) under the sampler, single GPU
# ---------------------------------------------------------------------------
def twostream_flops_per_forward(c):
    """Multiply-add = 2 FLOP, one sequence, backbone only (the condition encoders run once per batch)."""
    d, N = c["latent_dim"], c["num_points"]
    nl = c["num_latents"] + 2 + c["num_tokens_ppcd"] + c["num_tokens_depth"] + 1
    attn = lambda lq, lk: 4 * lq * d * d + 4 * lk * d * d + 4 * lq * lk * d
    mlp = lambda l: 16 * l * d * d
    block = attn(nl, N) + mlp(nl) + c["num_compute_layers"] * (attn(nl, nl) + mlp(nl)) + attn(N, nl) + mlp(N)
    return c["num_blocks"] * block + mlp(nl)


def run_twostream(args):
    import torch

    import pcd_b200 as P
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    assert int(os.environ.get("WORLD_SIZE", "1")) == 1, "single-GPU workload"
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    lib = P._lib.load()
    w = WORKLOADS[args.workload]
    B = args.batch or w["batch"]
    smax, churn, guidance = w["sigma_max"], w["churn"], w["guidance"]
    c = TWOSTREAM_CONFIG
    torch.manual_seed(1234)
    model = P.TwoStreamDenoiser(**c, device=dev, dtype=torch.bfloat16)
    with torch.no_grad():  # the reference zero-initialises ln_latent; a trained model does not keep it there
        model.denoiser_backbone.ln_latent.weight.fill_(1.0)
    diffusion = P.GaussianDiffusion(betas=P.get_named_beta_schedule("linear", 1000), model_mean_type="epsilon",
                                    model_var_type="fixed_small", loss_type="mse")
    mk = lambda m, graph: P.PointCloudSampler(dev, [m], [diffusion], [c["num_points"]], [], guidance_scale=[guidance],
                                              use_karras=[True], karras_steps=[64], sigma_min=[1e-3], sigma_max=[smax],
                                              s_churn=[churn], use_cuda_graph=graph)
    sampler = mk(model, True)
    host_kw = dict(class_labels=torch.randint(1, c["num_classes"], (B,)), viewpoints=torch.rand(B, 3),
                   partial_pcd=torch.rand(B, 1024, 3) - 0.5, depth_maps=torch.rand(B, 1, 512, 512))
    host_kw = {k: v.pin_memory() for k, v in host_kw.items()}
    dev_kw = {k: v.to(dev) for k, v in host_kw.items()}
    out_host = torch.empty(B, 3, c["num_points"]).pin_memory()

    def step_device():
        model._cond_cache.clear()  # every batch pays for its condition encoders once
        return sampler.sample_batch(B, dict(dev_kw))

    def step_e2e():
        kw = {k: v.to(dev, non_blocking=True) for k, v in host_kw.items()}
        out_host.copy_(sampler.sample_batch(B, kw))
        return out_host

    def timed(fn, k):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / k

    for _ in range(max(1, args.warmup)):
        step_device()
    clocks = ClockSampler(0)
    clocks.start()
    ms = timed(step_device, args.steps)
    clk = clocks.stop()
    stage = next(iter(sampler._graphs.values()))
    c0 = lib.pcd_launch_count()
    stage._enqueue()
    torch.cuda.synchronize()
    launches = int(lib.pcd_launch_count() - c0)
    step_e2e()
    ms_e2e = timed(step_e2e, args.steps)

    class Opaque(torch.nn.Module):  # public forward only: two B-sized calls per evaluation, no graph
        def __init__(self, inner):
            super().__init__()
            self.inner = inner

        def forward(self, x, t, **kw):
            return self.inner(x, t, **kw)
    plain = mk(Opaque(model), False)
    plain.sample_batch(2, {k: v[:2] for k, v in dev_kw.items()})
    ms_plain = timed(lambda: plain.sample_batch(B, dict(dev_kw)), 1)

    pk = peaks()
    evals = 127 * 2
    flops = twostream_flops_per_forward(c) * evals * B
    tf = flops / (ms * 1e-3) / 1e12
    h2d = sum(v.numel() * v.element_size() for v in host_kw.values())
    print(json.dumps({
        "metric": METRIC, "value": B / (ms * 1e-3), "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": args.workload, "batch": B, "model": "TwoStreamDenoiser, reference config.yaml:24-38",
                   "sampler": "64-step Heun, guidance 3, s_churn 0, one CUDA graph per stage",
                   "l2": "activations of one evaluation (>1 GB) exceed L2"},
        "e2e": {"value": B / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": out_host.numel() * 4},
        "gpu_launches": launches * args.steps, "launches_per_step": launches,
        "public_forward_two_calls_no_graph": {"value": B / (ms_plain * 1e-3), "unit": UNIT},
        "roofline": {"kernel": "whole step (backbone projections + attention)", "bound": "tensor", "achieved": tf,
                     "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": tf / pk["tf_sustained"], "traffic": None,
                     "peak_source": pk["source"] + ", sustained",
                     "note": "useful FLOP only (the zero-filled upper halves of the 32-wide head tiles are not counted)"},
        "clocks": clk}))


# ---------------------------------------------------------------------------
def time_kernel(fn, min_seconds=1.0, warmup=3):
    """Mean launch time over a >= `min_seconds` back-to-back run (CUDA events on the launching stream): long enough
    for the chip to settle at the power-capped clocks it also runs the step at, so the SUSTAINED peak applies."""
    import torch
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        fn()
    e1.record()
    torch.cuda.synchronize()
    est = max(e0.elapsed_time(e1) / 5 * 1e-3, 1e-6)
    iters = int(min(max(min_seconds / est, 10), 20000))
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3  # seconds per launch


def kernel_breakdown(P, B2, L, width, heads, layers, evals, Bc, C, N, pk, min_seconds=1.0):
    """Live CUDA-event timing of each hot kernel at the workload's shapes (operands larger than L2) ->
    roofline fractions against the SUSTAINED measured peaks."""
    import torch
    dev = torch.device("cuda")
    M = B2 * L
    bf = torch.bfloat16
    g = torch.Generator(device=dev).manual_seed(1)
    a_d = torch.randn(M, width, device=dev, generator=g).to(bf)
    a_4d = torch.randn(M, 4 * width, device=dev, generator=g).to(bf)
    h = torch.randn(M, width, device=dev, generator=g)
    mk = lambda n, k: (torch.randn(n, k, device=dev, generator=g) / math.sqrt(k)).to(bf)
    w_qkv, w_proj, w_fc, w_fc2 = mk(3 * width, width), mk(width, width), mk(4 * width, width), mk(width, 4 * width)
    bias = lambda n: torch.zeros(n, device=dev)
    qkv = torch.empty(M, 3 * width, device=dev, dtype=bf)
    hid = torch.empty(M, 4 * width, device=dev, dtype=bf)
    ops = P.ops
    res = {}

    def add(name, fn, launches, flops=None, bytes_=None):
        t = time_kernel(fn, min_seconds if launches else 0.2)
        r = {"ms": t * 1e3, "launches_per_step": launches, "ms_per_step": t * 1e3 * launches}
        if flops is not None:
            r.update(bound="tensor", achieved=flops / t / 1e12, peak=pk["tf_sustained"], unit="TFLOP/s")
        else:
            r.update(bound="hbm", achieved=bytes_ / t / 1e9, peak=pk["hbm"], unit="GB/s")
        r["frac"] = r["achieved"] / r["peak"]
        res[name] = r

    per = evals * layers
    y = torch.empty(M, width, device=dev, dtype=bf)
    fold = width % 256 == 0 and M >= 512   # the LayerNorm-folded forward (pcd_model_forward's default path)
    if fold:
        hb, stats = ops.cast_rowstats(h)
        colsum = lambda w: w.float().sum(dim=1).contiguous()
        cs_qkv, cs_fc = colsum(w_qkv), colsum(w_fc)
        add("gemm_qkv_lnfold", lambda: ops.linear_layernorm_folded(hb, stats, w_qkv, cs_qkv, bias(3 * width)), per,
            flops=2.0 * M * 3 * width * width)
        # residual update + bf16 copy + row statistics in the epilogue: K = width makes this one HBM-bound
        # (A 2K/N + h read 4 + h write 4 + bf16 copy 2 bytes per output element)
        add("gemm_attn_proj_resid_stats", lambda: ops.linear_residual_stats(a_d, w_proj, bias(width), h), per,
            bytes_=M * width * 12.0)
        res["gemm_attn_proj_resid_stats"]["tflops"] = 2.0 * M * width * width / (res["gemm_attn_proj_resid_stats"]["ms"] * 1e-3) / 1e12
        add("gemm_fc1_lnfold_gelu", lambda: ops.linear_layernorm_folded(hb, stats, w_fc, cs_fc, bias(4 * width), gelu=True),
            per, flops=2.0 * M * 4 * width * width)
        add("gemm_fc2_resid_stats", lambda: ops.linear_residual_stats(a_4d, w_fc2, bias(width), h), per,
            flops=2.0 * M * 4 * width * width)
        res["gemm_fc2_resid_stats"]["hbm_gbs"] = M * width * 18.0 / (res["gemm_fc2_resid_stats"]["ms"] * 1e-3) / 1e9
        # token assembly + ln_pre (emits the fp32 stream, its bf16 copy and the row statistics: 6 bytes per element)
        # and ln_post + output projection (reads the fp32 stream of the point tokens), once per evaluation
        n_pre = L - N
        x_in = torch.randn(Bc, C, N, device=dev, generator=g)   # guided: both halves of the 2B batch share x
        w_in = torch.randn(width, C, device=dev, generator=g)
        pre = torch.randn(B2, n_pre, width, device=dev, generator=g) if n_pre else None
        ones, zeros = torch.ones(width, device=dev), torch.zeros(width, device=dev)
        add("embed_tokens_ln_stats", lambda: ops.embed_tokens(x_in, w_in, zeros, pre, None, ones, zeros, seqs=B2, with_stats=True),
            evals, bytes_=M * width * 6.0)
        w_out = torch.randn(C, width, device=dev, generator=g)
        h3 = h.view(B2, L, width)
        add("ln_output_proj", lambda: ops.output_proj(h3, n_pre, ones, zeros, w_out, zeros[:C]), evals,
            bytes_=B2 * N * width * 4.0)
    else:
        add("gemm_qkv", lambda: ops.linear(a_d, w_qkv, bias(3 * width), out=qkv), per, flops=2.0 * M * 3 * width * width)
        add("gemm_attn_proj", lambda: ops.linear(a_d, w_proj, bias(width), out=y), per, flops=2.0 * M * width * width)
        add("gemm_fc1_gelu", lambda: ops.linear(a_d, w_fc, bias(4 * width), epilogue=1, out=hid), per,
            flops=2.0 * M * 4 * width * width)
        add("gemm_fc2", lambda: ops.linear(a_4d, w_fc2, bias(width), out=y), per, flops=2.0 * M * 4 * width * width)
        lnw, lnb = torch.ones(width, device=dev), torch.zeros(width, device=dev)
        y.normal_()
        add("add_layernorm", lambda: ops.add_layernorm(h, y, lnw, lnb, out_dtype=bf), 2 * per, bytes_=M * width * 12.0)
    qkv3 = qkv.view(B2, L, 3 * width)
    qkv3.normal_()
    add("flash_attention", lambda: ops.self_attention(qkv3, heads), per, flops=4.0 * L * L * 64 * heads * B2)
    # fused sampler update (state is tiny at the configured batch: L2-resident, launch bound)
    d = P.diffusion_from_config(P.DIFFUSION_CONFIGS["base40M"])
    plan = P.HeunPlan(d, 64, 1e-3, 120.0, 7.0, 3.0)
    for label, b in (("sampler_update_cfg_batch", Bc), ("sampler_update_large_batch", 4096)):
        st = P.k_diffusion.HeunState(d, plan, (b, C, N), dev, 3.0, True)
        st.x.normal_()
        mo = torch.randn(2 * b, C, N, device=dev)
        nz = torch.randn(b, C, N, device=dev)
        pred = torch.empty(b, C, N, device=dev)
        st.begin(nz)

        def upd():
            st.predictor(0, mo, pred)
            st.corrector(0, mo, nz)
        n_el = b * C * N * 4.0
        # algorithmic bytes per Heun step: predictor r(x,out_c,out_u) w(d,model_in,pred) +
        # corrector r(x,d,out_c,out_u,noise) w(x,model_in) = 13 array passes
        add(label, upd, 64 if b == Bc else 0, bytes_=13.0 * n_el)
    return res


def perceiver_bench(P, dev, batch, pk):
    """BASELINE config 3 (ii): one ResidualCrossAttentionBlock at the text-conditioning shape (queries = the 1026
    denoiser tokens of 2*batch guided sequences, data = 77 CLIP text tokens of width 768), as a sampler would run it:
    the key / value side cached per conditioning tensor, five launches per evaluation."""
    import torch
    c = PERCEIVER_TEXT
    S = 2 * batch
    torch.manual_seed(7)
    per = P.perceiver.SimplePerceiver(device=dev, dtype=torch.bfloat16, n_data=c["n_data"], width=c["width"],
                                      layers=c["layers"], heads=c["heads"], data_width=c["data_width"])
    x = torch.randn(S, c["n_q"], c["width"], device=dev)
    data = torch.randn(S, c["n_data"], c["data_width"], device=dev)
    blk = per.resblocks[0]
    rows = S * c["n_q"]
    stream = x.clone().view(rows, c["width"])
    state = P.ops.cast_rowstats(stream)
    kv = blk.keys_values(data)
    lib = P._lib.load()
    n0 = lib.pcd_launch_count()
    blk.step(stream, state, kv, S)
    launches = int(lib.pcd_launch_count() - n0)

    def fn():
        blk.step(stream, state, kv, S)
    t = time_kernel(fn, 1.0)
    d, lq, lkv = c["width"], c["n_q"], c["n_data"]
    # 20 d^2 Lq (q, proj, 16 d^2 MLP) + 4 Lq Lkv d (attention) per sequence; the step-invariant
    # 4 d data_width Lkv of c_kv is paid once per conditioning tensor, outside this loop
    flops = S * (20.0 * d * d * lq + 4.0 * lq * lkv * d)
    t_kv = time_kernel(lambda: (blk._kv_cache.clear(), blk.keys_values(data)), 0.3)
    return {"workload": "perceiver-text-1026x77", "baseline_config": 2, "sequences": S,
            "shape": f"Lq={lq} Lkv={lkv} width={d} heads={c['heads']} data_width={c['data_width']}",
            "ms_per_block": t * 1e3, "launches_per_block": launches, "achieved": flops / t / 1e12, "unit": "TFLOP/s",
            "peak": pk["tf_sustained"], "frac": flops / t / 1e12 / pk["tf_sustained"],
            "kv_side_ms_once_per_conditioning": t_kv * 1e3,
            "note": "one ResidualCrossAttentionBlock evaluation on the bf16 LayerNorm-folded path; c_kv(ln_2(data)) cached"}


# ---------------------------------------------------------------------------
class Workload:
    """One sampler workload on this rank: model, sampler, synthetic conditioning (host + device), timing helpers."""

    def __init__(self, P, name, dev, world, rank, batch_override=0):
        import torch
        self.P, self.name, self.dev, self.world, self.rank = P, name, dev, world, rank
        w = self.w = WORKLOADS[name]
        self.B, self.global_batch = local_batch(w, world, rank)
        if batch_override:
            self.B = batch_override
            self.global_batch = batch_override * world
        self.cfg = cfg = model_cfg(P.MODEL_CONFIGS, w["model"])
        torch.manual_seed(1234)  # the same weights on every rank (replicated model)
        self.model = P.model_from_config(cfg, dev, dtype=torch.bfloat16)
        if hasattr(self.model, "accept_grid_embeddings"):
            self.model.accept_grid_embeddings = True
        with torch.no_grad():  # reference init + re-randomised output_proj
            self.model.output_proj.weight.normal_(std=0.02)
        self.diffusion = P.diffusion_from_config(P.DIFFUSION_CONFIGS[w["diffusion"]])
        self.C, self.N = cfg["input_channels"], cfg["n_ctx"]
        self.sampler = P.PointCloudSampler(dev, [self.model], [self.diffusion], [self.N], ["R", "G", "B"],
                                           guidance_scale=[w["guidance"]], use_karras=[True], karras_steps=[64],
                                           sigma_min=[1e-3], sigma_max=[w["sigma_max"]], s_churn=[w["churn"]],
                                           use_cuda_graph=True)
        # synthetic conditioning of THIS rank's shard: unit-norm CLIP vectors, N(0,1) CLIP grids,
        # U(-0.5,0.5) xyz + U(0,255) rgb partial clouds
        gen = torch.Generator().manual_seed(1234 + 7919 * rank)
        cls = cfg["name"]
        B = self.B
        host_kw = {}
        if cls == "CLIPImagePointDiffusionTransformer":
            e = torch.randn(B, 768, generator=gen)
            host_kw["embeddings"] = e / e.norm(dim=1, keepdim=True)
        if "Grid" in cls:
            host_kw["embeddings"] = torch.randn(B, 1024, 256, generator=gen)
        if "Upsample" in cls:
            lr = torch.rand(B, self.C, cfg["cond_ctx"], generator=gen) - 0.5
            lr[:, 3:] = (lr[:, 3:] + 0.5) * 255.0
            host_kw["low_res"] = lr
        self.host_kw = {k: v.pin_memory() for k, v in host_kw.items()}
        self.dev_kw = {k: v.to(dev) for k, v in self.host_kw.items()}
        self.n_out = self.N + (cfg["cond_ctx"] if "Upsample" in cls else 0)
        self.out_host = torch.empty(self.global_batch if world > 1 else B, self.C, self.n_out).pin_memory()
        self.guided = w["guidance"] not in (0.0, 1.0)
        self.seq_len = self.model._prefix_layout()[0] + self.N

    # the product's public multi-GPU entry: shard -> PointCloudSampler.sample_batch -> all-gather of the clouds
    def sample(self, kw):
        return self.P.dist.sample_sharded(self.sampler.sample_batch, self.global_batch, kw, kwargs_are_local=True,
                                          local_batch=self.B)

    def step_device(self):
        return self.sample(dict(self.dev_kw))

    def step_e2e(self, events=None):
        kw = {k: v.to(self.dev, non_blocking=True) for k, v in self.host_kw.items()}
        if events:
            events[0].record()
        y = self.sample(kw)
        if events:
            events[1].record()
        self.out_host.copy_(y, non_blocking=False)  # device -> host read of the step's result
        return self.out_host

    def launches_per_step(self):
        """Kernels of this library one sampling pass launches (counted while the stage's CUDA graph was captured)."""
        return sum(stage.launches_per_replay for stage in self.sampler._graphs.values())

    def bytes_per_step(self):
        return (sum(v.numel() * v.element_size() for v in self.host_kw.values()),
                self.out_host.numel() * self.out_host.element_size())

    def flops_per_cloud(self):
        """SURVEY 8d work model: 24 d^2 L layers + 4 L^2 d layers per forward (multiply-add = 2 FLOP)."""
        d, L, layers = self.cfg["width"], self.seq_len, self.cfg["layers"]
        return (24.0 * d * d * L * layers + 4.0 * L * L * d * layers) * 127 * (2 if self.guided else 1)

    def release(self):
        import torch
        self.sampler._graphs.clear()
        self.sampler = self.model = None
        self.dev_kw = self.host_kw = self.out_host = None
        torch.cuda.empty_cache()


def make_timers(torch, dist, world, dev):
    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) * 1e-3
    return barrier, timed


def measure_other(P, name, dev, world, rank, torch, dist, pk):
    """One warm pass (graph capture + replay) and ONE timed end-to-end pass of a non-headline workload; the device
    time of the sampling itself is taken with events inside that same pass."""
    t_wall = time.perf_counter()
    wl = Workload(P, name, dev, world, rank)
    barrier, _ = make_timers(torch, dist, world, dev)
    wl.step_device()
    barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ev[2].record()
    wl.step_e2e(events=ev[:2])
    ev[3].record()
    barrier()
    t = torch.tensor([ev[0].elapsed_time(ev[1]), ev[2].elapsed_time(ev[3])], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    t_dev, t_e2e = float(t[0]) * 1e-3, float(t[1]) * 1e-3
    launches = wl.launches_per_step()
    h2d, d2h = wl.bytes_per_step()
    tf = wl.flops_per_cloud() * wl.global_batch / t_dev / 1e12 / world
    row = {"workload": name, "baseline_config": wl.w["baseline_config"], "scaling": wl.w["scaling"],
           "per_gpu_batch": wl.B, "global_batch": wl.global_batch, "points": wl.n_out, "seq_len": wl.seq_len,
           "forwards_per_eval": 2 if wl.guided else 1, "value": wl.global_batch / t_dev, "unit": UNIT,
           "ms_per_step": t_dev * 1e3, "steps": 1, "warmup": 1,
           "e2e": {"value": wl.global_batch / t_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
           "gpu_launches": launches,
           "whole_step": {"achieved": tf, "unit": "TFLOP/s per GPU", "peak": pk["tf_sustained"],
                          "frac": tf / pk["tf_sustained"], "note": "SURVEY 8d work model FLOP / device time"}}
    if rank == 0:
        try:
            kb = kernel_breakdown(P, (2 if wl.guided else 1) * wl.B, wl.seq_len, wl.cfg["width"], wl.cfg["heads"],
                                  wl.cfg["layers"], 127, wl.B, wl.C, wl.N, pk, min_seconds=0.4)
            dom = max((k for k in kb if kb[k]["launches_per_step"]), key=lambda k: kb[k]["ms_per_step"])
            row["dominant_kernel"] = {"kernel": dom, **{k: (round(v, 4) if isinstance(v, float) else v)
                                                        for k, v in kb[dom].items()}}
            row["kernel_frac"] = {k: round(v["frac"], 4) for k, v in kb.items() if v["launches_per_step"]}
        except Exception as ex:
            row["dominant_kernel"] = {"error": repr(ex)}
    wl.release()
    row["bench_wall_s"] = round(time.perf_counter() - t_wall, 1)
    return row


def traffic_for(workload, kernel):
    """dram__bytes_read + dram__bytes_write of one launch of `kernel` in `workload`, from the committed ncu capture
    (profiles/top_kernel_traffic.json, written by tools/gpu_profile_bench.sh); None when not captured."""
    tp = os.path.join(ROOT, "profiles", "top_kernel_traffic.json")
    if not os.path.exists(tp):
        return None
    t = json.load(open(tp))
    return (t.get(workload) or {}).get(kernel)


def run_b200(args):
    import torch
    import torch.distributed as dist

    import pcd_b200 as P

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus or world == 1, f"launched with WORLD_SIZE={world} but --gpus {args.gpus}"
    lib = P._lib.load()
    assert lib.pcd_check_device() == 0, lib.pcd_last_error().decode()
    pk = peaks()

    wl = Workload(P, args.workload, dev, world, rank, args.batch)
    barrier, timed = make_timers(torch, dist, world, dev)
    for _ in range(max(args.warmup, 1)):
        wl.step_device()
    clocks = ClockSampler(local)
    clocks.start()
    t_dev = timed(wl.step_device, args.steps)
    clk = clocks.stop()
    launches_per_step = wl.launches_per_step()
    wl.step_e2e()
    t_e2e = timed(wl.step_e2e, args.steps)

    value = wl.global_batch * args.steps / t_dev
    e2e_value = wl.global_batch * args.steps / t_e2e
    metric = METRIC if wl.n_out == 1024 else METRIC.replace("1024 pts", f"{wl.n_out} pts")
    h2d, d2h = wl.bytes_per_step()
    ws_bytes = sum(t.numel() for t in wl.model._ws.values())
    cfgline = workload_config(args.workload, world)
    if args.batch:
        cfgline.update(per_gpu_batch=wl.B, global_batch=wl.global_batch, note="--batch override")
    detail = dict(points=wl.n_out, seq_len=wl.seq_len, sequences_per_eval=(2 if wl.guided else 1) * wl.B,
                  cache=f"activations of one evaluation ({ws_bytes / 2 ** 30:.2f} GiB workspace) exceed the 126 MB L2; "
                        f"no explicit flush between timed iterations")
    step_ms = 1e3 * t_dev / args.steps
    line = {"metric": metric, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True,
            "scaling": wl.w["scaling"], "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": cfgline, "workload_detail": detail, "denoiser_ms_per_heun_step": step_ms / 64,
            "clocks": clk,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches_per_step * args.steps}
    if rank == 0:
        try:
            seqs_eval = (2 if wl.guided else 1) * wl.B
            kb = kernel_breakdown(P, seqs_eval, wl.seq_len, wl.cfg["width"], wl.cfg["heads"], wl.cfg["layers"], 127,
                                  wl.B, wl.C, wl.N, pk)
            dom = max((k for k in kb if kb[k]["launches_per_step"]), key=lambda k: kb[k]["ms_per_step"])
            kms = sum(v["ms_per_step"] for v in kb.values())
            # the same kernel seen from inside the step: its share of the (back-to-back timed) kernel total, applied
            # to the device-timed step and divided by its launches
            for v in kb.values():
                if v["launches_per_step"]:
                    v["share_of_step"] = v["ms_per_step"] / kms
                    v["in_step_ms"] = step_ms * v["share_of_step"] / v["launches_per_step"]
            if dom == "flash_attention":
                # hd = 64 attention is bound by the 16 ex2/clk/SM special-function rate (tools/ubench), not by the
                # tensor pipe: the achieved fraction of THAT ceiling at the clock the step ran at
                ex2_per_s = 16.0 * 148 * (clk.get("sm_mhz") or 1500.0) * 1e6
                kb[dom]["mufu_ceiling_tflops"] = 4.0 * 64 * ex2_per_s / 1e12   # FLOP per logit = 4*hd
                kb[dom]["frac_of_mufu_ceiling"] = kb[dom]["achieved"] / kb[dom]["mufu_ceiling_tflops"]
            r = kb[dom]
            in_step_achieved = r["achieved"] * r["ms"] / r["in_step_ms"]
            line["roofline"] = {"kernel": dom, "bound": r["bound"], "achieved": r["achieved"], "peak": r["peak"],
                                "unit": r["unit"], "frac": r["frac"], "traffic": traffic_for(args.workload, dom),
                                "in_step": {"ms": r["in_step_ms"], "achieved": in_step_achieved,
                                            "frac": in_step_achieved / r["peak"], "share_of_step": r["share_of_step"]},
                                "peak_source": pk["source"] + (", sustained: every kernel is timed back to back for >= 1 s, "
                                                               "i.e. at the power-capped clocks of the step itself"
                                                               if r["bound"] == "tensor" else ", copy bandwidth")}
            line["kernels"] = {k: {kk: (round(vv, 4) if isinstance(vv, float) else vv) for kk, vv in v.items()}
                               for k, v in kb.items()}
            line["kernel_ms_sum_per_step"] = kms
        except Exception as ex:  # the headline number must still be printed
            line["roofline"] = {"error": repr(ex)}
    wl.release()

    others = []
    if args.workload == HEADLINE and args.others != "none" and not args.batch:
        names = OTHER_WORKLOADS if args.others == "all" else OTHER_WORKLOADS[:2]
        for name in names:
            try:
                others.append(measure_other(P, name, dev, world, rank, torch, dist, pk))
            except Exception as ex:
                others.append({"workload": name, "error": repr(ex)[:300]})
                torch.cuda.empty_cache()
        if rank == 0:
            try:
                others.append(perceiver_bench(P, dev, 32, pk))
            except Exception as ex:
                others.append({"workload": "perceiver-text-1026x77", "error": repr(ex)[:300]})
    if rank == 0:
        if others:
            line["other_workloads"] = others
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"], _ = cpu_sample(args.workload, heun_steps=2)
            for row in others:
                if row.get("workload") in WORKLOADS and "error" not in row:
                    try:
                        row["cpu_baseline"], _ = cpu_sample(row["workload"], heun_steps=1)
                    except Exception as ex:
                        row["cpu_baseline"] = {"error": repr(ex)[:200]}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "eager-gpu"])
    ap.add_argument("--workload", default=HEADLINE, choices=list(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="override the per-GPU batch (debug)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs")
    ap.add_argument("--others", default="all", choices=["all", "fast", "none"],
                    help="other BASELINE workloads measured after the headline (fast: without the 300M model)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.impl == "eager-gpu":
        run_eager_gpu(args)
    elif args.workload.startswith("twostream"):
        run_twostream(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
