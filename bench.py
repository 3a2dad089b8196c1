#!/usr/bin/env python
"""Headline benchmark: completed clouds/s of the full 64-step Karras/Heun sampler.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference|eager-gpu]

A "step" is one complete sampling pass over one batch: 64 Heun steps = 127 denoiser
evaluations (each a 2B-sequence classifier-free-guidance forward) + the fused sampler
updates, replayed from one CUDA graph.  Workload at N=1 is BASELINE.json configs[1]:
base40M-imagevec (width 512, 12 layers, L = 1026), 1024 points, batch 64 per GPU, bf16,
guidance 3, s_churn 3, synthetic unit-norm CLIP embeddings, reference-init weights.
Multi-GPU (torchrun): one rank per GPU, batch sharded (weak scaling), no per-step
collective, one NCCL all-gather of the finished clouds per step.
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
WORKLOADS = {
    # name: (model config, diffusion config, per-GPU batch, sigma_max, s_churn, guidance)
    # BASELINE.json configs[1] -- the configuration the headline metric is quoted on:
    "base40M-imagevec-1024pt-b64": ("base40M-imagevec", "base40M-imagevec", 64, 120.0, 3.0, 3.0),
    # configs[2] (i): text-conditioned = same network fed a CLIP text vector (base40M-textvec)
    "base40M-textvec-1024pt-b32": ("base40M-textvec", "base40M-textvec", 32, 120.0, 3.0, 3.0),
    # configs[3]: upsampler 1024 -> 4096 points (L = 4353), unguided like the reference's text2pc stage 2
    "upsample-4096pt-b16": ("upsample", "upsample", 16, 160.0, 0.0, 0.0),
    # configs[4]: 300M denoiser (width 1024, 24 layers) with image grid + partial-cloud conditioning
    "base300M-upsample-4096pt-b8": ("base300M-upsample", "upsample", 8, 160.0, 0.0, 3.0),
    # the TwoStreamDenoiser of the reference's config.yaml (all four modalities, 57.5 M parameters)
    # under its own sampling settings (config.yaml:40-58: 32 samples, guidance 3, 64 steps, s_churn 0)
    "twostream-config-1024pt-b32": ("twostream", "linear-1000", 32, 120.0, 0.0, 3.0),
}
TWOSTREAM_CONFIG = dict(num_points=1024, num_latents=256, input_channels=3, output_channels=3, latent_dim=256, x_dim=256,
                        num_blocks=6, num_compute_layers=4, num_heads=8, num_classes=10, num_tokens_ppcd=256,
                        num_tokens_depth=128, active_modalities=["class", "view", "partial_pcd", "depth"])


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback (B200_PROFILING.md)")


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
def cpu_sample(workload, heun_steps=4, threads=None):
    """Time `heun_steps` Heun steps of the ORACLE sampler (reference algorithm, fp32 torch
    CPU) for ONE cloud of the workload and scale to clouds/s for the 64-step sampler."""
    import torch

    from oracle import cases, det
    from oracle import denoiser as D
    from oracle import sampler as S
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_oracle_golden import shapes_of

    mcfg, dcfg, _, smax, churn, guidance = WORKLOADS[workload]
    assert mcfg.startswith("base40M-"), "the CPU leg covers the base40M vector-conditioned workloads"
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    cfg = dict(cases.MODEL_CONFIGS[mcfg])
    sd = det.fill_state_dict(shapes_of(cfg), 201, mode="reference", width=cfg["width"])
    tab = S.Tables(**cases.DIFFUSION_CONFIGS["base"])
    e = det.normal((1, 768), 77)
    kw = dict(embeddings=torch.cat([e / e.norm(dim=1, keepdim=True), torch.zeros(1, 768)], 0))
    gen = S.heun_progressive(S.make_model_fn(sd, cfg), tab, (1, 6, cfg["n_ctx"]), steps=64, sigma_min=1e-3,
                             sigma_max=smax, s_churn=churn, guidance_scale=guidance, model_kwargs=kw,
                             noise_fn=cases.DetNoise(5))
    with torch.no_grad():
        next(gen)  # first yield comes after the first (cond+uncond) evaluation: start the clock there
        t0 = time.perf_counter()
        for _ in range(heun_steps):
            next(gen)
        dt = time.perf_counter() - t0
    per_cloud = dt / heun_steps * 64.0
    return dict(value=1.0 / per_cloud, unit=UNIT, cores=threads, kind="port",
                sample=f"oracle (reference algorithm, fp32 torch-CPU) on 1 cloud of {workload}: {heun_steps} of 64 "
                       f"Heun steps ({4 * heun_steps} B=1 denoiser forwards) in {dt:.1f}s, scaled x{64 // heun_steps}"), dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    times = []
    base = None
    for i in range(args.warmup + args.steps):
        base, dt = cpu_sample(args.workload, heun_steps=2)
        if i >= args.warmup:
            times.append(dt)
    per_cloud = (sum(times) / len(times)) / 2 * 64.0
    v = 1.0 / per_cloud
    base["value"] = v
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "note": "each step = bounded CPU sample (2 of 64 Heun steps, 1 cloud)"},
            "cpu_baseline": base,
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ---------------------------------------------------------------------------
# Same-box comparator: the reference algorithm (oracle port, plain PyTorch ops) run eagerly on the
# GPU, fp32 and bf16-autocast.  Reported beside the headline; none of this repo's kernels are on that path.
# ---------------------------------------------------------------------------
def run_eager_gpu(args):
    import torch

    from oracle import cases, det
    from oracle import sampler as S
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_oracle_golden import shapes_of

    if int(os.environ.get("RANK", "0")) != 0:
        return
    mcfg, dcfg, B, smax, churn, guidance = WORKLOADS[args.workload]
    if mcfg == "twostream":
        return run_eager_gpu_twostream(args)
    assert mcfg.startswith("base40M-"), "the eager comparator covers the base40M vector-conditioned workloads"
    B = args.batch or B
    dev = torch.device("cuda:0")
    torch.backends.cuda.matmul.allow_tf32 = False
    cfg = dict(cases.MODEL_CONFIGS[mcfg])
    sd = {k: v.to(dev) for k, v in det.fill_state_dict(shapes_of(cfg), 201, mode="reference", width=cfg["width"]).items()}
    tab = S.Tables(**cases.DIFFUSION_CONFIGS["base"])
    g = torch.Generator(device=dev).manual_seed(1234)
    e = torch.randn(B, 768, device=dev, generator=g)
    kw = dict(embeddings=torch.cat([e / e.norm(dim=1, keepdim=True), torch.zeros_like(e)], 0))
    heun_steps = max(1, args.steps)
    out = {}
    for mode in ("fp32", "bf16-autocast"):
        base_fn = S.make_model_fn(sd, cfg)
        if mode == "fp32":
            fn = base_fn
        else:
            def fn(x, t, _f=base_fn, **k):
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    return _f(x, t, **k).float()
        gen = S.heun_progressive(fn, tab, (B, 6, cfg["n_ctx"]), steps=64, sigma_min=1e-3, sigma_max=smax, s_churn=churn,
                                 guidance_scale=guidance, model_kwargs=kw,
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
                                 "note": f"reference algorithm as plain PyTorch ops on cuda:0 (oracle port; [B,H,L,L] attention "
                                         f"materialised, fp32 softmax); {heun_steps} of 64 Heun steps timed by wall clock "
                                         f"around synchronize, scaled to 64"}}))


def run_eager_gpu_twostream(args):
    """The reference's TwoStreamDenoiser algorithm (oracle restatement, plain PyTorch ops, condition encoders re-run
    in every forward like model.py:498-509) under the oracle's guided Heun loop on cuda:0, fp32 and bf16 autocast."""
    import torch

    import pcd_b200 as P
    from oracle import sampler as S
    from oracle import twostream as OT
    _, _, B, smax, churn, guidance = WORKLOADS[args.workload]
    B = args.batch or B
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
        gen = S.heun_progressive(fn, tab, (B, 3, c["num_points"]), steps=64, sigma_min=1e-3, sigma_max=smax, s_churn=churn,
                                 guidance_scale=guidance, model_kwargs=kw,
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
    _, _, B, smax, churn, guidance = WORKLOADS[args.workload]
    B = args.batch or B
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
                     "note": "useful FLOP only (the zero-padded halves of the 32-wide heads are not counted); ~290 "
                             "launches per evaluation: d = 256 projections at M = 2B x 1024 / 2B x 643 rows, "
                             "tensor-core attention, separate LayerNorm kernels"},
        "clocks": clk}))


# ---------------------------------------------------------------------------
def time_kernel(fn, iters=10, warmup=3):
    import torch
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3  # seconds per launch


def kernel_breakdown(P, model, B2, L, width, heads, layers, Bc, C, N, pk):
    """Live CUDA-event timing of each hot kernel at the workload's shapes (operands larger
    than L2: the activation matrices are 134-538 MB) -> roofline fractions."""
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
    out_h = torch.empty(M, width, device=dev)
    ops = P.ops
    res = {}

    def add(name, fn, launches, flops=None, bytes_=None):
        t = time_kernel(fn)
        r = {"ms": t * 1e3, "launches_per_step": launches, "ms_per_step": t * 1e3 * launches}
        if flops is not None:
            r.update(bound="tensor", achieved=flops / t / 1e12, peak=pk["tf_burst"], unit="TFLOP/s")
        else:
            r.update(bound="hbm", achieved=bytes_ / t / 1e9, peak=pk["hbm"], unit="GB/s")
        r["frac"] = r["achieved"] / r["peak"]
        res[name] = r

    evals = 127
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
        add("cast_rowstats", lambda: ops.cast_rowstats(h), evals, bytes_=M * width * 6.0)
    else:
        add("gemm_qkv", lambda: ops.linear(a_d, w_qkv, bias(3 * width), out=qkv), per, flops=2.0 * M * 3 * width * width)
        add("gemm_attn_proj", lambda: ops.linear(a_d, w_proj, bias(width), out=y), per, flops=2.0 * M * width * width)
        add("gemm_fc1_gelu", lambda: ops.linear(a_d, w_fc, bias(4 * width), epilogue=1, out=hid), per,
            flops=2.0 * M * 4 * width * width)
        add("gemm_fc2", lambda: ops.linear(a_4d, w_fc2, bias(width), out=y), per, flops=2.0 * M * 4 * width * width)
        lnw, lnb = torch.ones(width, device=dev), torch.zeros(width, device=dev)
        y.normal_()
        # fused residual-add + LayerNorm: reads h (fp32) + y (bf16), writes h (fp32) + xn (bf16)
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

    mcfg, dcfg, B, smax, churn, guidance = WORKLOADS[args.workload]
    if args.batch:
        B = args.batch
    if mcfg == "base300M-upsample":  # grid-upsample class with base300M dims
        cfg = dict(P.MODEL_CONFIGS["upsample"], width=1024, layers=24, heads=16)
    else:
        cfg = P.MODEL_CONFIGS[mcfg]
    torch.manual_seed(1234 + rank)
    model = P.model_from_config(cfg, dev, dtype=torch.bfloat16)
    if hasattr(model, "accept_grid_embeddings"):
        model.accept_grid_embeddings = True
    with torch.no_grad():  # reference init + re-randomised output_proj
        model.output_proj.weight.normal_(std=0.02)
    diffusion = P.diffusion_from_config(P.DIFFUSION_CONFIGS[dcfg])
    Cc, N = cfg["input_channels"], cfg["n_ctx"]
    sampler = P.PointCloudSampler(dev, [model], [diffusion], [N], ["R", "G", "B"], guidance_scale=[guidance],
                                  use_karras=[True], karras_steps=[64], sigma_min=[1e-3], sigma_max=[smax],
                                  s_churn=[churn], use_cuda_graph=True)
    # synthetic conditioning: unit-norm CLIP vectors, N(0,1) CLIP grids, U(-0.5,0.5) xyz +
    # U(0,255) rgb partial clouds
    cls = cfg["name"]
    host_kw = {}
    if cls == "CLIPImagePointDiffusionTransformer":
        e = torch.randn(B, 768)
        host_kw["embeddings"] = e / e.norm(dim=1, keepdim=True)
    if "Grid" in cls:
        host_kw["embeddings"] = torch.randn(B, 1024, 256)
    if "Upsample" in cls:
        lr = torch.rand(B, Cc, cfg["cond_ctx"]) - 0.5
        lr[:, 3:] = (lr[:, 3:] + 0.5) * 255.0
        host_kw["low_res"] = lr
    host_kw = {k: v.pin_memory() for k, v in host_kw.items()}
    dev_kw = {k: v.to(dev) for k, v in host_kw.items()}
    n_out = N + (cfg["cond_ctx"] if "Upsample" in cls else 0)
    out_host = torch.empty(B, Cc, n_out).pin_memory()
    gathered = None

    def step_device():
        y = sampler.sample_batch(B, dict(dev_kw))
        if world > 1:
            gather(y)
        return y

    def gather(y):
        nonlocal gathered
        if gathered is None:
            gathered = torch.empty((world * y.shape[0],) + tuple(y.shape[1:]), device=dev)
        dist.all_gather_into_tensor(gathered, y.contiguous())

    def step_e2e():
        kw = {k: v.to(dev, non_blocking=True) for k, v in host_kw.items()}
        y = sampler.sample_batch(B, kw)
        if world > 1:
            gather(y)
        out_host.copy_(y, non_blocking=False)  # device -> host read of the step's result
        return out_host

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

    for _ in range(max(args.warmup, 1)):
        step_device()
    clocks = ClockSampler(local)
    clocks.start()
    n0 = lib.pcd_launch_count()
    t_dev = timed(step_device, args.steps)
    clk = clocks.stop()
    # launches: the graph replays the kernels captured once; count them from one eager enqueue
    stage = next(iter(sampler._graphs.values()))
    c0 = lib.pcd_launch_count()
    stage._enqueue()
    torch.cuda.synchronize()
    launches_per_step = int(lib.pcd_launch_count() - c0)
    step_e2e()
    t_e2e = timed(step_e2e, args.steps)

    value = world * B * args.steps / t_dev
    e2e_value = world * B * args.steps / t_e2e
    metric = METRIC if n_out == 1024 else METRIC.replace("1024 pts", f"{n_out} pts")
    line = {"metric": metric, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t_dev / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": args.workload, "per_gpu_batch": B, "global_batch": world * B,
                       "points": n_out, "heun_steps": 64, "denoiser_evals_per_step": 127,
                       "sequences_per_eval": (2 if guidance not in (0.0, 1.0) else 1) * B,
                       "seq_len": model._prefix_layout()[0] + N, "guidance": guidance, "s_churn": churn,
                       "cache": "activations per forward (1.5 GB) exceed the 126 MB L2; no explicit flush",
                       "parallelism": f"batch-sharded x{world}, all-gather of finished clouds"},
            "denoiser_ms_per_heun_step": 1e3 * t_dev / args.steps / 64,
            "clocks": clk,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": sum(v.numel() * 4 for v in host_kw.values()),
                    "d2h_bytes_per_step": out_host.numel() * 4},
            "gpu_launches": launches_per_step * args.steps}
    if rank == 0:
        pk = peaks()
        clk_mhz = clk.get("sm_mhz")
        try:
            seqs_eval = (2 if guidance not in (0.0, 1.0) else 1) * B
            kb = kernel_breakdown(P, model, seqs_eval, model._prefix_layout()[0] + N, cfg["width"], cfg["heads"],
                                  cfg["layers"], B, Cc, N, pk)
            dom = max((k for k in kb if kb[k]["launches_per_step"]), key=lambda k: kb[k]["ms_per_step"])
            # Which measured peak applies (MEASURED_PEAKS.json holds a burst and a sustained bf16 figure): the kernels are
            # timed alone, but right after the long timed region, so the chip may still sit at its power-capped clocks.
            # If their times add up to the step itself they ran in the step's clock state -> sustained peak; if they add
            # up to clearly less they ran faster than inside the step -> burst peak.
            kms = sum(v["ms_per_step"] for v in kb.values())
            sustained = kms >= 0.97 * (1e3 * t_dev / args.steps)
            for v in kb.values():
                if v["bound"] == "tensor":
                    v["peak"] = pk["tf_sustained"] if sustained else pk["tf_burst"]
                    v["frac"] = v["achieved"] / v["peak"]
            if dom == "flash_attention":
                # hd = 64 attention is bound by the 16 ex2/clk/SM special-function rate (tools/ubench),
                # not by the tensor pipe: report the achieved fraction of THAT ceiling as well
                ex2_per_s = 16.0 * 148 * (clk_mhz or 1700.0) * 1e6
                line_mufu = 4.0 * 64 * ex2_per_s / 1e12   # FLOP per logit = 4*hd
                kb[dom]["mufu_ceiling_tflops"] = line_mufu
                kb[dom]["frac_of_mufu_ceiling"] = kb[dom]["achieved"] / line_mufu
            r = kb[dom]
            traffic = None
            tp = os.path.join(ROOT, "profiles", "top_kernel_traffic.json")
            if os.path.exists(tp):
                traffic = json.load(open(tp)).get(dom)
            line["roofline"] = {"kernel": dom, "bound": r["bound"], "achieved": r["achieved"], "peak": r["peak"],
                                "unit": r["unit"], "frac": r["frac"], "traffic": traffic,
                                "peak_source": pk["source"] + ("" if r["bound"] != "tensor" else
                                                               ", sustained (per-kernel times add up to the step: same clock state)"
                                                               if sustained else ", burst (kernels timed alone ran faster than inside the step)")}
            line["kernels"] = {k: {kk: (round(vv, 4) if isinstance(vv, float) else vv) for kk, vv in v.items()}
                               for k, v in kb.items()}
            line["kernel_ms_sum_per_step"] = sum(v["ms_per_step"] for v in kb.values())
        except Exception as ex:  # the headline number must still be printed
            line["roofline"] = {"error": repr(ex)}
        if world == 1 and not args.no_cpu:
            if args.workload.startswith("base40M"):
                base, _ = cpu_sample(args.workload, heun_steps=4)
                line["cpu_baseline"] = base
            else:
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                                        "sample": "not timed for this workload (CPU leg implemented for the "
                                                  "headline base40M workloads only)"}
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
    ap.add_argument("--workload", default="base40M-imagevec-1024pt-b64", choices=list(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="override the per-GPU batch (debug)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
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
