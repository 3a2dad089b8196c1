"""HBM roofline of the fused ancestral-sampling step (pcd_ddpm_step) at a batch far larger than L2."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pcd_b200 as P
dev = torch.device("cuda")
B, C, N = 4096, 6, 1024
d = P.diffusion_from_config(P.DIFFUSION_CONFIGS["base40M-imagevec"])     # learned_range, channel scaling
x = torch.randn(B, C, N, device=dev)
out = torch.randn(B, 2 * C, N, device=dev)
noise = torch.randn(B, C, N, device=dev)
t = torch.full((B,), 500, device=dev, dtype=torch.int64)
model = lambda x_, t_, **kw: out
def step():
    return d._ddpm_step(model, x, t, True, None, noise, ("x_next", "pred_xstart", "sample_unscaled"))
for _ in range(3): step()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20): step()
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 20
arrays = 1 + 2 + 1 + 3           # read x, eps + variance channels, noise; write x', pred, unscaled sample
gb = arrays * B * C * N * 4 / 1e9
pk = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else 6550.0
print(json.dumps({"kernel": "ddpm_step_kernel<4>", "B": B, "C": C, "N": N, "ms": ms, "algorithmic_GB": gb,
                  "achieved_GBps": gb / ms * 1e3, "peak_GBps": pk, "frac": gb / ms * 1e3 / pk,
                  "note": "through the host wrapper (three torch.empty_like per step)"}))
