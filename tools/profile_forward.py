"""Short workload for ncu: a few denoiser forwards + sampler updates at the bench shapes
(base40M-imagevec, 128 sequences x L=1026, bf16)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pcd_b200 as P

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--config", default="base40M-imagevec")
args = ap.parse_args()
dev = torch.device("cuda")
torch.manual_seed(0)
if args.config == "base300M-upsample":  # BASELINE config 5: grid-upsample class with base300M dims
    cfg = dict(P.MODEL_CONFIGS["upsample"], width=1024, layers=24, heads=16)
else:
    cfg = P.MODEL_CONFIGS[args.config]
model = P.model_from_config(cfg, dev, dtype=torch.bfloat16)
if hasattr(model, "accept_grid_embeddings"):
    model.accept_grid_embeddings = True
with torch.no_grad():
    model.output_proj.weight.normal_(std=0.02)
B = args.batch
cls = cfg["name"]
kw = {}
if cls == "CLIPImagePointDiffusionTransformer":
    emb = torch.randn(B, 768, device=dev)
    kw["embeddings"] = emb / emb.norm(dim=1, keepdim=True)
if "Grid" in cls:
    kw["embeddings"] = torch.randn(B, 1024, 256, device=dev)
if "Upsample" in cls:
    lr = torch.rand(B, cfg["input_channels"], cfg["cond_ctx"], device=dev) - 0.5
    lr[:, 3:] = (lr[:, 3:] + 0.5) * 255.0
    kw["low_res"] = lr
kw = {k: torch.cat([v, torch.zeros_like(v)], 0) for k, v in kw.items()}
x = torch.randn(B, cfg["input_channels"], cfg["n_ctx"], device=dev)
d = P.diffusion_from_config(P.DIFFUSION_CONFIGS["upsample" if "upsample" in args.config else args.config])
plan = P.HeunPlan(d, 64, 1e-3, 120.0, 7.0, 3.0)
st = P.k_diffusion.HeunState(d, plan, tuple(x.shape), dev, 3.0, True)
st.x.copy_(x * 120)
nz = torch.randn_like(x)
pred = torch.empty_like(x)
st.begin(nz)
for i in range(args.iters):
    out = model.forward_cfg(st.model_in, 1017, kw, doubled=True, out_channels=6)
    st.predictor(0, out, pred)
    out = model.forward_cfg(st.model_in, 1017, kw, doubled=True, out_channels=6)
    st.corrector(0, out, nz)
torch.cuda.synchronize()
print("ok", float(pred.abs().mean()))
