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
cfg = P.MODEL_CONFIGS[args.config]
model = P.model_from_config(cfg, dev, dtype=torch.bfloat16)
with torch.no_grad():
    model.output_proj.weight.normal_(std=0.02)
B = args.batch
emb = torch.randn(B, 768, device=dev)
emb = emb / emb.norm(dim=1, keepdim=True)
kw = dict(embeddings=torch.cat([emb, torch.zeros_like(emb)], 0))
x = torch.randn(B, cfg["input_channels"], cfg["n_ctx"], device=dev)
d = P.diffusion_from_config(P.DIFFUSION_CONFIGS[args.config])
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
