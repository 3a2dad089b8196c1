"""Small end-to-end run for compute-sanitizer: bf16 forward on the LayerNorm-folded path (width 512,
>= 512 tokens), persistent attention with several items per CTA and ragged tails, a short sampler run."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pcd_b200 as P
dev = torch.device("cuda")
torch.manual_seed(0)
cfg = dict(P.MODEL_CONFIGS["base40M-imagevec"], layers=2, n_ctx=300)
model = P.model_from_config(cfg, dev, dtype=torch.bfloat16)
with torch.no_grad():
    model.output_proj.weight.normal_(std=0.02)
B = 3
emb = torch.randn(B, 768, device=dev); emb = emb / emb.norm(dim=1, keepdim=True)
x = torch.randn(B, 6, 300, device=dev)
with torch.no_grad():
    y = model(x, torch.full((B,), 500, device=dev), embeddings=emb)
qkv = torch.randn(5, 333, 8 * 192, device=dev).bfloat16()
o = P.ops.self_attention(qkv, 8)
d = P.diffusion_from_config(P.DIFFUSION_CONFIGS["base40M-imagevec"])
s = P.PointCloudSampler(dev, [model], [d], [300], ["R", "G", "B"], guidance_scale=[3.0], use_karras=[True], karras_steps=[2],
                        sigma_min=[1e-3], sigma_max=[120], s_churn=[3], use_cuda_graph=False)
out = s.sample_batch(B, dict(embeddings=emb))
torch.cuda.synchronize()
print("ok", float(y.abs().mean()), float(o.float().abs().mean()), tuple(out.shape))
