"""Per-kernel timings at the bench shape (min of 3 x 10 launches): quick A/B probe (tools/ab.sh)."""
import math, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pcd_b200 as P
ops = P.ops
dev = torch.device("cuda")
B2, L, W, H = 128, 1026, 512, 8
if len(sys.argv) > 1:
    B2, L, W, H = [int(v) for v in sys.argv[1].split(",")]
M = B2 * L
bf = torch.bfloat16
g = torch.Generator(device=dev).manual_seed(1)
a_d = torch.randn(M, W, device=dev, generator=g).to(bf)
a_4d = torch.randn(M, 4 * W, device=dev, generator=g).to(bf)
h = torch.randn(M, W, device=dev, generator=g)
mk = lambda n, k: (torch.randn(n, k, device=dev, generator=g) / math.sqrt(k)).to(bf)
w_qkv, w_proj, w_fc, w_fc2 = mk(3 * W, W), mk(W, W), mk(4 * W, W), mk(W, 4 * W)
bias = lambda n: torch.zeros(n, device=dev)
hb, stats = ops.cast_rowstats(h)
cs = lambda w: w.float().sum(dim=1).contiguous()
cs_qkv, cs_fc = cs(w_qkv), cs(w_fc)
b3, b1, b4 = bias(3 * W), bias(W), bias(4 * W)
qkv = torch.randn(B2, L, 3 * W, device=dev, generator=g).to(bf)
hid = torch.empty(M, 4 * W, device=dev, dtype=bf)

def t(fn, it=10):
    best = 1e9
    for _ in range(3):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(it): fn()
        b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / it)
    return best * 1e3
xin = torch.randn(B2 // 2, 6, L - 2, device=dev, generator=g)
w_in, b_in = torch.randn(W, 6, device=dev, generator=g), bias(W)
pre = torch.randn(B2, 2, W, device=dev, generator=g)
ones = torch.ones(W, device=dev)
res = {
    "embed": t(lambda: ops.embed_tokens(xin, w_in, b_in, pre, None, ones, b_in, seqs=B2)),
    "embed_stats": t(lambda: ops.embed_tokens(xin, w_in, b_in, pre, None, ones, b_in, seqs=B2, with_stats=True)),
    "cast_rowstats": t(lambda: ops.cast_rowstats(h)),
    "output_proj": t(lambda: ops.output_proj(h.view(B2, L, W), 2, ones, b_in, w_in.t()[:6].contiguous(), b_in[:6])),
    "qkv_lnfold": t(lambda: ops.linear_layernorm_folded(hb, stats, w_qkv, cs_qkv, b3)),
    "proj_resid_stats": t(lambda: ops.linear_residual_stats(a_d, w_proj, b1, h)),
    "fc1_lnfold_gelu": t(lambda: ops.linear_layernorm_folded(hb, stats, w_fc, cs_fc, b4, gelu=True)),
    "fc2_resid_stats": t(lambda: ops.linear_residual_stats(a_4d, w_fc2, b1, h)),
    "fc1_plain_gelu": t(lambda: ops.linear(a_d, w_fc, b4, epilogue=1, out=hid)),
    "attention": t(lambda: ops.self_attention(qkv, H)),
}
print("  ".join(f"{k} {v:.1f}" for k, v in res.items()), " sum %.1f us" % sum(v for k, v in res.items() if k in ("qkv_lnfold", "proj_resid_stats", "fc1_lnfold_gelu", "fc2_resid_stats", "attention")))
