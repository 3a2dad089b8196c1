"""cuBLAS (torch.mm, bf16 in / bf16 out, no epilogue) at the four projection shapes of the bench, next to this
library's fused GEMMs -- a same-box yardstick for the roofline fractions."""
import math, os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pcd_b200 as P
ops = P.ops
dev = torch.device("cuda")
M, W = 128 * 1026, 512
g = torch.Generator(device=dev).manual_seed(1)
bf = torch.bfloat16
def t(fn, it=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / it * 1e-3
res = {}
h = torch.randn(M, W, device=dev, generator=g)
hb, stats = ops.cast_rowstats(h)
for name, N, K in (("c_qkv", 3 * W, W), ("c_proj", W, W), ("c_fc", 4 * W, W), ("mlp.c_proj", W, 4 * W)):
    a = torch.randn(M, K, device=dev, generator=g).to(bf)
    w = (torch.randn(N, K, device=dev, generator=g) / math.sqrt(K)).to(bf)
    wt = w.t()
    out = torch.empty(M, N, device=dev, dtype=bf)
    fl = 2.0 * M * N * K
    tc = t(lambda: torch.mm(a, wt, out=out))
    bias = torch.zeros(N, device=dev)
    if name == "c_qkv":
        cs = w.float().sum(dim=1).contiguous()
        tm = t(lambda: ops.linear_layernorm_folded(hb, stats, w, cs, bias))
        what = "LayerNorm folded + bias"
    elif name == "c_fc":
        cs = w.float().sum(dim=1).contiguous()
        tm = t(lambda: ops.linear_layernorm_folded(hb, stats, w, cs, bias, gelu=True))
        what = "LayerNorm folded + bias + GELU"
    else:
        tm = t(lambda: ops.linear_residual_stats(a, w, bias, h))
        what = "bias + fp32 residual in place + bf16 copy + row statistics"
    res[name] = dict(N=N, K=K, cublas_us=tc * 1e6, cublas_tflops=fl / tc / 1e12, ours_us=tm * 1e6, ours_tflops=fl / tm / 1e12,
                     ours_epilogue=what)
    print(f"{name:11s} N={N:5d} K={K:5d}: cuBLAS {tc*1e6:7.1f} us {fl/tc/1e12:7.1f} TF/s | ours {tm*1e6:7.1f} us {fl/tm/1e12:7.1f} TF/s ({what})")
json.dump(res, open("gpurun_out/cublas_compare.json", "w"), indent=1)
