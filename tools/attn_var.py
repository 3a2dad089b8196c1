"""Run-to-run variation of the attention kernel (variants 5 and 3) with the SM clock sampled next to each
timing, cold and after a GEMM burn-in that drives the chip into its power cap."""
import os, subprocess, sys, math
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pcd_b200 as P
lib = P._lib.load(); ops = P.ops
dev = torch.device("cuda")
B, L, H = 128, 1026, 8
qkv = (torch.randn(B, L, H * 192, device=dev) * 1.5).bfloat16()
a = torch.randn(B * L, 2048, device=dev).bfloat16(); w = (torch.randn(512, 2048, device=dev) / 45).bfloat16()
bias = torch.zeros(512, device=dev); y = torch.empty(B * L, 512, device=dev, dtype=torch.bfloat16)
def clock():
    o = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-i", "0"], capture_output=True, text=True).stdout.strip()
    return o
def t(fn, it=20):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it * 1e3
for phase in ("cold", "after 2 s of GEMMs"):
    if phase != "cold":
        for _ in range(8000): ops.linear(a, w, bias, out=y)
        torch.cuda.synchronize()
    for v in (8, 5, 8, 5, 8, 5):
        us = t(lambda: ops.self_attention(qkv, H, variant=v))
        print(f"{phase}: variant {v}: {us:7.1f} us   [sm MHz, W] {clock()}")
# long steady run of variant 5: 400 launches
for i in range(5):
    us = t(lambda: ops.self_attention(qkv, H, variant=5), it=100)
    print(f"steady v5 x100: {us:7.1f} us   {clock()}")
