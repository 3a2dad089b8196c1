"""Where the L = 1026 overhead of the grouped attention kernel goes: time (len_q, len_kv) in {1024, 1026}^2 on one buffer."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pcd_b200 as P
ops = P.ops
dev = torch.device("cuda")
B, L, H = 128, 1026, 8
torch.manual_seed(0)
qkv = (torch.randn(B, L, H * 192, device=dev) * 1.5).bfloat16()
out = torch.empty(B, L, H * 64, device=dev, dtype=torch.bfloat16)
def run(lq, lkv, variant=0):
    q = ops._operand(qkv, 0, L * H * 192, H * 192, 192)
    k = ops._operand(qkv, 64, L * H * 192, H * 192, 192)
    v = ops._operand(qkv, 128, L * H * 192, H * 192, 192)
    s = 64 ** -0.25
    ops.attention_packed(q, k, v, out, B, H, lq, lkv, s, s, None, variant)
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
import itertools
pairs = [(int(a), int(b)) for a, b in (p.split('x') for p in sys.argv[1].split(','))] if len(sys.argv) > 1 else [(1024, 1024), (1026, 1024), (1024, 1026), (1026, 1026)]
variants = [int(v) for v in os.environ.get("VARIANTS", "0").split(",")]
for lq, lkv in pairs:
    print(f"len_q {lq} len_kv {lkv}: " + "  ".join(f"v{v} {t(lambda: run(lq, lkv, v)):7.1f} us" for v in variants), flush=True)
