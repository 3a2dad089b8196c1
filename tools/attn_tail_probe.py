"""Where do the 2 extra tokens of L = 1026 cost time?  Times the attention kernel with independent query / key
lengths (cross-attention entry of the same kernel)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pcd_b200 as P
ops = P.ops
dev = torch.device("cuda")
B, H = 128, 8
def run(Lq, Lkv):
    q = torch.randn(B, Lq, H * 64, device=dev).bfloat16()
    kv = torch.randn(B, Lkv, H * 128, device=dev).bfloat16()
    fn = lambda: ops.cross_attention(q, kv, H)
    for _ in range(30): fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20): fn()
        b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / 20 * 1e3)
    print(f"Lq={Lq:5d} Lkv={Lkv:5d}: {best:7.1f} us")
import ast
for Lq, Lkv in (ast.literal_eval(sys.argv[1]) if len(sys.argv) > 1 else [(1024, 1024), (1024, 1026), (1026, 1024), (1026, 1026)]):
    run(Lq, Lkv)
