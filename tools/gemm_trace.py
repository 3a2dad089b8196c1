"""Timeline of one epilogue warp and of the MMA issuer of one CTA pair of the tcgen05 GEMM (tools build with
-DPCD_GEMM_TRACE):

    python tools/gemm_trace.py build                                    # here: tools/lib/libpcd_gtrace.so
    PCD_B200_LIB=tools/lib/libpcd_gtrace.so python tools/gemm_trace.py [qkv|proj|fc1|fc2]   # on the GPU box
"""
import ctypes as C
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
LIB = os.path.join(ROOT, "tools", "lib", "libpcd_gtrace.so")

if len(sys.argv) > 1 and sys.argv[1] == "build":
    import importlib
    b = importlib.import_module("a-multimodal-diffusion-based-model-for-point-cloud-completion_b200.build")
    print(b.build_variant(LIB, ["-DPCD_GEMM_TRACE"], "gtrace"))
    sys.exit(0)

import numpy as np
import torch

import pcd_b200 as P

which = sys.argv[1] if len(sys.argv) > 1 else "fc1"
lib = P._lib.load()
lib.pcd_gemm_trace_read.restype = C.c_int
lib.pcd_gemm_trace_read.argtypes = [C.c_void_p, C.c_int]
dev = torch.device("cuda")
M, W = 131328, 512
g = torch.Generator(device=dev).manual_seed(1)
h = torch.randn(M, W, device=dev, generator=g)
hb, stats = P.ops.cast_rowstats(h)
mk = lambda n, k: (torch.randn(n, k, device=dev, generator=g) / math.sqrt(k)).to(torch.bfloat16)
if which == "qkv":
    w = mk(3 * W, W); cs = w.float().sum(1).contiguous(); b = torch.zeros(3 * W, device=dev)
    fn = lambda: P.ops.linear_layernorm_folded(hb, stats, w, cs, b)
elif which == "fc1":
    w = mk(4 * W, W); cs = w.float().sum(1).contiguous(); b = torch.zeros(4 * W, device=dev)
    fn = lambda: P.ops.linear_layernorm_folded(hb, stats, w, cs, b, gelu=True)
elif which == "proj":
    a = torch.randn(M, W, device=dev, generator=g).to(torch.bfloat16); w = mk(W, W); b = torch.zeros(W, device=dev)
    fn = lambda: P.ops.linear_residual_stats(a, w, b, h)
else:
    a = torch.randn(M, 4 * W, device=dev, generator=g).to(torch.bfloat16); w = mk(W, 4 * W); b = torch.zeros(W, device=dev)
    fn = lambda: P.ops.linear_residual_stats(a, w, b, h)
for _ in range(3):
    fn()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    fn()
e1.record()
torch.cuda.synchronize()
print(which, f"{e0.elapsed_time(e1) / 10 * 1e3:.1f} us per launch (tracing build)")
T, PTS = 24, 24
buf = np.zeros(2 * T * PTS, dtype=np.uint64)
assert lib.pcd_gemm_trace_read(buf.ctypes.data, buf.size) == buf.size
t = buf.reshape(2, T, PTS).astype(np.int64)
base = t[0, 2, 0]
print("epilogue warp 0 of CTA 0, per tile: [tile start | barrier | acc ready] then per 32-column chunk [staging free, regs loaded, math done, store issued]; cycles since tile start")
for it in range(2, 14):
    e = t[0, it]
    nxt = t[0, it + 1, 0]
    s = f"tile {it:2d} @{e[0] - base:7d}: bar {e[1] - e[0]:5d} acc {e[2] - e[0]:5d} |"
    for c in range(4):
        if e[3 + 4 * c] == 0:
            break
        s += " [" + " ".join(f"{e[3 + 4 * c + q] - e[0]:5d}" for q in range(4)) + "]"
    print(s + f"  -> next tile {nxt - e[0]:6d}")
print("MMA issuer of CTA 0: [tile start -> acc buffer free -> first operands -> last operands]; next tile start")
for it in range(2, 14):
    m = t[1, it]
    print(f"tile {it:2d} @{m[0] - base:7d}: acc-free {m[1] - m[0]:5d}  first-ops {m[2] - m[0]:5d}  last-ops {m[3] - m[0]:5d}  -> next {t[1, it + 1, 0] - m[0]:6d}")
