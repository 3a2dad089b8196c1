"""Timeline of the softmax warps of one CTA of the grouped attention kernel (tools build with -DPCD_ATTN_TRACE):

    python tools/attn_trace.py build      # here (no GPU): tools/lib/libpcd_trace.so
    PCD_B200_LIB=tools/lib/libpcd_trace.so python tools/attn_trace.py [L] [variant]     # on the GPU box

Prints, per slot, the mean cycles between the stamps of a step (steady-state steps only) and a few raw steps."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
LIB = os.path.join(ROOT, "tools", "lib", "libpcd_trace.so")

if len(sys.argv) > 1 and sys.argv[1] == "build":
    import importlib
    b = importlib.import_module("a-multimodal-diffusion-based-model-for-point-cloud-completion_b200.build")
    print(b.build_variant(LIB, ["-DPCD_ATTN_TRACE"], "trace"))
    sys.exit(0)

import numpy as np
import torch

import pcd_b200 as P

L = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
variant = int(sys.argv[2]) if len(sys.argv) > 2 else 8
lib = P._lib.load()
lib.pcd_attn_trace_read.restype = C.c_int
lib.pcd_attn_trace_read.argtypes = [C.c_void_p, C.c_int]
B, H = 128, 8
qkv = (torch.randn(B, L, H * 192, device="cuda") * 1.5).bfloat16()
for _ in range(3):
    P.ops.self_attention(qkv, H, variant=variant)
torch.cuda.synchronize()
NT, STEPS, PTS = 3, 96, 12
buf = np.zeros(NT * STEPS * PTS, dtype=np.uint64)
assert lib.pcd_attn_trace_read(buf.ctypes.data, buf.size) == buf.size
t = buf.reshape(NT, STEPS, PTS).astype(np.int64)
num_kv = (L + 63) // 64
names = ["wait S_FULL", "LDTM", "row max (+S_FREE)", "wait PV_DONE", "(token)", "exponentials", "st wait + P_READY", "loop -> next step"]
for s in range(NT):
    steps = [j for j in range(2, min(STEPS, 4 * num_kv) - 1) if j % num_kv not in (0, num_kv - 1) and t[s, j, 0] > 0]
    d = np.array([[t[s, j, p + 1] - t[s, j, p] for p in range(7)] + [t[s, j + 1, 0] - t[s, j, 7]] for j in steps])
    whole = np.array([t[s, j + 1, 0] - t[s, j, 0] for j in steps])
    print(f"slot {s}: {len(steps)} steady steps, mean step {whole.mean():7.1f} cycles (min {whole.min()} max {whole.max()})")
    for n, m, mx in zip(names, d.mean(0), d.max(0)):
        print(f"    {n:22s} {m:7.1f}  (max {mx})")
print("per-step duration (cycles) of slot 0 / 1 / 2, item boundaries every", num_kv, "steps:")
for j in range(0, min(STEPS - 1, 3 * num_kv + 3)):
    mark = " <- first tile of an item" if j % num_kv == 0 else (" <- last tile" if j % num_kv == num_kv - 1 else "")
    print(f"  step {j:3d}: " + "  ".join(f"{int(t[s, j + 1, 0] - t[s, j, 0]):6d} (exp {int(t[s, j, 6] - t[s, j, 5]):5d}, S wait {int(t[s, j, 1] - t[s, j, 0]):5d}, PV wait {int(t[s, j, 4] - t[s, j, 3]):5d})" for s in range(NT)) + mark)
print("item boundaries (slot 0/1/2): loop end -> epilogue begin | wait last PV | O ld | stores issued | -> next item's first step")
for j in range(num_kv - 1, STEPS - 1, num_kv):
    print(f"  after step {j:3d}: " + "   ".join(" | ".join(str(int(x)) for x in (t[s, j, 8] - t[s, j, 7], t[s, j, 9] - t[s, j, 8], t[s, j, 10] - t[s, j, 9], t[s, j, 11] - t[s, j, 10], t[s, j + 1, 0] - t[s, j, 11])) for s in range(NT)))
for j in range(20, 22):
    print("step", j, " ".join(f"s{s}:" + ",".join(str(int(t[s, j, p] - t[0, 20, 0])) for p in range(8)) for s in range(NT)))
