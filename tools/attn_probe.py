"""Time the tensor-core attention variants at the bench shape and check them against torch."""
import math, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pcd_b200 as P
lib = P._lib.load()
dev = torch.device("cuda")
B, L, H = 128, 1026, 8
if len(sys.argv) > 1:
    L = int(sys.argv[1])
torch.manual_seed(0)
qkv = (torch.randn(B, L, H * 192, device=dev) * 1.5).bfloat16()
def ref(nb=2):
    x = qkv[:nb].float().view(nb, L, H, 192)
    q, k, v = x[..., :64], x[..., 64:128], x[..., 128:]
    w = torch.einsum("bthc,bshc->bhts", q, k) / 8.0
    return torch.einsum("bhts,bshc->bthc", torch.softmax(w, -1), v).reshape(nb, L, -1)
want = ref()
def t(fn, it=int(os.environ.get('ATTN_IT', '10'))):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / it
fl = 4.0 * L * L * 64 * H * B
variants = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [8, 9, 5]
for rep in range(int(os.environ.get("ATTN_REPS", "2"))):
    for v in variants:
        out = P.ops.self_attention(qkv, H, variant=v)
        torch.cuda.synchronize()
        err = float((out[:2].float() - want).norm() / want.norm())
        ms = t(lambda: P.ops.self_attention(qkv, H, variant=v))
        print(f"L={L} variant {v}: {ms*1e3:8.1f} us  {fl/ms/1e9:7.1f} TFLOP/s  rel err {err:.2e}", flush=True)
