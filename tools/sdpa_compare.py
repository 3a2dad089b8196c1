"""Library attention (torch SDPA backends, flash_attn if importable) at the bench's attention shape, next to this
library's tcgen05 kernel -- a same-box yardstick.  q, k, v: [B=128, H=8, L=1026, 64] bf16, scale 1/8."""
import json, os, sys
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pcd_b200 as P
dev = torch.device("cuda")
B, L, H = 128, 1026, 8
torch.manual_seed(0)
qkv = (torch.randn(B, L, H * 192, device=dev) * 1.5).bfloat16()
x = qkv.view(B, L, H, 3, 64)
q, k, v = (x[:, :, :, i].permute(0, 2, 1, 3).contiguous() for i in range(3))   # [B, H, L, 64]
fl = 4.0 * L * L * 64 * H * B
def t(fn, it=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / it * 1e-3
res = {}
def rec(name, fn):
    try:
        s = t(fn)
        res[name] = dict(us=s * 1e6, tflops=fl / s / 1e12)
        print(f"{name:28s} {s*1e6:8.1f} us  {fl/s/1e12:7.1f} TFLOP/s")
    except Exception as ex:
        res[name] = dict(error=repr(ex)[:200])
        print(f"{name:28s} unavailable: {repr(ex)[:120]}")
rec("pcd_b200 (tcgen05, variant 5)", lambda: P.ops.self_attention(qkv, H))
from torch.nn.attention import SDPBackend, sdpa_kernel
for name, be in (("torch SDPA flash", SDPBackend.FLASH_ATTENTION), ("torch SDPA cudnn", SDPBackend.CUDNN_ATTENTION),
                 ("torch SDPA mem-efficient", SDPBackend.EFFICIENT_ATTENTION)):
    def run(be=be):
        with sdpa_kernel(be):
            return F.scaled_dot_product_attention(q, k, v)
    rec(name, run)
try:
    from flash_attn import flash_attn_func
    qf, kf, vf = (x[:, :, :, i].contiguous() for i in range(3))   # [B, L, H, 64]
    rec("flash_attn 2 (package)", lambda: flash_attn_func(qf, kf, vf))
except Exception as ex:
    print("flash_attn unavailable:", repr(ex)[:100])
json.dump(res, open("gpurun_out/sdpa_compare.json", "w"), indent=1)
