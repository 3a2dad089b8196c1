"""Timings of the evaluation-side kernels at the reference's evaluation sizes (evaluation.py: 8192-point
predictions, FPS to 1024, Chamfer + F-score against 8192 / 1024 target points)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pcd_b200 as P
ops = P.ops
dev = torch.device("cuda")
B = 32
torch.manual_seed(0)
pred = torch.rand(B, 8192, 3, device=dev) - 0.5
gt = torch.rand(B, 8192, 3, device=dev) - 0.5
def t(fn, it=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / it
ms = t(lambda: ops.fscore_point_cloud_batch(pred, gt, 0.03))
pairs = 2.0 * B * 8192 * 8192
print(f"fscore  B={B} 8192x8192: {ms:8.3f} ms  {pairs * 8 / ms / 1e9:7.1f} TFLOP/s fp32 (8 flop per pair)  [reference materialises {B * 8192 * 8192 * 3 * 4 / 2**30:.0f} GiB]")
pc, gc = pred.permute(0, 2, 1).contiguous(), gt.permute(0, 2, 1).contiguous()
ms = t(lambda: ops.chamfer_distance_xyz(pc, gc))
print(f"chamfer B={B} 8192x8192: {ms:8.3f} ms  {pairs * 8 / ms / 1e9:7.1f} TFLOP/s fp32")
ms = t(lambda: ops.farthest_point_sample(pred, 1024, 0))
print(f"fps     B={B} 8192->1024: {ms:8.3f} ms  ({ms / 1023 * 1e3:.2f} us per selection round, one CTA per cloud)")
