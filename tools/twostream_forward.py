"""One guided evaluation (2B sequences) of the config.yaml TwoStreamDenoiser, for ncu launch lists:
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/launches_twostream.csv python tools/twostream_forward.py
(only the last, steady-state evaluation is inside the cudaProfilerStart/Stop range)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pcd_b200 as P  # noqa: E402
from bench import TWOSTREAM_CONFIG as c  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = P.TwoStreamDenoiser(**c, device=dev, dtype=torch.bfloat16)
kw = dict(class_labels=torch.randint(1, 10, (B,), device=dev), viewpoints=torch.rand(B, 3, device=dev),
          partial_pcd=torch.rand(B, 1024, 3, device=dev) - 0.5, depth_maps=torch.rand(B, 1, 512, 512, device=dev))
kw = {k: torch.cat([v, torch.zeros_like(v)]) for k, v in kw.items()}
x = torch.randn(B, 3, 1024, device=dev)
model.begin_trajectory(2 * B)
for i in range(3):  # evaluation 0 also runs the condition encoders; 1 and 2 are steady state
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = P.ops.launch_count()
    if i == 2:
        torch.cuda.profiler.start()
    e0.record()
    model.forward_cfg(x, 500 - i, kw, True)
    e1.record()
    torch.cuda.synchronize()
    if i == 2:
        torch.cuda.profiler.stop()
    print(f"evaluation {i}: {e0.elapsed_time(e1):.2f} ms, {P.ops.launch_count() - n0} launches")
