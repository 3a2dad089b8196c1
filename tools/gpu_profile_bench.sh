#!/bin/bash
# ncu launch list of the bench command itself (per-launch durations are cold-cache / serialised:
# compare SHARES with bench.py's live CUDA-event numbers, not absolutes).
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu"
MINE='regex:^(gemm_|attn_|layernorm|embed_tokens|output_proj|sampler_|timestep|rope_|cast_)'
# skip the eager enqueue used for graph capture bookkeeping: list 3 consecutive denoiser evaluations
ncu --metrics gpu__time_duration.sum --clock-control none -k "$MINE" -s 200 -c 300 --csv --log-file gpurun_out/launches_bench.csv $CMD > gpurun_out/ncu_launches_bench.log 2>&1
echo "bench launch list rc=$?"
python tools/summarize_launches.py gpurun_out/launches_bench.csv | head -20
