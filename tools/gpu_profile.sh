#!/bin/bash
# ncu evidence: per-launch durations of the step's kernels (+ optional full captures).
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
CMD="python tools/profile_forward.py --iters 2"
MINE='regex:^(gemm_|attn_|layernorm|embed_tokens|output_proj|sampler_|timestep)'
$CMD > gpurun_out/profile_plain.log 2>&1 || { echo "plain run failed"; tail gpurun_out/profile_plain.log; exit 1; }
# one forward = 89 launches of this library; skip the first (cold) forward, list the next two
# evaluations + the sampler kernels between them
ncu --metrics gpu__time_duration.sum --clock-control none -k "$MINE" -s 90 -c 182 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
if [ "${FULL:-1}" = "1" ]; then
ncu --set full --clock-control none --import-source on -k regex:"gemm_bf16_tc|attn_bf16_tc" -s 5 -c 5 -o gpurun_out/prof_hot $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"sampler_|layernorm" -s 4 -c 3 -o gpurun_out/prof_hbm $CMD > gpurun_out/ncu_full2.log 2>&1
echo "hbm capture rc=$?"
fi
ls -la gpurun_out/
