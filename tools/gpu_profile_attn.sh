#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
CMD="python tools/profile_forward.py --iters 1"
$CMD > gpurun_out/profile_plain.log 2>&1 || { echo "plain run failed"; tail gpurun_out/profile_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:"attn_bf16_tc" -s 3 -c 1 -o gpurun_out/prof_attn $CMD > gpurun_out/ncu_attn.log 2>&1
echo "attn capture rc=$?"
