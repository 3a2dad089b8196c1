#!/bin/bash
# ncu --set full capture of one tensor-core attention launch per variant at the bench shape: tools/gpu_profile_attn.sh "9 8 10"
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
L=${ATTN_L:-1026}
for V in ${1:-8}; do
  CMD="python tools/attn_probe.py $L $V"
  ATTN_REPS=1 ATTN_IT=2 $CMD > gpurun_out/attn_plain_$V.log 2>&1 || { echo "plain run failed"; tail gpurun_out/attn_plain_$V.log; exit 1; }
  cat gpurun_out/attn_plain_$V.log
  ATTN_REPS=1 ATTN_IT=2 ncu --set full --clock-control none --import-source on -k regex:"attn_bf16_tc" -s 3 -c 1 -f -o gpurun_out/prof_attn_v$V $CMD > gpurun_out/ncu_attn_$V.log 2>&1
  echo "attn capture v$V rc=$?"
done
