#!/bin/bash
# ncu --set full capture of one tensor-core attention launch at the bench shape (variant $1, default: library default)
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
V=${1:-}
CMD="python tools/attn_probe.py 1026 ${V}"
$CMD > gpurun_out/attn_plain.log 2>&1 || { echo "plain run failed"; tail gpurun_out/attn_plain.log; exit 1; }
cat gpurun_out/attn_plain.log
ncu --set full --clock-control none --import-source on -k regex:"attn_bf16_tc" -s 3 -c 1 -f -o gpurun_out/prof_attn${V} $CMD > gpurun_out/ncu_attn.log 2>&1
echo "attn capture rc=$?"
