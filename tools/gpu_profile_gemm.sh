#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 120 python tools/gemm_one.py > gpurun_out/gemm_plain.log 2>&1 || { echo "plain failed"; tail gpurun_out/gemm_plain.log; exit 1; }
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"gemm_bf16_tc" -s 2 -c 1 -o gpurun_out/prof_gemm python tools/gemm_one.py > gpurun_out/ncu_gemm.log 2>&1
echo "gemm capture rc=$?"
