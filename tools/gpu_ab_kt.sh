#!/bin/bash
# A/B of tools/kernel_times.py: product library vs a side build (tools/lib/$1), alternating in one box
cd "${GRAFT_REPO_ROOT:-/root/repo}"
SIDE=tools/lib/$1
for rep in 1 2; do
  echo "--- product"; timeout -s KILL 300 python tools/kernel_times.py 2>&1 | tail -1
  echo "--- side ($SIDE)"; PCD_B200_LIB=$PWD/$SIDE timeout -s KILL 300 python tools/kernel_times.py 2>&1 | tail -1
done
timeout 600 python -m pytest -q -p no:cacheprovider -m gpu tests/test_gpu_ops.py -k "gemm or linear or gelu" 2>&1 | tail -3
timeout 600 python -m pytest -q -p no:cacheprovider -m gpu tests/test_gpu_forward.py 2>&1 | tail -3
