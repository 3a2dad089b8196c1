#!/bin/bash
# quick A/B of the attention variants (correctness of each is asserted by attn_probe's rel err print + the variant tests)
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
PYT="python -m pytest -q -p no:cacheprovider --timeout 300 -m gpu"
timeout -s KILL 600 $PYT tests/test_gpu_ops.py -x -k "attention_bf16_variants or self_attention or rotary" > gpurun_out/attn_tests.log 2>&1; echo "tests exit=$?"; tail -3 gpurun_out/attn_tests.log
for L in 1026 1024; do
  timeout -s KILL 300 python tools/attn_probe.py $L ${VARIANTS:-8,9,5} 2>&1 | tee gpurun_out/probe$L.log
done
ATTN_IT=5 ATTN_REPS=1 timeout -s KILL 300 python tools/attn_probe.py 4353 ${VARIANTS:-8,9,5} 2>&1 | tee gpurun_out/probe4353.log
