#!/bin/bash
# GEMM work loop: parity of the tensor-core GEMM tests, the cuBLAS yardstick, and the epilogue / issuer timeline
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
PYT="python -m pytest -q -p no:cacheprovider --timeout 300 -m gpu"
timeout -s KILL 600 $PYT tests/test_gpu_ops.py -x -k "gemm_bf16 or residual_stats or layernorm_folded" > gpurun_out/gemm_tests.log 2>&1; echo "tests exit=$?"; tail -2 gpurun_out/gemm_tests.log
timeout -s KILL 300 python tools/cublas_compare.py 2>&1 | tee gpurun_out/cublas_compare.log | tail -8
for w in ${TRACE:-fc1 qkv}; do
  PCD_B200_LIB=$PWD/tools/lib/libpcd_gtrace.so timeout -s KILL 120 python tools/gemm_trace.py $w > gpurun_out/gemm_trace_$w.log 2>&1
  head -8 gpurun_out/gemm_trace_$w.log | cut -c1-220; grep -A5 "MMA issuer" gpurun_out/gemm_trace_$w.log | cut -c1-120
done
