#!/bin/bash
# wide geometry of the grouped attention kernel (two slots, 128-key steps): parity, then timing against the default
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest -q -p no:cacheprovider --timeout 120 -m gpu tests/test_gpu_ops.py -x -k "attention" 2>&1 | tail -8
timeout -s KILL 300 python -m pytest -q -p no:cacheprovider --timeout 120 -m gpu tests/test_gpu_guard.py -k "attention" 2>&1 | tail -3
VARIANTS=8,11 timeout -s KILL 300 python tools/attn_ragged_probe.py ${PAIRS:-1026x1026,1024x1024,1026x1026,1024x1024,4353x4353,1281x1281}
