#!/bin/bash
# Bring-up of the grouped attention kernel (attn_tc8.cu): parity first (isolated processes + timeouts so a hung
# kernel cannot take the rest down), then timing against the paired kernel, then the suites that sit on top of it.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
OUT=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt 2>&1
PYT="python -m pytest -q -p no:cacheprovider --timeout 300 -m gpu"
run() { # name, timeout, cmd...
  local name=$1 to=$2; shift 2
  echo "=== $name ($(date +%T))"
  timeout -s KILL $to "$@" > $OUT/$name.log 2>&1
  echo "    exit=$? ; tail:"; tail -n ${TAILN:-8} $OUT/$name.log | sed 's/^/    /'
}
run attn_small 300 $PYT tests/test_gpu_ops.py -x -k "self_attention_golden or self_attention_lengths or cross_attention"
run attn_var   600 $PYT tests/test_gpu_ops.py -k "attention_bf16_variants"
run rotary     300 $PYT tests/test_gpu_ops.py -k "rotary or perceiver"
TAILN=12 run probe1026 200 python tools/attn_probe.py 1026 8,9,5
TAILN=12 run probe1024 200 python tools/attn_probe.py 1024 8,9,5
TAILN=12 run probe4353 300 env ATTN_IT=5 python tools/attn_probe.py 4353 8,9,5
run forward    900 $PYT tests/test_gpu_forward.py
run sampler    900 $PYT tests/test_gpu_sampler.py
run twostream  600 $PYT tests/test_gpu_twostream.py
run ops_rest   600 $PYT tests/test_gpu_ops.py -k "not attention and not rotary and not perceiver"
TAILN=3 run bench 900 python bench.py --steps 2 --warmup 3 --no-cpu
grep -h '^{' $OUT/bench.log | tail -1 > $OUT/bench.json
echo "=== done ($(date +%T))"
