#!/bin/bash
# One GPU-box round: parity tests by risk group (isolated processes + timeouts so one hung
# kernel cannot take the others down), smoke(), a short bench.  Logs -> gpurun_out/.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
OUT=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt 2>&1
PYT="python -m pytest -q -p no:cacheprovider --timeout 240 -m gpu"
run() { # name, timeout, cmd...
  local name=$1 to=$2; shift 2
  echo "=== $name ($(date +%T))"
  timeout -s KILL $to "$@" > $OUT/$name.log 2>&1
  echo "    exit=$? ; tail:"; tail -n ${TAILN:-6} $OUT/$name.log | sed 's/^/    /'
}
run ops_simt   600 $PYT tests/test_gpu_ops.py -k "device or timestep or layernorm or gemm_f32 or chamfer or rotary or rowstats"
run probe 300 python tools/gemm_probe.py
run gemm_tc    600 $PYT tests/test_gpu_ops.py -k "gemm_bf16 or residual_stats or layernorm_folded"
run attn       600 $PYT tests/test_gpu_ops.py -k "attention or perceiver or rotary"
run pointcloud 300 $PYT tests/test_gpu_point_cloud.py
run forward    900 $PYT tests/test_gpu_forward.py
run sampler    900 $PYT tests/test_gpu_sampler.py -s
run twostream  400 $PYT tests/test_gpu_twostream.py
run ddpm       300 $PYT tests/test_gpu_ddpm.py
run smoke      600 python __graft_entry__.py --smoke
TAILN=3 run bench 900 python bench.py --steps 2 --warmup 3
grep -h '^{' $OUT/bench.log | tail -1 > $OUT/bench.json
echo "=== done ($(date +%T))"
