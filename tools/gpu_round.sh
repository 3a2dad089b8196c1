#!/bin/bash
# One GPU-box round: parity tests by risk group (isolated processes + timeouts so one hung
# kernel cannot take the others down), smoke(), the bench as the driver runs it.  Logs -> gpurun_out/.
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
  echo "    exit=$? ; tail:"; tail -n ${TAILN:-6} $OUT/$name.log | sed 's/^/    /'
}
run ops        900 $PYT tests/test_gpu_ops.py
run pointcloud 300 $PYT tests/test_gpu_point_cloud.py
run forward    900 $PYT tests/test_gpu_forward.py
run sampler    900 $PYT tests/test_gpu_sampler.py
run twostream  600 $PYT tests/test_gpu_twostream.py
run ddpm       300 $PYT tests/test_gpu_ddpm.py
run dist       300 $PYT tests/test_gpu_dist.py
run guard      600 $PYT tests/test_gpu_guard.py
run smoke      600 python __graft_entry__.py --smoke
TAILN=3 run bench 1800 python bench.py --steps ${BENCH_STEPS:-3} --warmup 3 ${BENCH_ARGS:-}
grep -h '^{' $OUT/bench.log | tail -1 > $OUT/bench.json
python tools/benchsum.py $OUT/bench.json 2>/dev/null | head -60
echo "=== done ($(date +%T))"
if [ "${EAGER:-0}" = "1" ]; then
  TAILN=3 run eager_gpu 900 python bench.py --impl eager-gpu --steps 2 --warmup 1
  grep -h '^{' $OUT/eager_gpu.log | tail -1 > $OUT/eager_gpu.json
fi
