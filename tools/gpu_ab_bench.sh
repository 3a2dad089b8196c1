#!/bin/bash
# A/B of the product library against a side build (tools/lib/$1) on the headline and TwoStream benches, same box
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
SIDE=$PWD/tools/lib/$1
one() { # label, env...
  local label=$1; shift
  env "$@" timeout -s KILL 600 python bench.py --steps 2 --warmup 3 --others none --no-cpu > gpurun_out/ab_$label.log 2>&1
  grep -h '^{' gpurun_out/ab_$label.log | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$label headline', round(d['value'],3), 'clouds/s', d['ms_per_step'])"
  env "$@" timeout -s KILL 600 python bench.py --workload twostream-config-1024pt-b32 --steps 2 --warmup 2 --no-cpu > gpurun_out/ab_ts_$label.log 2>&1
  grep -h '^{' gpurun_out/ab_ts_$label.log | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$label twostream', round(d['value'],3), 'clouds/s', d['ms_per_step'])"
}
timeout 900 python -m pytest -q -p no:cacheprovider -m gpu tests/test_gpu_sampler.py tests/test_gpu_forward.py -x 2>&1 | tail -2
one product A=1
one side PCD_B200_LIB=$SIDE
one product2 A=1
one side2 PCD_B200_LIB=$SIDE
