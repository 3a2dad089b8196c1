#!/bin/bash
# Round-2 profile set (one gpurun call, ncu only after the same command exited 0 without it):
#   1. launch list of the bench command itself (shares vs bench.py's CUDA-event shares)
#   2. ncu --set full of one attention launch and one launch of each projection inside a real denoiser forward
#   3. DRAM traffic of the attention kernel at the upsampler / 300M shapes (per-workload roofline.traffic)
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
OUT=gpurun_out
MINE='regex:^(gemm_|attn_|layernorm|embed_tokens|output_proj|sampler_|timestep|rope_|cast_)'
BENCH="python bench.py --steps 1 --warmup 3 --no-cpu --others none"
$BENCH > $OUT/prof_bench_plain.log 2>&1 || { echo "plain bench failed"; tail -5 $OUT/prof_bench_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -k "$MINE" -s 200 -c 320 --csv --log-file $OUT/launches_bench.csv $BENCH > $OUT/ncu_launches_bench.log 2>&1
echo "bench launch list rc=$?"
python tools/summarize_launches.py $OUT/launches_bench.csv | head -20

FWD="python tools/profile_forward.py --iters 2"
$FWD > $OUT/prof_fwd_plain.log 2>&1 || { echo "plain forward failed"; tail -5 $OUT/prof_fwd_plain.log; exit 1; }
# one evaluation = 1 cast + 12 x (qkv, attention, proj, fc1, fc2) + small kernels: skip the first evaluation
ncu --set full --clock-control none --import-source on -k regex:attn_bf16 -s 13 -c 1 -f -o $OUT/prof_attn_grouped $FWD > $OUT/ncu_attn.log 2>&1
echo "attention capture rc=$?"
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tc2 -s 52 -c 4 -f -o $OUT/prof_gemm_block $FWD > $OUT/ncu_gemm.log 2>&1
echo "gemm capture rc=$?"
# DRAM traffic of every hot kernel of the headline forward (one block) and of the attention kernel at the other shapes
TR="dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum"
ncu --metrics $TR --clock-control none -k "$MINE" -s 70 -c 8 --csv --log-file $OUT/traffic_base40M-imagevec-1024pt-b64.csv $FWD > /dev/null 2>&1
for cfg in "upsample:64:upsample-4096pt-b128" "base300M-upsample:16:base300M-upsample-4096pt-b64"; do
  IFS=: read c b name <<< "$cfg"
  python tools/profile_forward.py --iters 1 --config $c --batch $b > $OUT/prof_fwd_$c.log 2>&1 || { echo "plain $c failed"; continue; }
  ncu --metrics $TR --clock-control none -k "$MINE" -s 10 -c 8 --csv --log-file $OUT/traffic_$name.csv python tools/profile_forward.py --iters 1 --config $c --batch $b > /dev/null 2>&1
  echo "traffic $name rc=$? (batch $b x2 sequences)"
done
python tools/collect_traffic.py $OUT
