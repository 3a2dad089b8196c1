#!/bin/bash
# ncu evidence for profiles/: full captures of the hot kernels inside a real denoiser forward at the
# bench shape (base40M-imagevec, 128 sequences x L=1026, bf16).  Each capture runs only after the
# same command has exited 0 without ncu.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
CMD="python tools/profile_forward.py --iters 1"
$CMD > gpurun_out/profile_plain.log 2>&1 || { echo "plain run failed"; tail gpurun_out/profile_plain.log; exit 1; }
NCU="ncu --set full --clock-control none --import-source on -f"
# second forward (first is cold): one attention launch, then the four projections of a block
$NCU -k regex:"attn_bf16_tc" -s 13 -c 1 -o gpurun_out/prof_attn $CMD > gpurun_out/ncu_attn.log 2>&1; echo "attn capture rc=$?"
$NCU -k regex:"gemm_bf16_tc2" -s 52 -c 4 -o gpurun_out/prof_gemm $CMD > gpurun_out/ncu_gemm.log 2>&1; echo "gemm capture rc=$?"
ls -la gpurun_out/*.ncu-rep
