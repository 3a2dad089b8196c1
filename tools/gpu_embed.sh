#!/bin/bash
# embed_tokens with fused bf16 copy + statistics: parity, kernel times, headline line
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 600 python -m pytest -q -p no:cacheprovider -m gpu tests/test_gpu_ops.py -k "embed_tokens or cast_rowstats or output_proj" 2>&1 | tail -5
timeout 600 python -m pytest -q -p no:cacheprovider -m gpu tests/test_gpu_forward.py tests/test_gpu_sampler.py 2>&1 | tail -5
timeout 300 python tools/kernel_times.py 2>&1 | tail -2
timeout 900 python bench.py --steps 3 --warmup 3 --others none > gpurun_out/bench_embed.log 2>&1; grep -h '^{' gpurun_out/bench_embed.log | tail -1 > gpurun_out/bench_embed.json
python tools/benchsum.py gpurun_out/bench_embed.json 2>/dev/null | head -30
