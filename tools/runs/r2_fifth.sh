#!/bin/bash
# Round 2, fifth call (1 GPU): dropout kernels + the sampled head's hit-column path through the suite, then config-5
# lines (single-GPU shape and the sharded leg at N = 1) and the default bench line.
mkdir -p gpurun_out
{
echo "== whole GPU suite"; timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -8
echo "== bench config 5 (one rank's shard, replicated)"; timeout 600 python bench.py --workload config5 --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/r2b_bench_config5.json | cut -c1-1800
echo "== bench config 2, defaults"; timeout 900 python bench.py --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/r2b_bench_config2.json | cut -c1-6000
echo "== training step timing"; timeout 600 python bench.py --train --optimizer adamw --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-400
} > gpurun_out/r2_fifth.log 2>&1
tail -60 gpurun_out/r2_fifth.log
