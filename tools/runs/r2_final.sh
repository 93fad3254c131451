#!/bin/bash
# Round 2, final refresh on one GPU: the whole suite, smoke(), the default bench line (CPU baseline, Recall legs,
# sharded record) and the reference arm.
mkdir -p gpurun_out
{
echo "== whole GPU suite"; timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -8
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
echo "== bench, defaults"; timeout 900 python bench.py 2>&1 | tail -1 | tee gpurun_out/r2_final_bench_config2.json | cut -c1-400
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | tee gpurun_out/r2_final_bench_reference.json | cut -c1-300
echo "== config 1 (graph)"; timeout 600 python bench.py --workload config1 --steps 50 --warmup 10 --graph --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/r2_final_bench_config1.json | cut -c1-300
} > gpurun_out/r2_final.log 2>&1
cat gpurun_out/r2_final.log | cut -c1-500
