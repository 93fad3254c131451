#!/bin/bash
# Round 2, second call: resident-A head + two-warp attention are now the compiled defaults and the experimental gates
# are gone -- the whole `-m gpu` suite, the failing widened-input training test with its traceback, the device-side
# error table of the three head arithmetics, and a default bench line.
mkdir -p gpurun_out
{
echo "== widened-input training test (traceback)"; timeout 300 python -m pytest tests/test_gpu_zz_training.py -q -x -k widened --tb=long -p no:cacheprovider 2>&1 | tail -60
echo "== whole GPU suite, ungated"; timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -15
echo "== head arithmetics vs fp64 on the device"; timeout 300 python tools/precision_gpu.py 2>&1 | tail -8
echo "== bench, defaults"; timeout 600 python bench.py --steps 20 --warmup 5 2>&1 | tail -1
echo "== bench, nprod 2"; timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --nprod 2 2>&1 | tail -1 | cut -c1-400
} > gpurun_out/r2_second.log 2>&1
tail -40 gpurun_out/r2_second.log
