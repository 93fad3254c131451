#!/bin/bash
mkdir -p gpurun_out
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider -x > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest.log
echo "== microbench"; timeout 600 python tools/microbench.py 2>&1 | tee gpurun_out/microbench.log
echo "== bench config2"; timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c2.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_c2.log | cut -c1-300
