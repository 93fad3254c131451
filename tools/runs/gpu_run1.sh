#!/bin/bash
# first contact: smoke, parity tests, short bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke" ; timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/smoke.log
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -60 gpurun_out/pytest.log
echo "== bench config1"; timeout 600 python bench.py --workload config1 --steps 10 --warmup 3 > gpurun_out/bench_c1.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/bench_c1.log
echo "== bench config2"; timeout 900 python bench.py --steps 10 --warmup 3 --cpu-sessions 32 > gpurun_out/bench_c2.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/bench_c2.log
