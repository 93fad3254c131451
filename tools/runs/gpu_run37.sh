#!/bin/bash
mkdir -p gpurun_out
echo "== pytest (all gpu tests, 1 GPU)"; timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_r1c.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_r1c.log
echo "== bench config2"; timeout 600 python bench.py --steps 30 --warmup 5 --graph > gpurun_out/r1c_bench_config2.json 2> gpurun_out/bench_err.log; python -c "
import json
d=json.loads(open('gpurun_out/r1c_bench_config2.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['launch_ms'], d['roofline']['frac'], d['clocks'], d.get('cuda_graph'), d['cpu_baseline']['value'], d['gpu_launches'])"
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1c_bench_reference_arm.json 2>> gpurun_out/bench_err.log; cut -c1-160 gpurun_out/r1c_bench_reference_arm.json
echo "== bench config1"; timeout 300 python bench.py --workload config1 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r1c_bench_config1.json 2>> gpurun_out/bench_err.log; cut -c1-140 gpurun_out/r1c_bench_config1.json
echo "== launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1c_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/b_under_ncu.log 2>&1; wc -l gpurun_out/r1c_launches.csv
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
