#!/bin/bash
mkdir -p gpurun_out
echo "== bench config1"; timeout 600 python bench.py --workload config1 --steps 20 --warmup 5 > gpurun_out/bench_c1.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/bench_c1.log
echo "== bench config2"; timeout 900 python bench.py --steps 20 --warmup 5 --cpu-sessions 32 > gpurun_out/bench_c2.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/bench_c2.log
echo "== bench config2 nprod1"; timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --nprod 1 > gpurun_out/bench_c2_n1.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/bench_c2_n1.log
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 120 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1; echo "rc=$?"
echo "== ncu full head"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16x3_kernel -s 30 -c 6 -o gpurun_out/prof_gemm_r1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "rc=$?"
ls -la gpurun_out
