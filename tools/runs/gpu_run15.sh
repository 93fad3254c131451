#!/bin/bash
# round-1 evidence run: parity, bench (both arms), ncu launch list + full captures
mkdir -p gpurun_out
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest.log
echo "== smoke"; timeout 600 python __graft_entry__.py smoke 2>&1 | tail -2
echo "== bench config2 (with cpu baseline)"; timeout 1200 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_c2.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_c2.log | cut -c1-250
echo "== bench reference arm"; timeout 1200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_ref.log | cut -c1-250
echo "== bench config1"; timeout 600 python bench.py --workload config1 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c1.log 2>&1; tail -1 gpurun_out/bench_c1.log | cut -c1-200
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 160 -c 120 --csv --log-file gpurun_out/launches_r1c.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1; echo "rc=$?"
echo "== ncu full (head, LN gemm, dense gemm, attention, gather)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_bf16x3_kernel|attn_mma_kernel|embed_concat_kernel" -s 20 -c 22 -o gpurun_out/prof_r1c python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "rc=$?"
ls -la gpurun_out | tail -12
