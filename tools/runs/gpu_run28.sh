#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,power.limit,clocks.max.sm,temperature.gpu --format=csv | tail -1
for v in 0 1 0 1; do echo "T4R_GEMM_2CTA=$v"; T4R_GEMM_2CTA=$v timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['value'], d['ms_per_step'], 'head', d['roofline']['launch_ms'], 'frac', d['roofline']['frac'], d['clocks'], 'e2e', d['e2e']['value'])"; done
echo "== all gpu tests with the CTA-pair GEMM"
T4R_GEMM_2CTA=1 timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider 2>&1 | tail -3
