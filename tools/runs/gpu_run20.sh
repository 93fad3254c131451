#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q --timeout 600 -p no:cacheprovider -x 2>&1 | tail -3
timeout 200 python tools/microbench.py ffn ffn1 qkvp oproj attn 2>&1 | tail -6
timeout 300 python bench.py --steps 10 --warmup 3 2>&1 | tail -1 | cut -c1-330
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ffn_fused -c 1 -s 3 -o gpurun_out/r1b_ffn_fused2 -f python tools/microbench.py ffn > gpurun_out/ncu_ffn.log 2>&1; tail -1 gpurun_out/ncu_ffn.log
