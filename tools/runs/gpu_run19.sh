#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/microbench.py ffn 2>&1 | tail -2
T4R_GEMM_DEBUG=1 timeout 200 python tools/microbench.py ffn 2>&1 | tail -1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ffn_fused -c 1 -s 3 -o gpurun_out/r1b_ffn_fused -f python tools/microbench.py ffn > gpurun_out/ncu_ffn.log 2>&1; tail -2 gpurun_out/ncu_ffn.log
