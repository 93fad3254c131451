#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm2_bf16x3_kernel -c 1 -s 2 -o gpurun_out/r1c_head_pair -f python tools/microbench.py head > gpurun_out/ncu_head.log 2>&1; tail -2 gpurun_out/ncu_head.log
timeout 600 ncu --set full --clock-control none -k regex:ffn_fused_kernel -c 1 -s 3 -o gpurun_out/r1c_ffn -f python tools/microbench.py ffn > gpurun_out/ncu_ffn2.log 2>&1; tail -1 gpurun_out/ncu_ffn2.log
