#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_bf16x3_kernel" -s 10 -c 1 -o gpurun_out/prof_proj_r1 python tools/microbench.py proj > gpurun_out/ncu_proj.log 2>&1; echo "rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_bf16x3_kernel" -s 10 -c 1 -o gpurun_out/prof_oproj_r1 python tools/microbench.py oproj > gpurun_out/ncu_oproj.log 2>&1; echo "rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"attn_kernel" -s 2 -c 1 -o gpurun_out/prof_attn_r1b python tools/microbench.py attn > gpurun_out/ncu_attn.log 2>&1; echo "rc=$?"
ls -la gpurun_out | tail -5
