#!/bin/bash
mkdir -p gpurun_out
echo "== microbench"; timeout 600 python tools/microbench.py 2>&1 | tee gpurun_out/microbench.log
echo "== ncu dense"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_bf16x3_kernel" -s 8 -c 6 -o gpurun_out/prof_dense_r1 python tools/microbench.py qkv oproj > gpurun_out/ncu_dense.log 2>&1; echo "rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"attn_kernel" -s 2 -c 1 -o gpurun_out/prof_attn_r1 python tools/microbench.py attn > gpurun_out/ncu_attn.log 2>&1; echo "rc=$?"
ls -la gpurun_out
