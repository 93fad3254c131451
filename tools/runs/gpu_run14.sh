#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 40 --csv --log-file gpurun_out/launches_layer.csv python tools/microbench.py attn > gpurun_out/ncu_l.log 2>&1; echo "rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"attn_mma_kernel" -s 2 -c 1 -o gpurun_out/prof_attn_mma python tools/microbench.py attn > gpurun_out/ncu_attn.log 2>&1; echo "rc=$?"
