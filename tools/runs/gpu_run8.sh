#!/bin/bash
mkdir -p gpurun_out
echo "== normal"; timeout 300 python tools/microbench.py qkv ffn1 oproj ffn2 proj 2>&1 | tee gpurun_out/mb_normal.log
echo "== epilogue without global traffic"; T4R_GEMM_DEBUG=1 timeout 300 python tools/microbench.py qkv ffn1 oproj ffn2 proj 2>&1 | tee gpurun_out/mb_noepi.log
