#!/bin/bash
mkdir -p gpurun_out
T4R_GEMM_DEBUG=2 timeout 300 python tools/microbench.py qkv ffn1 proj 2>&1 | tee gpurun_out/mb_prof.log
