#!/bin/bash
for dbg in 0 16 32; do echo "T4R_GEMM_DEBUG=$dbg"; T4R_GEMM_DEBUG=$dbg timeout 200 python tools/microbench.py head64 head 2>&1 | grep "^head"; done
