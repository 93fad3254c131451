#!/bin/bash
for dbg in 4 8 12 13; do echo "T4R_GEMM_DEBUG=$dbg"; T4R_GEMM_DEBUG=$dbg timeout 200 python tools/microbench.py ffn 2>&1 | grep "fused ffn"; done
