#!/bin/bash
for dbg in 2 66; do echo "T4R_GEMM_DEBUG=$dbg (64 = staged residual read)"; T4R_GEMM_DEBUG=$dbg timeout 200 python tools/microbench.py ffn oproj 2>&1 | tail -7; done
timeout 600 python -m pytest tests/test_gpu_parity.py -q --timeout 300 -p no:cacheprovider -x -k "epilogues or ffn or xlnet or gpt2 or variants" 2>&1 | tail -2
