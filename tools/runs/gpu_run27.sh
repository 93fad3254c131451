#!/bin/bash
echo "== parity with the CTA-pair GEMM"
T4R_GEMM_2CTA=1 timeout 600 python -m pytest tests/test_gpu_parity.py -q --timeout 300 -p no:cacheprovider -x 2>&1 | tail -12
echo "== microbench 1-CTA"
timeout 300 python tools/microbench.py qkvp oproj ffn1 ffn2 proj head64 head 2>&1 | grep -v "^ "
echo "== microbench 2-CTA"
T4R_GEMM_2CTA=1 timeout 300 python tools/microbench.py qkvp oproj ffn1 ffn2 proj head64 head 2>&1 | grep -v "^ "
