#!/bin/bash
for v in 0 1; do echo "== T4R_FFN_2CTA=$v"; T4R_FFN_2CTA=$v timeout 600 python -m pytest tests/test_gpu_parity.py -q --timeout 300 -p no:cacheprovider -x -k "ffn or xlnet or gpt2 or hidden" 2>&1 | tail -2
T4R_FFN_2CTA=$v timeout 200 python tools/microbench.py ffn attn 2>&1 | grep -v "^ "; done
