#!/bin/bash
for v in 1 0; do echo "T4R_GEMM_2CTA=$v"; T4R_GEMM_2CTA=$v timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider --durations=8 2>&1 | tail -14; done
