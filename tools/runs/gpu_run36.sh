#!/bin/bash
for v in 0 1 0 1; do T4R_GEMM_2CTA=$v timeout 300 python tools/microbench.py head4 2>&1 | grep "^head" | sed "s/^/2cta=$v /"; done
