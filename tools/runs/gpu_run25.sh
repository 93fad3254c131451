#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parity.py -q --timeout 600 -p no:cacheprovider -x -k "head or end_to_end or sampled or smoothing" 2>&1 | tail -3
timeout 300 python tools/microbench.py head head64 2>&1 | tail -4
