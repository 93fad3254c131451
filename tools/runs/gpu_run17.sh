#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q --timeout 600 -p no:cacheprovider -k "embed or end_to_end or fixture" 2>&1 | tail -4
timeout 300 python tools/microbench.py embed 2>&1 | tee gpurun_out/mb_embed.log
