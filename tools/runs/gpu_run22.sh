#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_inputs.py -q --timeout 300 -p no:cacheprovider 2>&1 | tail -40
timeout 900 python -m pytest tests/test_gpu_parity.py -q --timeout 600 -p no:cacheprovider -x 2>&1 | tail -3
