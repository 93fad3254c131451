#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_fullsize.py -q --timeout 600 -p no:cacheprovider 2>&1 | tail -40
