#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_inputs.py -q --timeout 600 -p no:cacheprovider 2>&1 | tail -12
