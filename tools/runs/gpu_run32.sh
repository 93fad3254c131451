#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py -q --timeout 600 -p no:cacheprovider -x 2>&1 | tail -4
