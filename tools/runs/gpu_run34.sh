#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py -q --timeout 600 -p no:cacheprovider -x 2>&1 | tail -3
T4R_GEMM_DEBUG=2 timeout 200 python tools/microbench.py ffn 2>&1 | tail -4
timeout 200 python tools/microbench.py ffn oproj ffn2 attn 2>&1 | grep -v "^ "
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['value'], d['ms_per_step'], 'head', d['roofline']['launch_ms'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'])"
