#!/bin/bash
echo "== parity with the CTA-pair FFN"
T4R_FFN_2CTA=1 timeout 600 python -m pytest tests/test_gpu_parity.py -q --timeout 300 -p no:cacheprovider -x -k "ffn or xlnet or gpt2 or end_to_end or hidden or fixture" 2>&1 | tail -6
echo "== microbench"
for v in 0 1; do T4R_FFN_2CTA=$v timeout 200 python tools/microbench.py ffn attn 2>&1 | grep -v "^ " | sed "s/^/2cta=$v /"; done
for v in 0 1; do echo "T4R_FFN_2CTA=$v"; T4R_FFN_2CTA=$v timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['value'], d['ms_per_step'], 'head', d['roofline']['launch_ms'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'])"; done
