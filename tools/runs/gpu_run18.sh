#!/bin/bash
mkdir -p gpurun_out
echo "== encoder parity with fused FFN"
T4R_FFN_FUSED=1 timeout 600 python -m pytest tests/test_gpu_parity.py -q --timeout 300 -p no:cacheprovider -x -k "xlnet or gpt2 or encoder or end_to_end or fixture" 2>&1 | tail -8
echo "== layer timing unfused / fused"
T4R_FFN_FUSED=0 timeout 200 python tools/microbench.py attn 2>&1 | tail -2
T4R_FFN_FUSED=1 timeout 200 python tools/microbench.py attn 2>&1 | tail -2
echo "== bench unfused / fused"
T4R_FFN_FUSED=0 timeout 300 python bench.py --steps 10 --warmup 3 2>&1 | tail -1 | cut -c1-400
T4R_FFN_FUSED=1 timeout 300 python bench.py --steps 10 --warmup 3 2>&1 | tail -1 | cut -c1-400
