#!/bin/bash
# Round 2, eleventh call (1 GPU, short): the resident head's epilogue with the column scales fetched one chunk ahead:
# head tests, kernel timing (was 3.76 ms for nprod = 2 on the resident-A kernel), bench step.
mkdir -p gpurun_out
{
echo "== head tests"; timeout 900 python -m pytest tests/test_gpu_zz_mixed_head.py tests/test_gpu_parity.py tests/test_gpu_peer.py -q -k "head or model or sharded" -p no:cacheprovider 2>&1 | tail -4
echo "== head kernel timing"; timeout 600 python tools/microbench.py head headres head2 2>&1 | tail -8
echo "== bench"; for i in 1 2; do timeout 600 python bench.py --no-cpu-baseline --no-sharded --steps 20 --warmup 5 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), round(d['roofline']['launch_ms'],3), round(d['roofline']['frac'],3), d['stages_ms'])"; done
} > gpurun_out/r2_eleventh.log 2>&1
cat gpurun_out/r2_eleventh.log
