#!/bin/bash
# Round 2, ninth call (1 GPU, short): the fused FFN with sixteen epilogue warps (T4R_FFN_EPW=16, two-stage ring):
# parity of everything that goes through it, then the A/B of the kernel and of the bench step.
mkdir -p gpurun_out
{
echo "== tests with T4R_FFN_EPW=16"; T4R_FFN_EPW=16 timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "encoder or ffn or model or config or variants" -p no:cacheprovider 2>&1 | tail -4
for E in 8 16 8 16; do echo "== microbench ffn, EPW=$E"; T4R_FFN_EPW=$E timeout 300 python tools/microbench.py ffn attn 2>&1 | tail -2; done
for E in 8 16; do echo "== bench config 2, EPW=$E"; T4R_FFN_EPW=$E timeout 600 python bench.py --no-cpu-baseline --no-sharded --steps 20 --warmup 5 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), d['stages_ms'])"; done
} > gpurun_out/r2_ninth.log 2>&1
cat gpurun_out/r2_ninth.log
