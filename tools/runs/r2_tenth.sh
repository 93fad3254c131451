#!/bin/bash
# Round 2, tenth call (1 GPU, short): phase counters of the fused FFN's epilogue warps with 8 and 16 epilogue warps;
# timing of the training step with the warp-cooperative attention backward + its tests.
mkdir -p gpurun_out
{
for E in 8 16; do echo "== phase counters, EPW=$E"; T4R_GEMM_DEBUG=2 T4R_FFN_EPW=$E timeout 300 python tools/microbench.py ffn 2>&1 | tail -4; done
echo "== no-global-traffic epilogue (debug bit 1)"; for E in 8 16; do T4R_GEMM_DEBUG=1 T4R_FFN_EPW=$E timeout 300 python tools/microbench.py ffn 2>&1 | tail -1; done
echo "== training tests"; timeout 900 python -m pytest tests/test_gpu_zz_training.py tests/test_gpu_zz_plm.py -q -p no:cacheprovider 2>&1 | tail -4
echo "== training step timing"; timeout 600 python bench.py --train --optimizer adamw --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-330
} > gpurun_out/r2_tenth.log 2>&1
cat gpurun_out/r2_tenth.log
