#!/bin/bash
# Round 2, eighth call (1 GPU, short): (1) attention with ex2.approx / pairwise bias split: encoder parity tests + layer
# time; (2) A/B of the fused FFN with a 2-stage instead of a 3-stage operand ring (exp_build/libt4r_b200_ffn2.so).
mkdir -p gpurun_out
{
echo "== encoder / PLM / attention tests"; timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zz_attn64.py tests/test_gpu_zz_plm.py -q -k "encoder or attn or plm or model" -p no:cacheprovider 2>&1 | tail -4
echo "== microbench, shipped library"; timeout 300 python tools/microbench.py ffn oproj qkvp attn 2>&1 | tail -6
echo "== microbench, 2-stage FFN ring"; T4R_LIB_PATH=$PWD/exp_build/libt4r_b200_ffn2.so timeout 300 python tools/microbench.py ffn oproj attn 2>&1 | tail -5
echo "== again shipped"; timeout 300 python tools/microbench.py ffn attn 2>&1 | tail -3
} > gpurun_out/r2_eighth.log 2>&1
cat gpurun_out/r2_eighth.log
