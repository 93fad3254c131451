#!/bin/bash
# Round 2, thirteenth call (1 GPU, short): TMA-store epilogue of the planes-only dense GEMM (Q|K|V projection).
mkdir -p gpurun_out
{
echo "== bit-exactness"; timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "tma_store" -p no:cacheprovider 2>&1 | tail -12
for F in 0 1 0 1; do echo "== microbench qkvp, T4R_GEMM_TMA_STORE=$F"; T4R_GEMM_TMA_STORE=$F timeout 300 python tools/microbench.py qkvp 2>&1 | tail -1; done
echo "== encoder / model tests with the TMA stores"; T4R_GEMM_TMA_STORE=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zz_plm.py tests/test_gpu_zz_attn64.py -q -p no:cacheprovider 2>&1 | tail -4
for F in 0 1; do echo "== bench, T4R_GEMM_TMA_STORE=$F"; T4R_GEMM_TMA_STORE=$F timeout 600 python bench.py --no-cpu-baseline --no-sharded --steps 20 --warmup 5 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), d['stages_ms'])"; done
} > gpurun_out/r2_thirteenth.log 2>&1
cat gpurun_out/r2_thirteenth.log
