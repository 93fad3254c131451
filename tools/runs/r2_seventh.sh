#!/bin/bash
# Round 2, seventh call (1 GPU): row parts of the encoder on concurrent streams (bit-exactness test, A/B of the bench
# step for 1 / 2 / 4 parts, config 2 and config 5), and the launch list of one TRAINING step (what dominates its 0.4 s).
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --no-sharded --steps 20 --warmup 5"
{
echo "== encoder tests"; timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "encoder" -p no:cacheprovider 2>&1 | tail -5
for P in 1 2 4 1 2; do
echo "== bench config 2, T4R_ENC_PARTS=$P"; T4R_ENC_PARTS=$P timeout 600 $B 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), d['stages_ms'], round(d['e2e']['value']))"
done
for P in 1 2 4; do
echo "== bench config 5, T4R_ENC_PARTS=$P"; T4R_ENC_PARTS=$P timeout 600 $B --workload config5 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), d['stages_ms'])"
done
echo "== bench config 2 from a CUDA graph, parts 1 / 2"; for P in 1 2; do T4R_ENC_PARTS=$P timeout 600 $B --graph 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d.get('cuda_graph'))"; done
echo "== launch list of a training step"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r2_train_launches.csv python bench.py --train --optimizer adamw --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2_train_bench.log 2>&1
python tools/summarize_ncu.py launches gpurun_out/r2_train_launches.csv gpurun_out/r2_train_launches.txt | head -45
} > gpurun_out/r2_seventh.log 2>&1
tail -80 gpurun_out/r2_seventh.log
