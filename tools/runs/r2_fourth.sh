#!/bin/bash
# Round 2, fourth call (1 GPU): regression of the whole suite, the ncu evidence of this round's defaults (launch list of
# the bench step; --set full captures of the resident-A head, the gather under rotating batches, the fused FFN, the
# O-projection + LayerNorm GEMM, the attention kernel, the sampled head of config 5), and the bench lines of configs 1, 3, 5.
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --no-sharded"
{
echo "== whole GPU suite"; timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -8
echo "== launch list (3 steps of config 2)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv $B --steps 3 --warmup 3 > gpurun_out/r2_launches_bench.log 2>&1
python tools/summarize_ncu.py launches gpurun_out/r2_launches.csv gpurun_out/r2_launches_config2.txt | head -40
cap() {  # name, kernel regex, skip, extra bench args
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c 1 -o gpurun_out/r2_$1 -f $B --steps 3 --warmup 3 $4 > gpurun_out/r2_$1_ncu.log 2>&1
  python tools/summarize_ncu.py full gpurun_out/r2_$1.ncu-rep gpurun_out/r2_$1_full.txt > /dev/null 2>&1 && head -45 gpurun_out/r2_$1_full.txt
  python tools/ncu_lines.py gpurun_out/r2_$1.ncu-rep . 40 > gpurun_out/r2_$1_lines.txt 2>&1
  python tools/ncu_stalls.py gpurun_out/r2_$1.ncu-rep . 0 40 > gpurun_out/r2_$1_stalls.txt 2>&1
  ls -la gpurun_out/r2_$1.ncu-rep; [ "$1" = head ] || rm -f gpurun_out/r2_$1.ncu-rep   # the summaries / raw CSV / per-line stalls travel; one report is kept
}
echo "== ncu full: head"; cap head "head_resident_kernel" 4 ""
echo "== ncu full: gather"; cap gather "embed_concat_kernel" 5 ""
echo "== ncu full: ffn"; cap ffn "ffn_fused_kernel" 9 ""
echo "== ncu full: oproj+LN"; cap oproj "gemm2_bf16x3_kernel<256, *true" 9 ""
echo "== ncu full: attention"; cap attn "attn_mma_kernel" 9 ""
echo "== ncu full: sampled head (config 5 single-GPU shape)"; cap head5 "gemm2_bf16x3_kernel<256, *false, *true|head_resident" 2 "--workload config5"
echo "== bench config 3"; timeout 600 python bench.py --workload config3 --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/r2_bench_config3.json | cut -c1-1500
echo "== bench config 5 (one rank's shard, replicated)"; timeout 600 python bench.py --workload config5 --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/r2_bench_config5.json | cut -c1-1500
echo "== bench config 1"; timeout 600 python bench.py --workload config1 --steps 50 --warmup 10 --graph --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/r2_bench_config1.json | cut -c1-1200
echo "== bench config 2, defaults (CPU baseline + sharded record)"; timeout 900 python bench.py 2>&1 | tail -1 | tee gpurun_out/r2_bench_config2.json | cut -c1-3000
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | tee gpurun_out/r2_bench_reference.json | cut -c1-800
rm -f gpurun_out/*.ncu-rep.tmp
ls -la gpurun_out | head -40
} > gpurun_out/r2_fourth.log 2>&1
tail -150 gpurun_out/r2_fourth.log
