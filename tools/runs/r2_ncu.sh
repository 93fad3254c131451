#!/bin/bash
# Round 2, after r2_first.sh is green: ncu evidence for whichever head configuration won the A/B.
#   gpurun --timeout 1500 -- 'HEADCFG="T4R_HEAD_RESIDENT=1" NPROD=2 bash tools/runs/r2_ncu.sh'
# Produces (copy the summaries into profiles/ as r2_*):
#   gpurun_out/r2_launches.csv      per-launch durations of 3 bench steps  (shares of the step, not absolutes)
#   gpurun_out/r2_head.ncu-rep      --set full capture of the head kernel (tensor pipe %, L2->SM bytes, DRAM bytes)
mkdir -p gpurun_out
NPROD=${NPROD:-3}
export $HEADCFG
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --nprod $NPROD > gpurun_out/r2_launches_bench.log 2>&1
python tools/summarize_ncu.py launches gpurun_out/r2_launches.csv gpurun_out/r2_launches.txt || true
ncu --set full --clock-control none --import-source on -k regex:"head_resident_kernel|gemm2_bf16x3_kernel" -s 6 -c 1 \
    -o gpurun_out/r2_head -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --nprod $NPROD > gpurun_out/r2_head_ncu.log 2>&1
python tools/summarize_ncu.py full gpurun_out/r2_head.ncu-rep gpurun_out/r2_head_summary.txt || true
tail -5 gpurun_out/r2_launches.txt; cat gpurun_out/r2_head_summary.txt | head -20
