#!/bin/bash
# Round 2, twelfth call (1 GPU): coalesced label compaction, count-limited gather / mixed packing -- suite + bench.
mkdir -p gpurun_out
{
echo "== whole GPU suite"; timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -6
echo "== launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2c_launches.csv python bench.py --no-cpu-baseline --no-sharded --steps 3 --warmup 3 > /dev/null 2>&1
python tools/summarize_ncu.py launches gpurun_out/r2c_launches.csv gpurun_out/r2c_launches_config2.txt | head -24
echo "== bench"; for i in 1 2; do timeout 600 python bench.py --no-cpu-baseline --no-sharded --steps 20 --warmup 5 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), round(d['e2e']['pipelined_value'] or 0), round(d['roofline']['launch_ms'],3), d['stages_ms'])"; done
} > gpurun_out/r2_twelfth.log 2>&1
cat gpurun_out/r2_twelfth.log
