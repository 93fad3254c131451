#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sharded.py -q --timeout 600 -p no:cacheprovider -x 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 2>/dev/null | tail -1 > gpurun_out/r1c_bench_config2_n2.json; python -c "
import json
d=json.loads(open('gpurun_out/r1c_bench_config2_n2.json').read())
print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['launch_ms'])"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --workload config4 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r1c_bench_config4_n2.json; python -c "
import json
d=json.loads(open('gpurun_out/r1c_bench_config4_n2.json').read())
print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['launch_ms'], d['roofline']['frac'])"
