#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sharded.py -q --timeout 600 -p no:cacheprovider -x 2>&1 | tail -15
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload config4 --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -3 | cut -c1-1500 | tee gpurun_out/bench_config4_n2.json
