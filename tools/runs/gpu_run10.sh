#!/bin/bash
# 2-GPU validation: sharded lookup/head over NCCL + weak-scaling bench at N=2
mkdir -p gpurun_out
nvidia-smi -L
echo "== sharded test"; timeout 900 python -m pytest tests/test_gpu_sharded.py -q --timeout 600 -p no:cacheprovider 2>&1 | tail -5
echo "== bench N=1"; timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n1.log 2>&1; tail -1 gpurun_out/bench_n1.log | cut -c1-200
echo "== bench N=2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2.log 2>&1; tail -1 gpurun_out/bench_n2.log | cut -c1-200
