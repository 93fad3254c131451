#!/bin/bash
# Round 2, final check on two GPUs: the multi-process sharded tests (world 1 and 2, peer + nccl), the DDP training test,
# and the bench line at N = 2 as the driver launches it.
mkdir -p gpurun_out
{
echo "== multi-process tests"; timeout 1200 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_peer.py tests/test_gpu_zz_training.py -q -p no:cacheprovider 2>&1 | tail -6
echo "== bench N=2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29577 bench.py --gpus 2 --steps 20 --warmup 5 2> gpurun_out/final_n2.err | tail -1 | tee gpurun_out/r2_final_bench_n2.json | cut -c1-300; tail -2 gpurun_out/final_n2.err
} > gpurun_out/r2_final_2gpu.log 2>&1
cat gpurun_out/r2_final_2gpu.log
