#!/bin/bash
# Round 2, third call (2 GPUs): the peer-memory kernels (single process), the sharded multi-process tests in both
# formulations at world 1 and 2, the widened-input training test, then bench.py with its `sharded` record at N = 1, 2.
mkdir -p gpurun_out
{
nvidia-smi -L; nvidia-smi topo -m 2>/dev/null | head -8
echo "== peer kernels, one process"; timeout 300 python -m pytest tests/test_gpu_peer.py -q -x -p no:cacheprovider 2>&1 | tail -15
echo "== sharded path, world 1 and 2, peer + nccl"; timeout 900 python -m pytest tests/test_gpu_sharded.py -q -p no:cacheprovider 2>&1 | tail -25
echo "== widened training"; timeout 300 python -m pytest tests/test_gpu_zz_training.py -q -k widened -p no:cacheprovider 2>&1 | tail -5
echo "== bench N=1 (with sharded record)"; timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_n1.err | tail -1 > gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err; cut -c1-6000 gpurun_out/bench_n1.json
echo "== bench N=2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 2> gpurun_out/bench_n2.err | tail -1 > gpurun_out/bench_n2.json; tail -3 gpurun_out/bench_n2.err; cut -c1-6000 gpurun_out/bench_n2.json
} > gpurun_out/r2_third.log 2>&1
tail -80 gpurun_out/r2_third.log
