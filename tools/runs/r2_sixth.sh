#!/bin/bash
# Round 2, sixth call (8 GPUs): the bench line with its sharded record at N = 8 (then N = 4 if time allows), as the driver
# launches it; N = 1 first on the same box so that the note file gives the sharded legs their own efficiency.
mkdir -p gpurun_out
{
nvidia-smi -L | head -8
echo "== N=1"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench8_n1.err | tail -1 > gpurun_out/r2_scale_n1.json; tail -2 gpurun_out/bench8_n1.err; cut -c1-300 gpurun_out/r2_scale_n1.json
for N in 8 4; do
echo "== N=$N"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$N bench.py --gpus $N --steps 10 --warmup 3 2> gpurun_out/bench8_n$N.err | tail -1 > gpurun_out/r2_scale_n$N.json; tail -3 gpurun_out/bench8_n$N.err; cut -c1-300 gpurun_out/r2_scale_n$N.json
done
} > gpurun_out/r2_sixth.log 2>&1
tail -40 gpurun_out/r2_sixth.log
