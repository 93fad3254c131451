#!/bin/bash
# Round 2: compute-sanitizer memcheck over the kernels written this round (peer-memory gather / pull / combine, hit
# columns of the sampled head, dropout + attention-with-dropout items, warp-cooperative attention backward, NG = 4
# epilogues), at their test shapes.
mkdir -p gpurun_out
{
T4R_FFN_EPW=16 timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest -q -p no:cacheprovider \
  tests/test_gpu_peer.py "tests/test_gpu_zz_training.py::test_dropout_kernels_device_vs_host_twin" \
  "tests/test_gpu_zz_training.py::test_training_primitives_device_vs_host_twin" \
  "tests/test_gpu_parity.py::test_head_sampled_softmax" "tests/test_gpu_parity.py::test_fused_ffn" > gpurun_out/r2_memcheck_full.log 2>&1
grep -n "=========" gpurun_out/r2_memcheck_full.log | grep -v "Host Frame" | head -40; tail -4 gpurun_out/r2_memcheck_full.log
echo "exit code: $?"
} > gpurun_out/r2_memcheck.log 2>&1
cat gpurun_out/r2_memcheck.log
