#!/bin/bash
# Round 2, first gpurun call: validate what was written after round 1's GPU budget ran out.
#   gpurun --timeout 1500 -- 'bash tools/runs/r2_first.sh'
# Order: proven suite first (must stay green), then the element-wise check of the 2-unit product, then the fused
# heads on all three kernels, then timings.  Every step has its own timeout; a trap in an experimental kernel only
# loses that step (separate processes).
mkdir -p gpurun_out
{
echo "== proven GPU suite skipped (driver ran it at the end of round 1: 85 passed)"
echo "== device packing == host twin"; T4R_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_zz_mixed_head.py -q -x -k "packing" -p no:cacheprovider 2>&1 | tail -5
echo "== 2-unit product, element-wise (dense epilogue), then each partial product alone"; T4R_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_zz_mixed_head.py -q -k "materialised or partial_product" -p no:cacheprovider 2>&1 | tail -12
echo "== resident-A kernel vs default kernel (shipped arithmetic)"; T4R_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_zz_mixed_head.py -q -k "resident_head_kernel" -p no:cacheprovider 2>&1 | tail -8
echo "== fused head nprod=2 on single / pair / resident"; T4R_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_zz_mixed_head.py -q -k "full_softmax_nprod2 or requires or model_training or config2_full_size" -p no:cacheprovider 2>&1 | tail -15
echo "== two-warp tensor-path attention for 32 < L <= 64"; T4R_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_zz_attn64.py -q -p no:cacheprovider 2>&1 | tail -8
timeout 300 python tools/microbench.py attn50 2>&1 | tail -3
echo "== permutation language modeling (mask kernel, two-stream attention, model)"; T4R_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_zz_plm.py -q -p no:cacheprovider 2>&1 | tail -8
echo "== training step (N3): primitives vs host twins, gradients vs autograd, SGD steps"; T4R_TEST_EXPERIMENTAL=1 timeout 900 python -m pytest tests/test_gpu_zz_training.py -q -p no:cacheprovider 2>&1 | tail -8
echo "== memcheck of the new small kernels (packing, PLM mask, training primitives)"
T4R_TEST_EXPERIMENTAL=1 timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest -q -p no:cacheprovider \
  tests/test_gpu_zz_mixed_head.py::test_device_packing_matches_host_twin_bit_exactly tests/test_gpu_zz_plm.py::test_mask_kernel_matches_host_twin \
  tests/test_gpu_zz_training.py::test_training_primitives_device_vs_host_twin tests/test_gpu_zz_training.py::test_fused_adamw_on_gpu \
  tests/test_gpu_zz_training.py::test_widened_input_block_training_on_gpu 2>&1 | tail -6
echo "== timings (config-2 head shape): nprod 3 / 1, resident-A, nprod 2"; timeout 600 python tools/microbench.py head headres head2 2>&1 | tail -12
echo "== bench A/B"; for a in "" "--nprod 2"; do timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline $a 2>&1 | tail -1 | cut -c1-700; done
echo "== bench with the resident-A head"; T4R_HEAD_RESIDENT=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-700
T4R_HEAD_RESIDENT=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --nprod 2 2>&1 | tail -1 | cut -c1-700
echo "== training step timing (fwd + bwd + optimizer; not the BASELINE metric)"; for o in sgd adamw; do timeout 600 python bench.py --train --optimizer $o --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-500; done
} > gpurun_out/r2_first.log 2>&1
tail -60 gpurun_out/r2_first.log
