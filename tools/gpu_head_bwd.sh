#!/bin/bash
# head backward on the tensor cores: isolated parity test first (hard timeout: a hand-rolled mbarrier pipeline can hang),
# then the train-step tests, then the same-box A/B against the SIMT heads
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_gpu_backward.py -k "head_layer_bwd" -m gpu -x -q --timeout 120 2>&1 | tail -25 | tee gpurun_out/head_bwd_test.txt
if grep -q "passed" gpurun_out/head_bwd_test.txt && ! grep -q "failed" gpurun_out/head_bwd_test.txt; then
  timeout -s KILL 900 python -m pytest tests/test_gpu_backward.py tests/test_gpu_edge_cases.py -m gpu -x -q --timeout 300 2>&1 | tail -8 | tee gpurun_out/head_bwd_test2.txt
  bash tools/gpu_ab_env.sh SR4D_HEAD_SIMT
fi
