#!/bin/bash
# CTA-pair weight multicast (SR4D_TC_CLUSTER=1): parity tests under the switch with hard timeouts, then the same-box A/B
mkdir -p gpurun_out
SR4D_TC_CLUSTER=1 timeout -s KILL 200 python -m pytest tests/test_gpu_forward.py -m gpu -x -q --timeout 60 2>&1 | tail -4 | tee gpurun_out/cluster_tests.txt
if grep -q "passed" gpurun_out/cluster_tests.txt && ! grep -q "failed\|error" gpurun_out/cluster_tests.txt; then
  SR4D_TC_CLUSTER=1 timeout -s KILL 400 python -m pytest tests/test_gpu_backward.py tests/test_gpu_edge_cases.py -m gpu -x -q --timeout 100 2>&1 | tail -3
  timeout -s KILL 500 bash tools/gpu_ab_env.sh SR4D_TC_CLUSTER
fi
