#!/bin/bash
# chained forward launches: parity tests (hard timeouts: a grid barrier can hang), then same-box A/B with SR4D_NO_CHAIN
mkdir -p gpurun_out
timeout -s KILL 150 python -m pytest tests/test_gpu_forward.py -m gpu -x -q --timeout 60 2>&1 | tail -6 | tee gpurun_out/chain_tests.txt
if grep -q "passed" gpurun_out/chain_tests.txt && ! grep -q "failed\|error" gpurun_out/chain_tests.txt; then
  timeout -s KILL 300 python -m pytest tests/test_gpu_integration.py tests/test_gpu_edge_cases.py -m gpu -x -q --timeout 100 2>&1 | tail -4
  timeout -s KILL 400 bash tools/gpu_ab_env.sh SR4D_NO_CHAIN
fi
