#!/bin/bash
# batched weight gradient: backward / edge-case / integration tests with hard timeouts, then the A/B against per-layer launches
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_backward.py tests/test_gpu_edge_cases.py -m gpu -x -q --timeout 200 2>&1 | tail -6 | tee gpurun_out/wgb_tests.txt
if grep -q "passed" gpurun_out/wgb_tests.txt && ! grep -q "failed\|error" gpurun_out/wgb_tests.txt; then
  timeout -s KILL 300 python -m pytest tests/test_gpu_integration.py tests/test_gpu_c_client.py -m gpu -x -q --timeout 100 2>&1 | tail -3
  timeout -s KILL 500 bash tools/gpu_ab_env.sh SR4D_WGRAD_UNBATCHED
fi
