#!/bin/bash
# head backward on the tensor cores: GPU test suite (hard timeouts: a hand-rolled mbarrier pipeline can hang), then the
# same-box A/B against the SIMT heads
mkdir -p gpurun_out
rm -f gpurun_out/test_bars.jsonl
SR4D_RECORD_BARS=gpurun_out/test_bars.jsonl timeout -s KILL 1200 python -m pytest tests -m gpu -x -q --timeout 300 2>&1 | tail -8 | tee gpurun_out/pytest_gpu_head_bwd.txt
bash tools/gpu_ab_env.sh SR4D_HEAD_SIMT
