#!/bin/bash
# small-kernel iteration: forward/backward parity tests + ncu launch list of two B=8 forwards and one train step
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_forward.py tests/test_gpu_backward.py tests/test_graph_golden.py -m gpu -x -q --timeout 120 2>&1 | tail -6
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_small.csv python tools/train_once.py 8 1 > gpurun_out/ncu_small.log 2>&1; tail -2 gpurun_out/ncu_small.log | cut -c1-200
python tools/ncu_summary.py launches gpurun_out/launches_small.csv | grep -v "conv64_tc\|wgrad64_tc" | head -24
