#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_forward.py -x -q --timeout 120 2>&1 | tail -3
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_fwd.csv python tools/fwd_once.py 8 2 > gpurun_out/ncu_launch_fwd.log 2>&1; tail -2 gpurun_out/ncu_launch_fwd.log
