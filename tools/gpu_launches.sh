#!/bin/bash
# ncu launch list (gpu__time_duration, no clock control) of two train steps + summary
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_train.csv python tools/train_once.py 8 2 > gpurun_out/ncu_launch.log 2>&1; tail -1 gpurun_out/ncu_launch.log
python tools/ncu_summary.py launches gpurun_out/launches_train.csv 2>/dev/null | head -40 | tee gpurun_out/launches_train_summary.txt
