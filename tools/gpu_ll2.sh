#!/bin/bash
mkdir -p gpurun_out
timeout 420 python -m pytest tests -m gpu -x -q --timeout 150 2>&1 | tail -4
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_train.csv python tools/train_once.py 8 2 > gpurun_out/ncu_launch.log 2>&1; tail -2 gpurun_out/ncu_launch.log
