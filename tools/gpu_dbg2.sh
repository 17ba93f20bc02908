#!/bin/bash
mkdir -p gpurun_out
SR4D_TC_DEBUG=1 timeout 200 python tools/train_once.py 8 1 2>&1 | grep "tc dbg" | sort | uniq -c | sort -rn | awk '{$1="";print}' | sort -u -k3,6 | head -12 > gpurun_out/tc_dbg.log; cat gpurun_out/tc_dbg.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_train.csv python tools/train_once.py 8 2 > gpurun_out/ncu_launch.log 2>&1; tail -2 gpurun_out/ncu_launch.log
