#!/bin/bash
# ncu launch lists (train step, forward) and full captures of the dominant kernels; run under gpurun
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_train.csv python tools/train_once.py 8 2 > gpurun_out/ncu_launch.log 2>&1; tail -2 gpurun_out/ncu_launch.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_fwd.csv python tools/fwd_once.py 8 2 > gpurun_out/ncu_launch_fwd.log 2>&1; tail -2 gpurun_out/ncu_launch_fwd.log
# forward: launches 30..59 are the second pass; 49.. are HR (TY=24)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv64_tc_kernel -s 50 -c 2 -f -o gpurun_out/prof_conv64_fwd_hr python tools/fwd_once.py 8 2 > gpurun_out/ncu_full1.log 2>&1; tail -2 gpurun_out/ncu_full1.log
# dgrad (TY=26 instantiation only): first HR launches of the second step
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:conv64_tc_kernel<26>" -s 30 -c 2 -f -o gpurun_out/prof_conv64_dgrad_hr python tools/train_once.py 8 2 > gpurun_out/ncu_full2.log 2>&1; tail -2 gpurun_out/ncu_full2.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wgrad64_tc_kernel -s 30 -c 2 -f -o gpurun_out/prof_wgrad64_hr python tools/train_once.py 8 2 > gpurun_out/ncu_full3.log 2>&1; tail -2 gpurun_out/ncu_full3.log
ls -la gpurun_out
