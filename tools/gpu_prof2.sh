#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:head_out_kernel -s 1 -c 1 -f -o gpurun_out/prof_head_out python tools/fwd_once.py 8 2 > gpurun_out/ncu_full4.log 2>&1; tail -2 gpurun_out/ncu_full4.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv64_tc_kernel -s 98 -c 2 -f -o gpurun_out/prof_conv64_dgrad_hr python tools/train_once.py 8 2 > gpurun_out/ncu_full2.log 2>&1; tail -2 gpurun_out/ncu_full2.log
