#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:head2_bwd_kernel -s 3 -c 1 -f -o gpurun_out/prof_head2_bwd python tools/train_once.py 8 2 > gpurun_out/ncu_full5.log 2>&1; tail -2 gpurun_out/ncu_full5.log
