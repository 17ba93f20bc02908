#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k "regex:stem_conv_kernel|conv1x1_cat_kernel|upsample_kernel" -s 4 -c 4 -f -o gpurun_out/prof_small_fwd python tools/fwd_once.py 8 2 > gpurun_out/ncu_full6.log 2>&1; tail -2 gpurun_out/ncu_full6.log
