#!/bin/bash
mkdir -p gpurun_out
for rep in 1 2; do for m in 0 1 2 3; do SR4D_TC_EXP_SKIP=$m timeout 120 python tools/exp_skip.py 150 2>&1 | tail -1; done; done | tee gpurun_out/exp_skip.txt
