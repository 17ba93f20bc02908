#!/bin/bash
# ncu --set full of one launch of a kernel inside a train step:  bash tools/gpu_prof_kernel.sh <kernel regex> <skip> <out name>
mkdir -p gpurun_out
K=${1:-head_bwd_tc_kernel}; S=${2:-1}; O=${3:-prof_$K}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s $S -c 1 -f -o gpurun_out/$O python tools/train_once.py 8 1 > gpurun_out/ncu_$O.log 2>&1; tail -1 gpurun_out/ncu_$O.log
ls -la gpurun_out/$O.ncu-rep
