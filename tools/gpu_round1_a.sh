#!/bin/bash
# First GPU call of the round: parity tests, bench (both arms), ncu launch list and full captures.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; cat gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
# launch list of exactly one warm train step (step 2 of 2)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_train.csv python tools/train_once.py 8 2 > gpurun_out/ncu_launch.log 2>&1; tail -3 gpurun_out/ncu_launch.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_fwd.csv python tools/fwd_once.py 8 2 > gpurun_out/ncu_launch_fwd.log 2>&1; tail -3 gpurun_out/ncu_launch_fwd.log
# full captures: the HR conv forward, dgrad, wgrad (skip the first step's launches)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv64_tc_kernel -s 48 -c 3 -f -o gpurun_out/prof_conv64_tc_fwd python tools/fwd_once.py 8 2 > gpurun_out/ncu_full1.log 2>&1; tail -3 gpurun_out/ncu_full1.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wgrad64_tc_kernel -s 30 -c 2 -f -o gpurun_out/prof_wgrad64_tc python tools/train_once.py 8 2 > gpurun_out/ncu_full2.log 2>&1; tail -3 gpurun_out/ncu_full2.log
ls -la gpurun_out
