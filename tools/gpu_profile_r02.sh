#!/bin/bash
# round 2 evidence: ncu launch list of two train steps and --set full captures of the three tensor-core kernel classes
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_train.csv python tools/train_once.py 8 2 > gpurun_out/ncu_launch.log 2>&1; tail -1 gpurun_out/ncu_launch.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_fwd.csv python tools/fwd_once.py 8 2 > gpurun_out/ncu_launch_fwd.log 2>&1; tail -1 gpurun_out/ncu_launch_fwd.log
# conv64_tc_kernel launches of a train step: 0-18 LR forward, 19-29 HR forward, 30-32 HR head dgrads, 33.. HR block dgrads
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv64_tc_kernel -s 22 -c 1 -f -o gpurun_out/prof_r02_conv64_fwd_hr python tools/train_once.py 8 1 > gpurun_out/ncu_full1.log 2>&1; tail -1 gpurun_out/ncu_full1.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv64_tc_kernel -s 34 -c 1 -f -o gpurun_out/prof_r02_conv64_dgrad_hr python tools/train_once.py 8 1 > gpurun_out/ncu_full2.log 2>&1; tail -1 gpurun_out/ncu_full2.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wgrad64_tc2_kernel -s 4 -c 1 -f -o gpurun_out/prof_r02_wgrad2_hr python tools/train_once.py 8 1 > gpurun_out/ncu_full3.log 2>&1; tail -1 gpurun_out/ncu_full3.log
ls -la gpurun_out/*.ncu-rep
