#!/bin/bash
# round 2, GPU call 3: two-group dgrad epilogue + wgrad flush experiment + launch list + ncu captures
mkdir -p gpurun_out
rm -f gpurun_out/test_bars.jsonl
SR4D_RECORD_BARS=gpurun_out/test_bars.jsonl timeout 1200 python -m pytest tests/test_gpu_backward.py tests/test_gpu_forward.py tests/test_gpu_edge_cases.py -m gpu -q --timeout 300 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_r02_3.txt
for v in 0 100000 192; do
  SR4D_WGRAD_FLUSH=$v timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra-configs > gpurun_out/c3_flush_$v.json 2>gpurun_out/c3.err || tail -5 gpurun_out/c3.err
  python - "$v" gpurun_out/c3_flush_$v.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
print("flush", sys.argv[1], "step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["ms_per_step"], 3), {a: round(b, 3) for a, b in d["kernel_classes_ms_per_step"].items()}, "fwd", round(d["forward"]["ms_per_step"], 3))
PY
done 2>&1 | tee gpurun_out/c3_flush.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_train.csv python tools/train_once.py 8 2 > gpurun_out/ncu_launch.log 2>&1; tail -2 gpurun_out/ncu_launch.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wgrad64_tc2_kernel -s 30 -c 1 -f -o gpurun_out/prof_wgrad2_hr python tools/train_once.py 8 2 > gpurun_out/ncu_full3.log 2>&1; tail -2 gpurun_out/ncu_full3.log
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:conv64_tc_kernel<26" -s 30 -c 1 -f -o gpurun_out/prof_dgrad_single_hr python tools/train_once.py 8 2 > gpurun_out/ncu_full2.log 2>&1; tail -2 gpurun_out/ncu_full2.log
ls -la gpurun_out | tail -12
