#!/bin/bash
# one optimisation iteration: parity tests that touch the changed kernels, MMA-warp wait breakdown, sustained forward loop, default bench
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_forward.py tests/test_gpu_backward.py tests/test_gpu_edge_cases.py tests/test_graph_golden.py -m gpu -x -q --timeout 120 2>&1 | tail -12 > gpurun_out/pytest_iter.log; cat gpurun_out/pytest_iter.log
SR4D_TC_DEBUG=1 timeout 200 python tools/fwd_once.py 8 2 2>&1 | grep "tc dbg" | tail -30 | grep "B=8" | sort -u -k3,6 | cut -c1-330 > gpurun_out/tc_dbg_iter.txt; cat gpurun_out/tc_dbg_iter.txt
timeout 120 python tools/exp_skip.py 150 2>&1 | tail -1 | tee gpurun_out/fwd_loop_iter.txt
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_iter.json 2> gpurun_out/bench_iter.err; tail -3 gpurun_out/bench_iter.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_iter.json").read().strip().splitlines()[-1])
print("ms_per_step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"])
print(d["kernel_classes_ms_per_step"])
print("fwd", d["forward"]["ms_per_step"], d["clocks"])
PY
