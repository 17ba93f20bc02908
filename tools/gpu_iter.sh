#!/bin/bash
# one optimisation iteration: backward parity tests, then the default bench (prints the per-class kernel times)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_backward.py tests/test_gpu_edge_cases.py -m gpu -x -q --timeout 120 2>&1 | tail -12 > gpurun_out/pytest_iter.log; cat gpurun_out/pytest_iter.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_iter.json 2> gpurun_out/bench_iter.err; tail -3 gpurun_out/bench_iter.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_iter.json").read().strip().splitlines()[-1])
print("ms_per_step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"])
print(d["kernel_classes_ms_per_step"])
print("fwd", d["forward"]["ms_per_step"], d["clocks"])
PY
