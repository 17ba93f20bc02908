#!/bin/bash
# iteration call: backward + edge-case tests, short bench, ncu launch list of two train steps
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_backward.py tests/test_gpu_edge_cases.py tests/test_gpu_forward.py -m gpu -x -q --timeout 300 2>&1 | tail -5
for i in 1 2; do
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra-configs > gpurun_out/quick.json 2>gpurun_out/quick.err || tail -5 gpurun_out/quick.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/quick.json").read().strip().splitlines()[-1])
print("step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["ms_per_step"], 3), {a: round(b, 3) for a, b in d["kernel_classes_ms_per_step"].items()}, "fwd", round(d["forward"]["ms_per_step"], 3), "clk", d["clocks"]["sm_mhz"])
PY
done
bash tools/gpu_launches.sh | tail -42
