#!/bin/bash
# first thing to run in the next round: validate the experimental single-operand backward options on hardware
# (gated parity tests), then a same-box A/B of the bench with and without them
mkdir -p gpurun_out
SR4D_TEST_EXPERIMENTAL=1 timeout 400 python -m pytest tests/test_gpu_backward.py -m gpu -q --timeout 150 -k single_operand 2>&1 | tail -8
for rep in 1 2; do for m in "" "--experimental-backward"; do
  timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline $m > gpurun_out/exp_bwd.json 2>/dev/null
  python - "$m" <<'PY'
import json, sys
d = json.loads(open("gpurun_out/exp_bwd.json").read().strip().splitlines()[-1])
print("flags", repr(sys.argv[1]), "step", round(d["ms_per_step"], 3), {a: round(b, 3) for a, b in d["kernel_classes_ms_per_step"].items()})
PY
done; done 2>&1 | tee gpurun_out/exp_bwd.txt
