#!/bin/bash
# parity tests of the new build, then same-box A/B of two builds: libsr4d_prev.so (previous commit) vs libsr4d.so (new), three alternations
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_backward.py -m gpu -x -q --timeout 150 -k "train_step or identical" 2>&1 | tail -5
cp 4dflownet_b200/libsr4d.so /tmp/new.so
for rep in 1 2; do for which in prev new; do
  if [ $which = prev ]; then cp 4dflownet_b200/libsr4d_prev.so 4dflownet_b200/libsr4d.so; else cp /tmp/new.so 4dflownet_b200/libsr4d.so; fi
  timeout -s KILL 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extra-configs > gpurun_out/ab2_$which.json 2>/dev/null
  python - $which <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/ab2_{sys.argv[1]}.json").read().strip().splitlines()[-1])
print(sys.argv[1], "step", round(d["ms_per_step"], 3), {a: round(b, 3) for a, b in d["kernel_classes_ms_per_step"].items()}, "fwd", round(d["forward"]["ms_per_step"], 3), "clk", d["clocks"]["sm_mhz"])
PY
done; done 2>&1 | tee gpurun_out/ab_builds.txt
cp /tmp/new.so 4dflownet_b200/libsr4d.so
