#!/bin/bash
# same-box A/B of two builds: libsr4d.so (new) vs libsr4d_prev.so (previous commit), three alternations
mkdir -p gpurun_out
cp 4dflownet_b200/libsr4d.so /tmp/new.so
for rep in 1 2 3; do for which in prev new; do
  if [ $which = prev ]; then cp 4dflownet_b200/libsr4d_prev.so 4dflownet_b200/libsr4d.so; else cp /tmp/new.so 4dflownet_b200/libsr4d.so; fi
  timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extra-configs > gpurun_out/ab2_$which.json 2>/dev/null
  python - $which <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/ab2_{sys.argv[1]}.json").read().strip().splitlines()[-1])
print(sys.argv[1], "step", round(d["ms_per_step"], 3), {a: round(b, 3) for a, b in d["kernel_classes_ms_per_step"].items()}, "clk", d["clocks"]["sm_mhz"])
PY
done; done 2>&1 | tee gpurun_out/ab2.txt
cp /tmp/new.so 4dflownet_b200/libsr4d.so
timeout -s KILL 400 python -m pytest tests/test_gpu_forward.py tests/test_gpu_backward.py -m gpu -x -q --timeout 100 2>&1 | tail -3
