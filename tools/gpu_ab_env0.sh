#!/bin/bash
# same-box A/B of one build with an environment switch set to 0 (feature off) vs unset (default):  bash tools/gpu_ab_env0.sh SR4D_TC_CLUSTER
mkdir -p gpurun_out
V=${1:-SR4D_TC_CLUSTER}
for rep in 1 2 3; do for off in 1 0; do
  if [ $off = 1 ]; then export $V=0; else unset $V; fi
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/abenv0_$off.json 2>/dev/null
  python - $V $off <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/abenv0_{sys.argv[2]}.json").read().strip().splitlines()[-1])
k = d["kernel_classes_ms_per_step"]
print(sys.argv[1], "off" if sys.argv[2] == "1" else "default", "step", round(d["ms_per_step"], 3), "fwd", round(d["forward"]["ms_per_step"], 3), "conv classes", round(sum(k.values()), 3),
      {a: (round(b["value"], 1)) for a, b in d["other_configs"].items()}, "clk", d["clocks"]["sm_mhz"])
PY
done; done 2>&1 | tee gpurun_out/abenv0_$V.txt
