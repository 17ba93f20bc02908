#!/bin/bash
# same-box A/B of one build with / without an environment switch:  bash tools/gpu_ab_env.sh SR4D_HEAD_SIMT
mkdir -p gpurun_out
V=${1:-SR4D_HEAD_SIMT}
for rep in 1 2; do for on in 0 1; do
  if [ $on = 1 ]; then export $V=1; else unset $V; fi
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/abenv_$on.json 2>/dev/null
  python - $V $on <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/abenv_{sys.argv[2]}.json").read().strip().splitlines()[-1])
k = d["kernel_classes_ms_per_step"]
print(sys.argv[1], "=", sys.argv[2], "step", round(d["ms_per_step"], 3), "fwd", round(d["forward"]["ms_per_step"], 3), "conv classes", round(sum(k.values()), 3),
      {a: (round(b["value"], 1)) for a, b in d["other_configs"].items()})
PY
done; done 2>&1 | tee gpurun_out/abenv_$V.txt
