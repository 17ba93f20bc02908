#!/bin/bash
# same-box A/B of the conv producer order (SR4D_TC_XSPLIT=0: plane requested at the pass boundary; 1: mid-pass)
mkdir -p gpurun_out
for rep in 1 2; do for m in 0 1; do
  SR4D_TC_XSPLIT=$m timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/ab_$m.json 2>/dev/null
  python - $m <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/ab_{sys.argv[1]}.json").read().strip().splitlines()[-1])
k = d["kernel_classes_ms_per_step"]
print("xsplit", sys.argv[1], "step", round(d["ms_per_step"], 3), "fwd", round(d["forward"]["ms_per_step"], 3), {a: round(b, 3) for a, b in k.items()})
PY
done; done 2>&1 | tee gpurun_out/ab.txt
