#!/bin/bash
# same-box A/B of two builds of libsr4d.so: tools/probe/libsr4d_prev.so (a previous commit, built by hand) vs the in-tree build
mkdir -p gpurun_out
cp 4dflownet_b200/libsr4d.so /tmp/libsr4d_new.so
for rep in 1 2; do for m in prev new; do
  if [ $m = prev ]; then cp tools/probe/libsr4d_prev.so 4dflownet_b200/libsr4d.so; else cp /tmp/libsr4d_new.so 4dflownet_b200/libsr4d.so; fi
  timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/ab_$m.json 2>/dev/null
  python - $m <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/ab_{sys.argv[1]}.json").read().strip().splitlines()[-1])
k = d["kernel_classes_ms_per_step"]
print("build", sys.argv[1], "step", round(d["ms_per_step"], 3), "fwd", round(d["forward"]["ms_per_step"], 3), {a: round(b, 3) for a, b in k.items()})
PY
done; done 2>&1 | tee gpurun_out/ab_builds.txt
cp /tmp/libsr4d_new.so 4dflownet_b200/libsr4d.so
