#!/bin/bash
# chained forward launches: parity tests (hard timeouts: the dependency waits can hang), then same-box A/B with SR4D_NO_CHAIN
mkdir -p gpurun_out
timeout -s KILL 150 python -m pytest tests/test_gpu_forward.py -m gpu -x -q --timeout 60 2>&1 | tail -6 | tee gpurun_out/chain_tests.txt
if grep -q "passed" gpurun_out/chain_tests.txt && ! grep -q "failed\|error" gpurun_out/chain_tests.txt; then
  timeout -s KILL 300 python -m pytest tests/test_gpu_integration.py tests/test_gpu_edge_cases.py -m gpu -x -q --timeout 100 2>&1 | tail -4
  timeout -s KILL 400 bash tools/gpu_ab_env.sh SR4D_NO_CHAIN
  for n in 8 64; do echo "SR4D_CHAIN_TILES_PER_SM=$n"; SR4D_CHAIN_TILES_PER_SM=$n timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/chain_$n.json 2>/dev/null; python - $n <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/chain_{sys.argv[1]}.json").read().strip().splitlines()[-1])
print("step", round(d["ms_per_step"], 3), "fwd", round(d["forward"]["ms_per_step"], 3), {a: round(b["value"], 1) for a, b in d["other_configs"].items()})
PY
  done
fi
