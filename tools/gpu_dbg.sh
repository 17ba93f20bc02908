#!/bin/bash
mkdir -p gpurun_out
SR4D_TC_DEBUG=1 timeout 300 python tools/train_once.py 8 1 2>&1 | sort | uniq -c | sort -rn | head -30 > gpurun_out/tc_dbg.log; cat gpurun_out/tc_dbg.log
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_n1.json'))
print('train ms', d['ms_per_step'], 'patches/s', d['value'], 'fwd ms', d['forward']['ms_per_step'], d['kernel_classes_ms_per_step'], d['clocks'])
PY
