#!/bin/bash
# final check of the session: smoke, the whole gpu suite, bench (no profiler), then the launch lists of the final build
mkdir -p gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python -m pytest tests -m gpu -x -q --timeout 200 2>&1 | tail -6 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
( time timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err ) 2>&1 | grep real; tail -2 gpurun_out/bench_n1.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1])
print('value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value'], 1), 'fwd', round(d['forward']['value'], 1), d['clocks'])
print(d.get('kernel_classes_ms_per_step'))
print({k: round(v['value'], 1) for k, v in d['other_configs'].items()})
PY
timeout 120 python tools/time_fwd.py 1 2>&1 | tail -2
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_fwd_b1.csv python tools/fwd_once.py 1 3 > gpurun_out/ncu_b1.log 2>&1; tail -1 gpurun_out/ncu_b1.log
python tools/ncu_summary.py launches gpurun_out/launches_fwd_b1.csv | head -20
