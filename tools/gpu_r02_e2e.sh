#!/bin/bash
# e2e upload overlap (HR targets on a side stream underneath the forward, deferred metric fold): new equality test, the whole gpu suite, bench
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_edge_cases.py -m gpu -x -q --timeout 200 -k "host_buffers or reduces_the_loss" 2>&1 | tail -5
timeout 600 python -m pytest tests -m gpu -x -q --timeout 200 2>&1 | tail -6 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
( time timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err ) 2>&1 | grep real; tail -2 gpurun_out/bench_n1.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1])
print('value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value'], 1), round(d['e2e']['ms_per_step'], 3), d['clocks'])
print(d.get('kernel_classes_ms_per_step'))
PY
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
