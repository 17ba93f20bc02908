#!/bin/bash
# final check of the round: smoke, the whole gpu suite, both bench arms (no profiler), sanitizer on smoke
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
( time timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err ) 2>&1 | grep real; cut -c1-300 gpurun_out/bench_ref.json
timeout 400 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_memcheck_smoke.txt 2>&1; tail -3 gpurun_out/sanitizer_memcheck_smoke.txt
timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_racecheck_smoke.txt 2>&1; tail -3 gpurun_out/sanitizer_racecheck_smoke.txt
timeout 400 compute-sanitizer --tool synccheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_synccheck_smoke.txt 2>&1; tail -3 gpurun_out/sanitizer_synccheck_smoke.txt
