#!/bin/bash
# thin tiles of the fused dgrad (last interior z column + halo column as a 4-row-per-line tile): parity tests, A/B
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_backward.py tests/test_gpu_edge_cases.py -m gpu -x -q --timeout 150 2>&1 | tail -6
for v in 0 1; do
  SR4D_DGRAD_THIN=$v timeout 200 python bench.py --steps 10 --warmup 3 > gpurun_out/thin_$v.json 2> gpurun_out/thin_$v.err; tail -1 gpurun_out/thin_$v.err | cut -c1-200
done
python - <<'PY'
import json
for v in (0, 1):
    d = json.loads(open(f'gpurun_out/thin_{v}.json').read().strip().splitlines()[-1])
    print('SR4D_DGRAD_THIN', v, 'value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 3), d['clocks']['sm_mhz'], d.get('kernel_classes_ms_per_step'))
PY
