#!/bin/bash
# 2-GPU check: bench.py under torchrun (weak scaling, dp_parity pre-flight inside) + the 2-GPU DP parity tests
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -3 gpurun_out/bench_n2.err | cut -c1-300
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_n2.json').read().strip().splitlines()[-1])
print('2 gpus: value', round(d['value'], 1), 'ms/step', round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['value'], 1), round(d['e2e']['ms_per_step'], 2), 'fwd', round(d['forward']['value'], 1), d['clocks'], json.dumps(d.get('dp_parity'))[:300])
print({k: round(v['value'], 1) for k, v in d['other_configs'].items()})
PY
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -x -q --timeout 200 2>&1 | tail -3
