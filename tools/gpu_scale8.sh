#!/bin/bash
# 8-GPU check of the final build: bench.py under torchrun (weak scaling, dp_parity pre-flight inside)
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; tail -3 gpurun_out/bench_n8.err | cut -c1-300
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_n8.json').read().strip().splitlines()[-1])
print('8 gpus: value', round(d['value'], 1), 'ms/step', round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['value'], 1), 'fwd', round(d['forward']['value'], 1), d['clocks'], json.dumps(d.get('dp_parity'))[:300])
print({k: round(v['value'], 1) for k, v in d['other_configs'].items()})
PY
