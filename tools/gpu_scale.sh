#!/bin/bash
# scaling check: bench.py at N GPUs under torchrun (+ config 5 sharded inference)
N=${1:-8}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; tail -3 gpurun_out/bench_n$N.err | cut -c1-300
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 tools/bench_configs.py --config 5 2> gpurun_out/cfg5_n$N.err | tail -1 > gpurun_out/cfg5_n$N.json; cat gpurun_out/cfg5_n$N.json | cut -c1-600
python - <<PY
import json
d = json.loads(open('gpurun_out/bench_n$N.json').read().strip().splitlines()[-1])
print($N, 'gpus: value', round(d['value'], 1), 'ms/step', round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['value'], 1), 'fwd', round(d['forward']['value'], 1), d['clocks'])
PY
