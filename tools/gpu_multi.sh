#!/bin/bash
# 2-GPU (or N-GPU) checks: DP parity test + torchrun bench lines
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -8
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -3 gpurun_out/bench_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; tail -5 gpurun_out/bench_n$N.err
python - <<PY
import json
for n in (1, $N):
    try:
        d = json.loads(open(f'gpurun_out/bench_n{n}.json').read().strip().splitlines()[-1])
        print(n, 'gpus: value', round(d['value'], 1), 'ms/step', round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['value'], 1), 'fwd', round(d['forward']['value'], 1))
    except Exception as e:
        print(n, 'failed', e)
PY
