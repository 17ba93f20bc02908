#!/bin/bash
# 2-GPU validation: NCCL data-parallel parity test + bench at N=2 (dp_parity preflight inside)
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q --timeout 300 2>&1 | tail -5 | tee gpurun_out/pytest_gpu_multi_r02.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err || tail -20 gpurun_out/bench_n2.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_n2.json").read().strip().splitlines()[-1])
print("N=2: value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "ms", round(d["ms_per_step"], 3), "dp_parity", json.dumps(d.get("dp_parity")))
for k, v in d["other_configs"].items(): print(k, round(v["value"], 1), v["unit"])
PY
