#!/bin/bash
# round 2, GPU call 1: (1) the gated single-operand backward parity tests, (2) gradient parity table at the
# config-2 geometry (fp64 oracle on the host), (3) same-box bench A/B default vs --experimental-backward,
# (4) the whole GPU suite
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader
nproc; free -g | head -2
SR4D_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_backward.py -m gpu -q --timeout 200 -k single_operand 2>&1 | tail -15 | tee gpurun_out/exp_tests.txt
timeout 1500 python tools/grad_parity.py --cases "${CASES:-24,2,8,4,1;24,2,8,4,8}" --out gpurun_out/grad_parity.txt 2>&1 | tail -140
for rep in 1 2; do for m in "" "--experimental-backward"; do
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline $m > gpurun_out/exp_bwd_$rep"_"${m:+exp}.json 2>gpurun_out/exp_bwd.err
  python - "$m" gpurun_out/exp_bwd_$rep"_"${m:+exp}.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
print("flags", repr(sys.argv[1]), "step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["ms_per_step"], 3), {a: round(b, 3) for a, b in d["kernel_classes_ms_per_step"].items()}, "fwd", round(d["forward"]["ms_per_step"], 3))
PY
done; done 2>&1 | tee gpurun_out/exp_bwd.txt
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_r02_1.txt
