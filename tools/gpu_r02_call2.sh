#!/bin/bash
# round 2, GPU call 2: the stacked single-plane wgrad kernel (wgrad_tc2.cu) and the chain-limited slabs of the two-plane
# kernel: (1) layer-level backward tests (small first: a hang would show here, under a short timeout),
# (2) the whole GPU suite with the measured gradient errors recorded, (3) gradient parity table, (4) bench A/B.
mkdir -p gpurun_out
rm -f gpurun_out/test_bars.jsonl
timeout 120 python -m pytest tests/test_gpu_backward.py -m gpu -x -q --timeout 100 -k "conv64_layer_bwd and 6-2" 2>&1 | tail -15
timeout 300 python -m pytest tests/test_gpu_backward.py -m gpu -q --timeout 100 -k "conv64_layer_bwd" 2>&1 | tail -25
SR4D_RECORD_BARS=gpurun_out/test_bars.jsonl timeout 1200 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -40 | tee gpurun_out/pytest_gpu_r02_2.txt
timeout 1500 python tools/grad_parity.py --out gpurun_out/grad_parity.txt > gpurun_out/grad_parity.log 2>&1; grep -v "^  conv3d" gpurun_out/grad_parity.txt | head -60
for rep in 1 2; do for m in "" "--two-plane-backward"; do
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline $m > gpurun_out/ab_$rep"_"${m:+two}.json 2>gpurun_out/ab.err || tail -5 gpurun_out/ab.err
  python - "$m" gpurun_out/ab_$rep"_"${m:+two}.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
print("flags", repr(sys.argv[1]), "step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["ms_per_step"], 3), {a: round(b, 3) for a, b in d["kernel_classes_ms_per_step"].items()}, "fwd", round(d["forward"]["ms_per_step"], 3))
PY
done; done 2>&1 | tee gpurun_out/ab_r02_2.txt
