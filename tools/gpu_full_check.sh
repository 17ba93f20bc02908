#!/bin/bash
# whole GPU suite (recording the gradient bars) + default bench + ncu launch list of two train steps
mkdir -p gpurun_out
rm -f gpurun_out/test_bars.jsonl
SR4D_RECORD_BARS=gpurun_out/test_bars.jsonl timeout -s KILL 1200 python -m pytest tests -m gpu -x -q --timeout 300 2>&1 | tail -6 | tee gpurun_out/pytest_gpu_full.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err || tail -5 gpurun_out/bench_full.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_full.json").read().strip().splitlines()[-1])
print("step", round(d["ms_per_step"], 3), "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), {a: round(b, 3) for a, b in d["kernel_classes_ms_per_step"].items()})
print("fwd", round(d["forward"]["ms_per_step"], 3), "clk", d["clocks"], {a: round(b["value"], 1) for a, b in d["other_configs"].items()})
print("roofline", {k: d["roofline"][k] for k in ("kernel", "achieved", "frac", "ms_per_launch")}, "cpu", d.get("cpu_baseline", {}).get("value"))
PY
bash tools/gpu_launches.sh | tail -40
