"""tests/gradient_bars.json from a recording of the GPU suite:
    SR4D_RECORD_BARS=gpurun_out/test_bars.jsonl python -m pytest tests -m gpu      (on a B200)
    python tools/make_gradient_bars.py gpurun_out/test_bars.jsonl
Every named gradient-parity check then asserts err <= 2 x the measured value (tests/conftest.py::bars)."""
import json
import os
import sys

src = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/test_bars.jsonl"
out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "gradient_bars.json")
table = {}
for line in open(src):
    r = json.loads(line)
    table[r["name"]] = max(table.get(r["name"], 0.0), float(r["err"]))
with open(out, "w") as f:
    json.dump(dict(sorted(table.items())), f, indent=1)
print(f"{len(table)} checks -> {out}")
