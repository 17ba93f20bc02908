#!/bin/bash
# round 2, GPU call 4: full suite with bars, wgrad issue-loop timing, TC debug stamps, sanitizer, full bench line
mkdir -p gpurun_out
rm -f gpurun_out/test_bars.jsonl
SR4D_RECORD_BARS=gpurun_out/test_bars.jsonl timeout 1500 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_r02_4.txt
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra-configs > gpurun_out/c4_bench_quick.json 2>gpurun_out/c4.err || tail -5 gpurun_out/c4.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/c4_bench_quick.json").read().strip().splitlines()[-1])
print("step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["ms_per_step"], 3), {a: round(b, 3) for a, b in d["kernel_classes_ms_per_step"].items()}, "fwd", round(d["forward"]["ms_per_step"], 3))
PY
SR4D_TC_DEBUG=1 timeout 300 python tools/train_once.py 8 1 2> gpurun_out/tc_debug.txt | tail -1; grep -c "tc dbg" gpurun_out/tc_debug.txt; sort gpurun_out/tc_debug.txt | uniq -c | sort -rn | head -3 >/dev/null
python - <<'PY'
import re, collections
agg = collections.OrderedDict()
for line in open("gpurun_out/tc_debug.txt"):
    m = re.search(r"Do=(\d+) B=(\d+) ty=(\d+) dgrad=(\d+): MMA-warp wait cycles avg/CTA: t_empty (\d+)  x_full (\d+)  w_full (\d+)  of total (\d+) \| prologue (\d+)  tail after last MMA issue (\d+)  CTA lifetime (\d+) \| first entry -> last exit ([\d.]+) us", line)
    if not m: continue
    k = (m.group(1), m.group(3), m.group(4))
    a = agg.setdefault(k, [0] + [0.0] * 8)
    a[0] += 1
    for i in range(8): a[1 + i] += float(m.group(5 + i))
print("Do ty dgrad |  n | t_empty x_full w_full total | prologue tail lifetime | span us")
for k, a in agg.items():
    print(k, a[0], [round(v / a[0]) for v in a[1:8]], round(a[8] / a[0], 1))
PY
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_memcheck_smoke.txt 2>&1; tail -4 gpurun_out/sanitizer_memcheck_smoke.txt
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_racecheck_smoke.txt 2>&1; tail -4 gpurun_out/sanitizer_racecheck_smoke.txt
timeout 600 compute-sanitizer --tool synccheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_synccheck_smoke.txt 2>&1; tail -4 gpurun_out/sanitizer_synccheck_smoke.txt
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/c4_bench_full.json 2>gpurun_out/c4_full.err || tail -5 gpurun_out/c4_full.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/c4_bench_full.json").read().strip().splitlines()[-1])
print("FULL: value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "cpu", d.get("cpu_baseline", {}).get("value"), "roofline", {k: d["roofline"][k] for k in ("kernel", "achieved", "frac")})
for k, v in d["other_configs"].items(): print(k, {a: (round(b, 2) if isinstance(b, float) else b) for a, b in v.items() if a != "workload"})
PY
