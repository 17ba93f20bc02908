#!/bin/bash
# SR4D_TC_DEBUG cycle stamps of the MMA warps (conv fwd / dgrad, stacked wgrad) for one batch-8 train step
mkdir -p gpurun_out
SR4D_WGRAD_UNBATCHED=1 SR4D_TC_DEBUG=1 timeout 300 python tools/train_once.py 8 1 2> gpurun_out/tc_debug.txt | tail -1; grep -c "dbg" gpurun_out/tc_debug.txt
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
grep "wgrad2 dbg" gpurun_out/tc_debug.txt | sort | uniq -c | sort -rn | head -4 | cut -c1-260
