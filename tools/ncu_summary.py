"""Summarise ncu outputs (run here, no GPU needed).
  python tools/ncu_summary.py launches <launches.csv>          # per-kernel totals / shares
  python tools/ncu_summary.py raw <report.ncu-rep>              # key metrics per profiled launch
  python tools/ncu_summary.py stalls <report.ncu-rep> [N]       # top-N stalled SASS instructions
"""
import collections, csv, subprocess, sys, io

def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr, agg = None, collections.OrderedDict()
    for r in rows:
        if r[0] == 'ID':
            hdr = r; continue
        if hdr is None: continue
        d = dict(zip(hdr, r))
        try: v = float(d['Metric Value'].replace(',', ''))
        except ValueError: continue
        k = d['Kernel Name'].split('(')[0][-48:] + ' grid=' + d['Grid Size']
        a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"{'kernel':80s} {'n':>5s} {'total us':>10s} {'avg us':>9s} share")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:80s} {a[0]:5d} {a[1]/1e3:10.1f} {a[1]/a[0]/1e3:9.1f} {a[1]/tot:.3f}")
    print(f"total {tot/1e3:.1f} us")

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_subpipe_hmma_cycles_active', 'lts__t_bytes.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__m_xbar2l1tex_read_bytes.sum', 'lts__t_sector_hit_rate.pct', 'launch__registers_per_thread',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__cycles_elapsed.max', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sectors_srcunit_tex_op_write.sum', 'sm__inst_executed_pipe_uniform']

def raw(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print('==', d.get('Kernel Name', '')[:100], 'grid', d.get('Grid Size'))
        for k in hdr:
            if any(k == x or (k.startswith(x) and k[len(x):] in ('', '.per_second')) for x in KEYS):
                print(f"   {k:75s} {d[k]:>16s} {units[hdr.index(k)]}")

def stalls(path, n=30, which=0):
    out = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    # one section per profiled launch: a "Kernel Name" row, a header row, then one row per SASS instruction
    starts = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name'] + [len(rows)]
    lo, hi = starts[which], starts[which + 1]
    print('kernel:', rows[lo][1][:100])
    hdr = rows[lo + 1]; idx = {k: i for i, k in enumerate(hdr)}
    sec = [r for r in rows[lo + 2:hi] if len(r) >= len(hdr)]
    tot = sum(int(r[idx['# Samples']] or 0) for r in sec)
    print('total samples', tot)
    names = [k for k in hdr if k.startswith('stall_') and 'Not Issued' not in k]
    agg = collections.Counter()
    for r in sec:
        for k in names: agg[k] += int(r[idx[k]] or 0)
    print('by reason:', [(k, v) for k, v in agg.most_common(8)])
    for r in sorted(sec, key=lambda r: -int(r[idx['# Samples']] or 0))[:n]:
        st = {k: int(r[idx[k]] or 0) for k in names}
        main = sorted(((k, v) for k, v in st.items() if v), key=lambda kv: -kv[1])[:2]
        print(r[idx['# Samples']].rjust(7), r[idx['Instructions Executed']].rjust(9), r[idx['Address']][-5:], r[idx['Source']][:80].ljust(80), main)

if __name__ == '__main__':
    cmd = sys.argv[1]
    if cmd == 'launches': launches(sys.argv[2])
    elif cmd == 'raw': raw(sys.argv[2])
    else: stalls(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 30)
