"""Summarise an ncu report per CUDA source line: python tools/ncu_lines.py report.ncu-rep [min_pct]"""
import collections as C
import csv
import subprocess
import sys

rep = sys.argv[1]
minpct = float(sys.argv[2]) if len(sys.argv) > 2 else 0.6
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'launch__grid_size', 'smsp__inst_executed.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct']
stall = [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('_per_issue_active.ratio')]
for r in rows[2:3]:
    print('---', r[hdr.index('Kernel Name')][:60])
    for k in keys:
        if k in hdr:
            print('  ', k, r[hdr.index(k)], rows[1][hdr.index(k)])
    st = sorted(((float(r[hdr.index(k)]), k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')) for k in stall), reverse=True)
    print('   stalls:', ', '.join('%s %.2f' % (k, v) for v, k in st[:8]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
agg = C.OrderedDict()
for r in rows[3:]:
    if len(r) < 8:
        continue
    if r[0] == 'Line No':
        break
    if r[0] and r[2] == '-':
        try:
            agg[int(r[0])] = (r[1].strip(), int(r[7]), int(r[6]))
        except ValueError:
            pass
tot = sum(v[1] for v in agg.values()) or 1
ts = sum(v[2] for v in agg.values()) or 1
print('total warp-inst', tot, 'samples', ts)
for ln, (s, n, sm) in sorted(agg.items()):
    if 100.0 * n / tot >= minpct or 100.0 * sm / ts >= 2 * minpct:
        print(f"{ln:4d} {100*n/tot:5.1f}% inst {100*sm/ts:5.1f}% smp | {s[:115]}")
