"""Summarise an ncu report per CUDA source line (all source files of the kernel):
python tools/ncu_lines.py report.ncu-rep [min_pct] [kernel-name-substring]   (default: the first kernel of the report)"""
import csv
import os
import subprocess
import sys

rep = sys.argv[1]
minpct = float(sys.argv[2]) if len(sys.argv) > 2 else 0.6
want = sys.argv[3] if len(sys.argv) > 3 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'launch__grid_size', 'smsp__inst_executed.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']
stall = [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('_per_issue_active.ratio')]
sel = [r for r in rows[2:] if want is None or want in r[hdr.index('Kernel Name')]][:1]
for r in sel:
    print('---', r[hdr.index('Kernel Name')][:60])
    for k in keys:
        if k in hdr:
            print('  ', k, r[hdr.index(k)], rows[1][hdr.index(k)])
    st = sorted(((float(r[hdr.index(k)]), k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')) for k in stall), reverse=True)
    print('   stalls:', ', '.join('%s %.2f' % (k, v) for v, k in st[:8]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
agg = {}
fname = "?"
cols = None
first_fn = None
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        fname = os.path.basename(r[1])
        continue
    if r[0] == 'Function Name':
        if first_fn is None and (want is None or want in r[1]):
            first_fn = r[1]
        cur_fn = r[1]
        continue
    if r[0] == 'Line No':
        cols = {h: i for i, h in enumerate(r)}
        continue
    if cols is None or len(r) < 8 or cur_fn != first_fn:
        continue
    if r[0] and r[2] == '-':
        try:
            agg[(fname, int(r[0]))] = (r[1].strip(), int(r[cols['Instructions Executed']]), int(r[cols['# Samples']]))
        except ValueError:
            pass
tot = sum(v[1] for v in agg.values()) or 1
ts = sum(v[2] for v in agg.values()) or 1
print('total warp-inst', tot, 'samples', ts)
for (f, ln), (s, n, sm) in sorted(agg.items()):
    if 100.0 * n / tot >= minpct or 100.0 * sm / ts >= 2 * minpct:
        print(f"{f[:12]:12s}{ln:5d} {100*n/tot:5.1f}% inst {100*sm/ts:5.1f}% smp | {s[:105]}")
