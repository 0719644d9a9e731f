"""Per-SASS-instruction executed counts of the first kernel of an ncu report, bucketed by source-line ranges.
python tools/ncu_sass_regions.py report.ncu-rep units file.cuh 'name:lo-hi,name:lo-hi,...' [dump.txt]"""
import collections
import csv
import subprocess
import sys

rep, units, fkey, spec = sys.argv[1], float(sys.argv[2]), sys.argv[3], sys.argv[4]
dump = sys.argv[5] if len(sys.argv) > 5 else None
regions = []
for part in spec.split(","):
    name, rng = part.split(":")
    lo, hi = rng.split("-")
    regions.append((name, int(lo), int(hi)))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
out, cur, fname, cols, seen, first_fn, fn = [], None, None, None, set(), None, None
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        fname = r[1].split('/')[-1]
        continue
    if r[0] == 'Function Name':
        fn = r[1]
        first_fn = first_fn or fn
        continue
    if r[0] == 'Line No':
        cols = {h: i for i, h in enumerate(r)}
        continue
    if fn != first_fn:
        continue
    if r[0] and r[2] == '-':
        cur = (fname, int(r[0]))
        continue
    if r[0] == '' and r[2].startswith('0x'):
        a = int(r[2], 16)
        if a in seen:
            continue
        seen.add(a)
        out.append((a, r[3].strip(), int(r[cols['Instructions Executed']]), cur))
out.sort()
base = out[0][0]
tot = sum(o[2] for o in out)
b, ops = collections.Counter(), collections.Counter()
for a, t, n, c in out:
    reg = 'inlined:' + c[0][:16]
    if c[0].startswith(fkey):
        reg = 'other'
        for name, lo, hi in regions:
            if lo <= c[1] <= hi:
                reg = name
                break
    b[reg] += n
    op = t.split()
    op = op[1] if op[0].startswith('@') else op[0]
    ops[op.split('.')[0]] += n
print("total %d warp-inst = %.1f per unit" % (tot, tot / units))
for k, v in sorted(b.items(), key=lambda x: -x[1]):
    print("%-28s %8.1f per unit %5.1f%%" % (k, v / units, 100 * v / tot))
print(", ".join("%s %.1f" % (k, v / units) for k, v in ops.most_common(24)))
if dump:
    with open(dump, 'w') as f:
        for a, t, n, c in out:
            f.write("%05x %9d %5.2f%% %-14s:%-4d %s\n" % (a - base, n, 100 * n / tot, c[0][:14], c[1], t))
