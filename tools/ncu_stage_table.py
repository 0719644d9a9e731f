"""Per-kernel table from an ncu --set full report: duration, DRAM bytes, achieved GB/s against the measured peak,
issue utilisation, occupancy, top stall reasons.  python tools/ncu_stage_table.py report.ncu-rep [peak_gbs]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
peak = float(sys.argv[2]) if len(sys.argv) > 2 else 6573.5
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}


def val(r, k):
    try:
        v = float(r[col[k]].replace(",", ""))
    except (KeyError, ValueError):
        return float("nan")
    u = units[col[k]]
    scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}.get(u, 1.0)
    return v * scale


stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
print("%-44s %9s %9s %9s %8s %6s %6s %6s  %s" % ("kernel", "us", "rd MB", "wr MB", "GB/s", "frac", "issue%", "occ%", "top stalls (warps per issue)"))
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    name = r[col["Kernel Name"]].split("(")[0][:44]
    us = val(r, "gpu__time_duration.sum")
    rd, wr = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
    gbs = (rd + wr) / us / 1e3 if us else 0.0
    st = sorted(((float(r[col[k]]), k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""))
                 for k in stall), reverse=True)[:3]
    print("%-44s %9.1f %9.2f %9.2f %8.0f %6.3f %6.1f %6.1f  %s" % (
        name, us, rd / 1e6, wr / 1e6, gbs, gbs / peak, val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        val(r, "sm__warps_active.avg.pct_of_peak_sustained_active"), ", ".join("%s %.1f" % (k, v) for v, k in st)))
