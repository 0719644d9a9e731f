#!/bin/bash
# Summarise the ncu reports / launch lists / bench lines of one GPU call (gpurun_out/<tag>_*) into profiles/.
# usage: tools/make_profiles.sh <tag of the assign/launch/bench files> [tag of the stage-kernel report]
set -e
cd "$(dirname "$0")/.."
T=${1:-r02h}
S=${2:-r02}
O=gpurun_out
P=${PROFILES_DIR:-profiles}
mkdir -p $P
[ -f $P/assign_traffic.json ] || cp profiles/assign_traffic.json $P/assign_traffic.json
{ echo "# ncu --set full, one launch each, 720p x 32 (clock control none); GB/s = (dram read + write) / duration; frac of MEASURED_PEAKS hbm_gbs 6573.5";
  python tools/ncu_stage_table.py $O/${S}_stages_b32.ncu-rep;
  if [ -f $O/${T}_lab_b32.ncu-rep ]; then echo "# the Lab kernel of the final build"; python tools/ncu_stage_table.py $O/${T}_lab_b32.ncu-rep | tail -n +2; fi; } > $P/r02_stage_kernels_b32.txt
{ echo "# ncu --set full, one launch each, 3840x2160 K=4000 x 8"; python tools/ncu_stage_table.py $O/${S}_stages_D8.ncu-rep; } > $P/r02_stage_kernels_D8.txt
{ python tools/ncu_lines.py $O/${T}_assign5_b32.ncu-rep 1.0;
  echo; echo "# warp instructions per 128-pixel tile by region of assign5.cuh (76 800 tiles per launch)";
  python tools/ncu_sass_regions.py $O/${T}_assign5_b32.ncu-rep 76800 assign5 'prologue:129-177,walk:178-194,list:195-278,wait:279-289,tilesetup:290-327,distance:328-366,labels:367-387,update:388-482,out:483-516,prepare_tail:517-539,helpers:50-128'; } > $P/r02_assign5_b32_ncu_summary.txt
[ -f $O/${T}_assign5_D8.ncu-rep ] && python tools/ncu_lines.py $O/${T}_assign5_D8.ncu-rep 1.5 > $P/r02_assign5_D8_ncu_summary.txt
python tools/summarize_launches.py $O/${T}_launches_b32.csv > $P/r02_launches_b32_summary.txt
cp $O/${T}_launches_b32.csv $P/r02_launches_b32.csv
if [ -f $O/${T}_launches_b1.csv ]; then python tools/summarize_launches.py $O/${T}_launches_b1.csv > $P/r02_launches_b1_summary.txt; cp $O/${T}_launches_b1.csv $P/r02_launches_b1.csv; fi
# SASS evidence of the TMA path
{ echo "# cuobjdump -sass fast_slic_b200/libfslic_b200.so, k_assign5<128,3,true,4>: TMA / mbarrier / REDUX / IMMA instructions";
  cuobjdump -sass fast_slic_b200/libfslic_b200.so | awk '/Function : _Z9k_assign5ILi128ELi3ELb1ELi4E/{f=1} f&&/Function :/&&!/k_assign5ILi128ELi3ELb1ELi4E/{f=0} f' | grep -v "^\s*/\* 0x" > /tmp/a5_final.sass;
  grep -n "UTMALDG\|UTMASTG\|UTMACMDFLUSH\|SYNCS\|REDUX\|IMMA\|REDG" /tmp/a5_final.sass | cut -c1-120;
  echo; echo "# the distance loop (4 candidates per trip: LDS.U16 -> VABSDIFF4.U8.ACC -> IMAD -> VIMNMX3)";
  L=$(grep -n "VIMNMX3" /tmp/a5_final.sass | head -1 | cut -d: -f1); sed -n "$((L-14)),$((L+62))p" /tmp/a5_final.sass | cut -c1-110; } > $P/r02_assign_sass.txt
for f in bench_n1 bench_D32 bench_reference; do [ -f $O/${T}_$f.json ] && cp $O/${T}_$f.json $P/r02_$f.json; done
for f in r02_bench_n2 r02_bench_D32_n2 r02_bench_n4 r02_bench_n8 r02_bench_D32_n8 r02_pcie_n2 r02_pcie_n4 r02_pcie_n8 r02_topo_n8; do
  [ -f $O/$f.json ] && cp $O/$f.json $P/$f.json; [ -f $O/$f.txt ] && cp $O/$f.txt $P/$f.txt; done
python - <<PY
import csv, json, subprocess
import os
p = "$P/assign_traffic.json"
d = json.load(open(p))
for key, tag in (("B_batch32", "b32"), ("D_batch8", "D8")):
    rep = "$O/${T}_assign5_%s.ncu-rep" % tag
    if not os.path.exists(rep):
        continue
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h, u, r = rows[0], rows[1], rows[2]
    def val(k):
        v = float(r[h.index(k)]); unit = u[h.index(k)]
        return v * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}.get(unit, 1.0)
    rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
    d[key] = {"kernel": r[h.index("Kernel Name")].split("(")[0], "dram_bytes_read": rd, "dram_bytes_write": wr, "traffic": rd + wr,
              "source": "ncu --set full --clock-control none, one launch (summary in profiles/r02_assign5_%s_ncu_summary.txt)" % tag}
    print("traffic", key, rd + wr)
json.dump(d, open(p, "w"), indent=1)
PY
ls -la $P | tail -30
