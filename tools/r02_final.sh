#!/bin/bash
# Final round-2 evidence on one GPU: tests, launch lists (32 images / one image), ncu --set full of the assign kernel
# and of one launch of every other stage kernel, default + D32 bench lines, single-image probe.
O=gpurun_out
TAG=${1:-r02z}
mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu > $O/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 $O/${TAG}_pytest.log
for spec in "B 32 b32" "B 1 b1"; do
  set -- $spec
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${TAG}_launches_$3.csv \
    python tools/run_workload.py --workload $1 --batch $2 --iters 1 --timing 0 > /dev/null 2>&1
done
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_assign5 -s 2 -c 1 -f -o $O/${TAG}_assign5_b32 \
  python tools/run_workload.py --workload B --batch 32 --iters 1 --timing 0 > $O/ncu_a5_b32.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_assign5 -s 2 -c 1 -f -o $O/${TAG}_assign5_D8 \
  python tools/run_workload.py --workload D --batch 8 --iters 1 --timing 0 > $O/ncu_a5_D8.log 2>&1
timeout 400 ncu --set full --clock-control none -k regex:'k_rgb_to|k_prepare|k_ccl_|k_cca_|k_kept_|k_scan_blocks' -c 26 -f -o $O/${TAG}_stages_b32 \
  python tools/run_workload.py --workload B --batch 32 --iters 1 --timing 0 > $O/ncu_stages.log 2>&1
timeout 400 ncu --set full --clock-control none -k regex:'k_rgb_to|k_prepare|k_ccl_|k_cca_|k_kept_|k_scan_blocks' -c 26 -f -o $O/${TAG}_stages_D8 \
  python tools/run_workload.py --workload D --batch 8 --iters 1 --timing 0 > $O/ncu_stages_D8.log 2>&1
timeout 500 python bench.py --steps 40 --warmup 5 > $O/${TAG}_bench_n1.json 2> $O/bench_n1.err
echo "bench rc=$?"; python - <<PY
import json
d=json.load(open("$O/${TAG}_bench_n1.json"))
print("value %.0f MP/s  %.3f ms/step | seq %.3f ms | e2e %.0f (blocking %.0f) | assign %.1f us frac %.3f | parity %s | cpu %s" % (d["value"], d["ms_per_step"], d["sequential"]["ms_per_step"], d["e2e"]["value"], d["e2e"]["blocking"]["value"], d["roofline"]["avg_launch_us"], d["roofline"]["frac"], d.get("parity_checked"), d["cpu_baseline"]["value"]))
print(d["roofline"]["stage_ms_last_step"])
PY
tail -3 $O/bench_n1.err
timeout 400 python bench.py --impl reference --steps 6 --warmup 3 > $O/${TAG}_bench_reference.json 2> $O/bench_ref.err
tail -c 700 $O/${TAG}_bench_reference.json; echo
timeout 300 python bench.py --workload D --batch 32 --steps 16 --warmup 3 --no-cpu-baseline --extra-batched 0 > $O/${TAG}_bench_D32.json 2> $O/bench_D32.err
python - <<PY
import json
d=json.load(open("$O/${TAG}_bench_D32.json"))
print("D32 value %.0f MP/s  %.3f ms/step | seq %.3f ms | e2e %.0f | assign %.1f us frac %.3f | parity %s" % (d["value"], d["ms_per_step"], d["sequential"]["ms_per_step"], d["e2e"]["value"], d["roofline"]["avg_launch_us"], d["roofline"]["frac"], d.get("parity_checked")))
PY
python tools/single_probe.py --one
python tools/select_probe.py | tail -5
# summarise on the box (the reports of the stage kernels are too large to travel back) and keep only the assign report
PROFILES_DIR=$O/profiles_${TAG} bash tools/make_profiles.sh ${TAG} ${TAG} > $O/make_profiles.log 2>&1
tail -3 $O/make_profiles.log
rm -f $O/${TAG}_stages_b32.ncu-rep $O/${TAG}_stages_D8.ncu-rep $O/${TAG}_assign5_D8.ncu-rep
du -sh $O
