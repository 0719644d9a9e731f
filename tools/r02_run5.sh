#!/bin/bash
O=gpurun_out
TAG=${1:-r02e}
mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu > $O/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -6 $O/${TAG}_pytest.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${TAG}_launches_b32.csv \
    python tools/run_workload.py --workload B --batch 32 --iters 1 --timing 0 > /dev/null 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_assign5 -s 2 -c 1 -f -o $O/${TAG}_assign5_b32 \
  python tools/run_workload.py --workload B --batch 32 --iters 1 --timing 0 > $O/ncu_a5_b32.log 2>&1
timeout 200 ncu --set full --import-source on --clock-control none -k regex:k_rgb_to -c 1 -f -o $O/${TAG}_lab_b32 \
  python tools/run_workload.py --workload B --batch 32 --iters 1 --timing 0 > /dev/null 2>&1
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --extra-batched 0 > $O/${TAG}_bench_n1.json 2> $O/bench_n1.err
echo "bench rc=$?"; python - <<PY
import json
d=json.load(open("$O/${TAG}_bench_n1.json"))
print("value %.0f MP/s  %.3f ms/step | seq %.3f ms | e2e %.0f (blocking %.0f) | assign %.1f us frac %.3f | parity %s" % (d["value"], d["ms_per_step"], d["sequential"]["ms_per_step"], d["e2e"]["value"], d["e2e"]["blocking"]["value"], d["roofline"]["avg_launch_us"], d["roofline"]["frac"], d.get("parity_checked")))
print(d["roofline"]["stage_ms_last_step"])
PY
tail -3 $O/bench_n1.err
timeout 300 python bench.py --workload D --batch 32 --steps 16 --warmup 3 --no-cpu-baseline --extra-batched 0 > $O/${TAG}_bench_D32.json 2> $O/bench_D32.err
python - <<PY
import json
d=json.load(open("$O/${TAG}_bench_D32.json"))
print("D32 value %.0f MP/s  %.3f ms/step | seq %.3f ms | e2e %.0f | assign %.1f us frac %.3f | parity %s" % (d["value"], d["ms_per_step"], d["sequential"]["ms_per_step"], d["e2e"]["value"], d["roofline"]["avg_launch_us"], d["roofline"]["frac"], d.get("parity_checked")))
PY
