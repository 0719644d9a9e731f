#!/bin/bash
# quick check of a build: GPU tests, default bench line, single-image probe
O=gpurun_out
TAG=${1:-r02j}
mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu > $O/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 $O/${TAG}_pytest.log
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --extra-batched 0 > $O/${TAG}_bench_n1.json 2> $O/bench_n1.err
echo "bench rc=$?"; python - <<PY
import json
d=json.load(open("$O/${TAG}_bench_n1.json"))
print("value %.0f MP/s  %.3f ms/step | seq %.3f ms | e2e %.0f (blocking %.0f) | assign %.1f us frac %.3f | parity %s" % (d["value"], d["ms_per_step"], d["sequential"]["ms_per_step"], d["e2e"]["value"], d["e2e"]["blocking"]["value"], d["roofline"]["avg_launch_us"], d["roofline"]["frac"], d.get("parity_checked")))
print(d["roofline"]["stage_ms_last_step"])
print(d["roofline"].get("cca_stage_ms_last_step"))
PY
tail -3 $O/bench_n1.err
python tools/single_probe.py --one
