#!/bin/bash
# bench lines of the final build: default (with the CPU baseline leg), reference arm, 4K x 32
O=gpurun_out
mkdir -p $O
timeout 500 python bench.py > $O/r02f_bench_n1.json 2> $O/bench_n1.err
echo "bench rc=$?"; tail -2 $O/bench_n1.err
timeout 400 python bench.py --impl reference > $O/r02f_bench_reference.json 2> $O/bench_ref.err
echo "ref rc=$?"
timeout 300 python bench.py --workload D --batch 32 --steps 16 --warmup 3 --no-cpu-baseline --extra-batched 0 > $O/r02f_bench_D32.json 2> $O/bench_D32.err
python - <<PY
import json
for f in ("n1", "reference", "D32"):
    d = json.load(open("$O/r02f_bench_%s.json" % f))
    print(f, d.get("value"), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), ((d.get("e2e") or {}).get("blocking") or {}).get("value"),
          (d.get("roofline") or {}).get("frac"), (d.get("cpu_baseline") or {}).get("value"), d.get("parity_checked"))
PY
