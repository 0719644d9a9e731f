#!/bin/bash
# Round-end evidence capture (run on the GPU box via gpurun): tests, ncu launch lists, ncu --set full of the
# dominant kernel, bench lines.  Everything lands in gpurun_out/ and is summarised into profiles/ afterwards.
R=${1:-r01}
O=gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
for spec in "B 32 b32" "B 1 b1" "D 8 D8"; do
  set -- $spec
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${R}_launches_$3.csv \
    python tools/run_workload.py --workload $1 --batch $2 --iters 1 --timing 0 > /dev/null 2>&1
done
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_assign_warp -s 2 -c 1 -f -o $O/${R}_assign_b32 \
  python tools/run_workload.py --workload B --batch 32 --iters 1 --timing 0 > $O/ncu_b32.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_assign_warp -s 2 -c 1 -f -o $O/${R}_assign_D8 \
  python tools/run_workload.py --workload D --batch 8 --iters 1 --timing 0 > $O/ncu_D8.log 2>&1
timeout 400 python bench.py --impl reference --steps 6 --warmup 3 > $O/${R}_bench_reference.json 2> $O/bench_ref.err
timeout 400 python bench.py > $O/${R}_bench_n1.json 2> $O/bench_n1.err
timeout 300 python bench.py --workload D --batch 8 --steps 40 --warmup 5 --no-cpu-baseline --extra-batched 0 > $O/${R}_bench_D8.json 2> $O/bench_D8.err
timeout 300 python bench.py --workload C --batch 32 --steps 60 --warmup 5 --no-cpu-baseline --extra-batched 0 > $O/${R}_bench_C32.json 2> $O/bench_C32.err
tail -c 600 $O/${R}_bench_n1.json; echo
