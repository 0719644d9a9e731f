#!/bin/bash
# Round-2 GPU call 2: first contact of the TMA-staged assign kernel.
O=gpurun_out
mkdir -p $O
timeout 180 python __graft_entry__.py smoke > $O/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $O/r02_smoke.log
timeout 300 python -m pytest tests/test_parity_gpu.py -q -m gpu -k "tma_kernel or ldg_kernel" > $O/r02_pytest_tma.log 2>&1
echo "pytest tma rc=$?"; tail -15 $O/r02_pytest_tma.log
timeout 900 python -m pytest tests -q -m gpu > $O/r02_pytest2.log 2>&1
echo "pytest all rc=$?"; tail -8 $O/r02_pytest2.log
timeout 300 compute-sanitizer --tool memcheck python tools/sanitize_case.py > $O/r02_sanitizer_memcheck.txt 2>&1; echo "memcheck rc=$?"; tail -3 $O/r02_sanitizer_memcheck.txt
for spec in "B 32 b32" "D 8 D8"; do
  set -- $spec
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02_launches_$3.csv \
    python tools/run_workload.py --workload $1 --batch $2 --iters 1 --timing 0 > /dev/null 2>&1
done
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_assign5 -s 2 -c 1 -f -o $O/r02_assign5_b32 \
  python tools/run_workload.py --workload B --batch 32 --iters 1 --timing 0 > $O/ncu_a5_b32.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_assign5 -s 2 -c 1 -f -o $O/r02_assign5_D8 \
  python tools/run_workload.py --workload D --batch 8 --iters 1 --timing 0 > $O/ncu_a5_D8.log 2>&1
timeout 500 python bench.py --steps 40 --warmup 5 > $O/r02b_bench_n1.json 2> $O/bench_n1.err
echo "bench rc=$?"; tail -c 2500 $O/r02b_bench_n1.json; tail -5 $O/bench_n1.err
timeout 300 python bench.py --workload D --batch 32 --steps 20 --warmup 3 --no-cpu-baseline --extra-batched 0 > $O/r02b_bench_D32.json 2> $O/bench_D32.err
echo "bench D32 rc=$?"; tail -c 1200 $O/r02b_bench_D32.json
