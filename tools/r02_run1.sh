#!/bin/bash
# Round-2 GPU call 1: the new parity tests, ncu --set full of every non-assign stage kernel, a default bench line.
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/r02_smi.txt 2>&1
timeout 900 python -m pytest tests -x -q -m gpu > $O/r02_pytest1.log 2>&1
echo "pytest rc=$?"; tail -5 $O/r02_pytest1.log
# every stage kernel except the assign kernel (one iterate, batch 32 of 720p): Lab, k_prepare, the whole CCA chain
timeout 600 ncu --set full --import-source on --clock-control none \
  -k regex:'k_rgb_to_quad|k_prepare|k_ccl_|k_cca_|k_kept_|k_scan_blocks' -c 24 -f -o $O/r02_stages_b32 \
  python tools/run_workload.py --workload B --batch 32 --iters 1 --timing 0 > $O/ncu_stages.log 2>&1
echo "ncu stages rc=$?"
timeout 300 ncu --set full --clock-control none -k regex:'k_rgb_to_quad|k_ccl_|k_cca_|k_kept_' -c 16 -f -o $O/r02_stages_D8 \
  python tools/run_workload.py --workload D --batch 8 --iters 1 --timing 0 > $O/ncu_stages_D8.log 2>&1
echo "ncu stages D8 rc=$?"
timeout 400 python bench.py --impl reference --steps 6 --warmup 2 > $O/r02a_bench_reference.json 2> $O/bench_ref.err
echo "ref rc=$?"; tail -c 400 $O/r02a_bench_reference.json
timeout 500 python bench.py --steps 40 --warmup 5 > $O/r02a_bench_n1.json 2> $O/bench_n1.err
echo "bench rc=$?"; tail -c 1500 $O/r02a_bench_n1.json; tail -5 $O/bench_n1.err
