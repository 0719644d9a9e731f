#!/bin/bash
# select-loop variants (with / without warp barriers) + compute-sanitizer memcheck / racecheck on the small workload
O=gpurun_out
mkdir -p $O
echo "== FSLIC_SELSYNC=1 (default)"; python tools/select_probe.py 2>&1 | tail -5
echo "== FSLIC_SELSYNC=0"; FSLIC_SELSYNC=0 python tools/select_probe.py 2>&1 | tail -5
timeout 900 compute-sanitizer --tool memcheck --leak-check no python tools/sanitize_case.py > $O/r02_sanitizer_memcheck.txt 2>&1
echo "memcheck rc=$?"; tail -4 $O/r02_sanitizer_memcheck.txt
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all python tools/sanitize_case.py > $O/r02_sanitizer_racecheck.txt 2>&1
echo "racecheck rc=$?"; tail -4 $O/r02_sanitizer_racecheck.txt
FSLIC_SELSYNC=0 timeout 900 compute-sanitizer --tool racecheck --racecheck-report all python tools/sanitize_case.py > $O/r02_sanitizer_racecheck_nosync.txt 2>&1
echo "racecheck (no barriers) rc=$?"; tail -4 $O/r02_sanitizer_racecheck_nosync.txt
