#!/bin/bash
O=gpurun_out
mkdir -p $O
timeout 900 compute-sanitizer --tool synccheck python tools/sanitize_case.py > $O/r02_sanitizer_synccheck.txt 2>&1
echo "synccheck rc=$?"; tail -3 $O/r02_sanitizer_synccheck.txt
timeout 900 compute-sanitizer --tool initcheck python tools/sanitize_case.py > $O/r02_sanitizer_initcheck.txt 2>&1
echo "initcheck rc=$?"; tail -3 $O/r02_sanitizer_initcheck.txt
grep -c "Uninitialized" $O/r02_sanitizer_initcheck.txt
