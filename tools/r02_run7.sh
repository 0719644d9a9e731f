#!/bin/bash
# GPU tests only (tail of the log) + optional -k filter
O=gpurun_out
TAG=${1:-r02p}
mkdir -p $O
timeout 1200 python -m pytest tests -q -m gpu ${2:+-k "$2"} > $O/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -25 $O/${TAG}_pytest.log
