#!/bin/bash
# Builds libfslic_b200.so (sm_100a only) next to the Python package.
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 \
  -Xcompiler -fPIC,-O2 -shared -ccbin /usr/bin/g++ ${FSLIC_NVCC_EXTRA} \
  -o ../libfslic_b200.so capi.cu
echo "built $(cd .. && pwd)/libfslic_b200.so"
