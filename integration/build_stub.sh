#!/bin/bash
# Builds integration/cfast_slic_b200*.so against include/fslic_b200.h and fast_slic_b200/libfslic_b200.so
# (Cython -> C++ -> shared object; the same three steps the reference's setup.py performs for cfast_slic.pyx).
set -e
cd "$(dirname "$0")"
PY=${PYTHON:-python}
OUT=${1:-.}
INC=$($PY -c "import sysconfig; print(sysconfig.get_paths()['include'])")
NPI=$($PY -c "import numpy; print(numpy.get_include())")
EXT=$($PY -c "import sysconfig; print(sysconfig.get_config_var('EXT_SUFFIX'))")
$PY -m cython --cplus -3 cfast_slic_b200.pyx -o $OUT/cfast_slic_b200.cpp
g++ -std=c++14 -O2 -fPIC -shared -w -DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION -I$INC -I$NPI -I../include \
    $OUT/cfast_slic_b200.cpp -o $OUT/cfast_slic_b200$EXT -L../fast_slic_b200 -lfslic_b200 -Wl,-rpath,"$(cd ../fast_slic_b200 && pwd)"
echo "built $OUT/cfast_slic_b200$EXT"
