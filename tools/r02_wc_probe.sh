#!/bin/bash
N=${1:-4}
O=gpurun_out
for mode in "" "--wc"; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/pcie_probe_multi.py $mode > $O/r02_pcie_wc_n$N$mode.txt 2>&1
tail -$((N+2)) $O/r02_pcie_wc_n$N$mode.txt
done
