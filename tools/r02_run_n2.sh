#!/bin/bash
# 2-GPU call: NCCL label gather loop of bench.py, per-rank NUMA binding, host-link probe with both GPUs busy.
O=gpurun_out
N=${1:-2}
mkdir -p $O
nvidia-smi topo -m > $O/r02_topo_n$N.txt 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/pcie_probe_multi.py > $O/r02_pcie_n$N.txt 2>&1
cat $O/r02_pcie_n$N.txt | tail -6
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 30 --warmup 5 --extra-batched 0 > $O/r02_bench_n$N.json 2> $O/bench_n$N.err
echo "bench rc=$?"; tail -3 $O/bench_n$N.err
python - <<PY
import json
d=json.load(open("$O/r02_bench_n$N.json"))
print("N=%d value %.0f MP/s %.3f ms/step | e2e %.0f | gather %s | parity %s" % (d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], {k: d["gather"][k] for k in ("value","ms_per_step","own_shard_intact")} if d.get("gather") else None, d.get("parity_checked")))
print(d["e2e"]["numa"])
PY
if [ "$N" = "2" ]; then
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --workload D --batch 32 --steps 12 --warmup 3 --extra-batched 0 > $O/r02_bench_D32_n$N.json 2> $O/bench_D32_n$N.err
echo "bench D rc=$?"; tail -c 600 $O/r02_bench_D32_n$N.json
fi
