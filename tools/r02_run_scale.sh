#!/bin/bash
# N-GPU call (N = 4 or 8): host-link probe with all GPUs busy, default bench line under torchrun (gather loop included),
# and for N = 8 the BASELINE configs[3] line (32 4K images per GPU = 256 over the box).
O=gpurun_out
N=${1:-8}
mkdir -p $O
nvidia-smi topo -m > $O/r02_topo_n$N.txt 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/pcie_probe_multi.py > $O/r02_pcie_n$N.txt 2>&1
tail -$((N+3)) $O/r02_pcie_n$N.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 30 --warmup 5 --extra-batched 0 > $O/r02_bench_n$N.json 2> $O/bench_n$N.err
echo "bench rc=$?"; tail -2 $O/bench_n$N.err
python - <<PY
import json
d=json.load(open("$O/r02_bench_n$N.json"))
print("N=%d value %.0f MP/s %.3f ms/step | e2e %.0f (%.3f ms/step) | gather %s | parity %s" % (d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], {k: d["gather"][k] for k in ("value","ms_per_step","own_shard_intact")} if d.get("gather") else None, d.get("parity_checked")))
print(d["e2e"]["numa"])
PY
if [ "$N" = "8" ]; then
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --workload D --batch 32 --steps 10 --warmup 3 --extra-batched 0 > $O/r02_bench_D32_n$N.json 2> $O/bench_D32_n$N.err
echo "bench D rc=$?"
python - <<PY
import json
d=json.load(open("$O/r02_bench_D32_n$N.json"))
print("D32 N=%d value %.0f MP/s %.3f ms/step | e2e %.0f | gather %s | parity %s" % (d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], {k: d["gather"][k] for k in ("value","ms_per_step")} if d.get("gather") else None, d.get("parity_checked")))
PY
fi
