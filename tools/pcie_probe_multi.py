"""Host-link probe with every GPU of the box busy at once (torchrun --nproc-per-node N tools/pcie_probe_multi.py):
each rank, bound to its GPU's NUMA node, moves 90 MB up and 60 MB down per round on two streams -- the traffic of one
720p x 32 host step -- alone first, then all ranks together.  Names the host-side limiter of the e2e scaling."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from bench import bind_to_gpu_numa_node, usable_cores

rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
numa = bind_to_gpu_numa_node(lr)
if world > 1:
    dist.init_process_group("gloo")
n, m = 90 * 1000 * 1000, 60 * 1000 * 1000
WC = "--wc" in sys.argv  # upload source in write-combined pinned memory (cudaHostAllocWriteCombined): no CPU-cache snoops
if WC:
    from cuda.bindings import runtime as rt
    err, wc_ptr = rt.cudaHostAlloc(n, rt.cudaHostAllocWriteCombined)
    assert err == rt.cudaError_t.cudaSuccess, err
d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
h_in = None if WC else torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(m, dtype=torch.uint8).pin_memory(); d_out = torch.empty(m, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def rounds(k):
    for _ in range(k):
        with torch.cuda.stream(s1):
            if WC:
                rt.cudaMemcpyAsync(d_in.data_ptr(), wc_ptr, n, rt.cudaMemcpyKind.cudaMemcpyHostToDevice, s1.cuda_stream)
            else:
                d_in.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()


def timed(k=20):
    rounds(3)
    t0 = time.perf_counter()
    rounds(k)
    return (time.perf_counter() - t0) / k


solo = None
for r in range(world):   # one rank at a time
    if world > 1:
        dist.barrier()
    if r == rank:
        solo = timed()
if world > 1:
    dist.barrier()
together = timed()
line = ("[write-combined upload buffer] " if WC else "") + "rank %d (%s): alone %.3f ms/round (H2D %.1f + D2H %.1f GB/s) | all %d ranks at once %.3f ms/round (H2D %.1f + D2H %.1f GB/s)" % (
    rank, numa, 1e3 * solo, n / solo / 1e9, m / solo / 1e9, world, 1e3 * together, n / together / 1e9, m / together / 1e9)
if world > 1:
    out = [None] * world
    dist.all_gather_object(out, line)
    if rank == 0:
        print("usable cores %d" % usable_cores())
        print("\n".join(out))
        tot = sum(float(l.split("at once ")[1].split(" ms")[0]) for l in out) / world
        print("aggregate while all ranks copy: %.1f GB/s up + %.1f GB/s down" % (world * n / (tot / 1e3) / 1e9, world * m / (tot / 1e3) / 1e9))
    dist.destroy_process_group()
else:
    print(line)
