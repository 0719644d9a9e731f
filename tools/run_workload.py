"""Small driver for profiling: runs `iters` iterate() calls of one workload (no torch input-generation noise
inside the profiled range beyond the first image synthesis)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import COMPACTNESS, MAX_ITER, STRIDE, WORKLOADS, synth_images_torch  # noqa: E402
from fast_slic_b200 import get_engine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="B")
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--iters", type=int, default=2)
ap.add_argument("--timing", type=int, default=1)
a = ap.parse_args()
H, W, K, msf = WORKLOADS[a.workload]
dev = torch.device("cuda", 0)
imgs = synth_images_torch(a.batch, H, W, 1234, 12.0, dev)
eng = get_engine(H, W, K, a.batch, 0)
pr = eng.initialize_clusters(imgs)
p = eng.params(COMPACTNESS, msf, STRIDE, True, MAX_ITER, collect_timing=a.timing)
for i in range(a.iters):
    cl = pr.clone()
    eng.iterate(imgs, cl, p)
    torch.cuda.synchronize()
    if a.timing:
        print(eng.stage_ms(), eng.assign_kernel_time())
for b in range(min(a.batch, 32) if os.environ.get("FSLIC_COUNTERS") else 0):
    print(b, eng.cca_counters(b))
