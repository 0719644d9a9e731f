"""Where the std::partial_sort replay (k_cca_select) spends its time: per-phase clock counts for the images of the
default bench batch that need it.  FSLIC_SELPROF=1 is set here before the context is created."""
import os
import sys

os.environ["FSLIC_SELPROF"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import COMPACTNESS, MAX_ITER, STRIDE, WORKLOADS, synth_images_torch
from fast_slic_b200 import Engine

H, W, K, msf = WORKLOADS["B"]
B = 32
eng = Engine(H, W, K, B)
dev = torch.device("cuda", 0)
imgs = synth_images_torch(B, H, W, 77, 12.0, dev)
cl = eng.initialize_clusters(imgs)
p = eng.params(COMPACTNESS, msf, STRIDE, True, MAX_ITER)
lab = torch.empty((B, H, W), dtype=torch.int16, device=dev)
c0 = cl.clone()
for rep in range(3):
    cl.copy_(c0)
    eng.iterate(imgs, cl, p, lab)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ms = []
for rep in range(10):
    cl.copy_(c0)
    e0.record()
    eng.iterate(imgs, cl, p, lab)
    e1.record()
    torch.cuda.synchronize()
    ms.append(e0.elapsed_time(e1))
print("sequential iterate, batch %d: median %.3f ms" % (B, sorted(ms)[5]))
mhz = 1965.0
for b in range(B):
    c = eng.cca_counters(b)
    if not c["need_sim"]:
        continue
    pr = eng.select_profile(b)
    us = lambda v: v / mhz
    print("image %2d: ncomp %d queued %d chunks %d | replacements %d, loop trips %d | total %.0f us = filter %.0f + build %.0f + replay %.0f"
          " | %.0f clocks per trip" % (b, pr["ncomp"], pr["queued"], pr["chunks"], c["heap_ops"], pr["trips"], us(pr["total"]),
                                       us(pr["filter"]), us(pr["build"]), us(pr["replay"]), pr["replay"] / max(pr["trips"], 1)))
