import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import COMPACTNESS, MAX_ITER, STRIDE, WORKLOADS, synth_images_torch
from fast_slic_b200 import get_engine, CLUSTER_DTYPE
H, W, K, msf = WORKLOADS["B"]
B = 32
dev = torch.device("cuda", 0)
imgs = synth_images_torch(B, H, W, 77, 12.0, dev)
eng = get_engine(H, W, K, 32, 0)
pr = eng.initialize_clusters(imgs)
p = eng.params(COMPACTNESS, msf, STRIDE, True, MAX_ITER)
h_img = torch.empty((B, H, W, 3), dtype=torch.uint8).pin_memory(); h_img.copy_(imgs)
h_cl0 = torch.empty(pr.shape, dtype=torch.uint8).pin_memory(); h_cl0.copy_(pr)
h_cl = torch.empty(pr.shape, dtype=torch.uint8).pin_memory()
h_lab = torch.empty((B, H, W), dtype=torch.int16).pin_memory()
img_np, lab_np = h_img.numpy(), h_lab.numpy()
cl0_np = h_cl0.numpy().view(CLUSTER_DTYPE).reshape(B, K); cl_np = h_cl.numpy().view(CLUSTER_DTYPE).reshape(B, K)
for i in range(4):
    cl_np[...] = cl0_np
    if i == 3: os.environ["FSLIC_TRACE"] = "1"
    t0 = time.perf_counter()
    eng.iterate_host(img_np, cl_np, p, lab_np)
    print("step %d %.3f ms" % (i, 1e3 * (time.perf_counter() - t0)))
