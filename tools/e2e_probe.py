import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import COMPACTNESS, MAX_ITER, STRIDE, WORKLOADS, synth_images_torch
from fast_slic_b200 import get_engine, CLUSTER_DTYPE
H, W, K, msf = WORKLOADS["B"]
for B in (1, 32):
    dev = torch.device("cuda", 0)
    imgs = synth_images_torch(B, H, W, 77, 12.0, dev)
    eng = get_engine(H, W, K, 32, 0)
    pr = eng.initialize_clusters(imgs)
    p = eng.params(COMPACTNESS, msf, STRIDE, True, MAX_ITER)
    h_img = torch.empty((B, H, W, 3), dtype=torch.uint8).pin_memory(); h_img.copy_(imgs)
    h_cl0 = torch.empty(pr.shape, dtype=torch.uint8).pin_memory(); h_cl0.copy_(pr)
    h_cl = torch.empty(pr.shape, dtype=torch.uint8).pin_memory()
    h_lab = torch.empty((B, H, W), dtype=torch.int16).pin_memory()
    d_img = torch.empty_like(imgs); d_cl = torch.empty_like(pr); d_lab = torch.empty((B, H, W), dtype=torch.int16, device=dev)
    def a():
        d_img.copy_(h_img, non_blocking=True); d_cl.copy_(h_cl0, non_blocking=True)
        eng.iterate(d_img, d_cl, p, d_lab)
        h_lab.copy_(d_lab, non_blocking=True); h_cl.copy_(d_cl, non_blocking=True)
        torch.cuda.synchronize()
    img_np, lab_np = h_img.numpy(), h_lab.numpy()
    cl0_np = h_cl0.numpy().view(CLUSTER_DTYPE).reshape(B, K); cl_np = h_cl.numpy().view(CLUSTER_DTYPE).reshape(B, K)
    def b():
        cl_np[...] = cl0_np
        eng.iterate_host(img_np, cl_np, p, lab_np)
    pg_img = np.ascontiguousarray(img_np.copy()); pg_lab = np.empty_like(lab_np); pg_cl = cl0_np.copy()
    def c():
        pg_cl[...] = cl0_np
        eng.iterate_host(pg_img, pg_cl, p, pg_lab)
    for name, f in (("torch copies + iterate", a), ("iterate_host pinned", b), ("iterate_host pageable", c)):
        for _ in range(3): f()
        t0 = time.perf_counter()
        n = 20
        for _ in range(n): f()
        print("B=%d %-26s %.3f ms/step" % (B, name, 1e3 * (time.perf_counter() - t0) / n))
