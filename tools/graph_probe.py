"""Single-image host calls with and without FSLIC_GRAPH=1: blocking latency and 8 requests in flight."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import synth_images_torch
from fast_slic_b200 import Engine, CLUSTER_DTYPE

H, W, K = 720, 1280, 1600
MP = H * W / 1e6
imgs = synth_images_torch(16, H, W, 77, 12.0, torch.device("cuda", 0)).cpu().pin_memory().numpy()
n_req = 8
engs = [Engine(H, W, K, 1) for _ in range(n_req)]
p = engs[0].params(10.0, 0.0, 3, True, 10)
pr = engs[0].initialize_clusters_host(imgs[:1]).view(np.uint8).reshape(1, K, 32).copy()
cls = [torch.empty((1, K, 32), dtype=torch.uint8).pin_memory().numpy() for _ in range(n_req)]
labs = [torch.empty((1, H, W), dtype=torch.int16).pin_memory().numpy() for _ in range(n_req)]
def blocking(n):
    for i in range(n):
        cls[0][...] = pr
        engs[0].iterate_host(imgs[i % 16][None], cls[0].view(CLUSTER_DTYPE).reshape(1, K), p, labs[0])
def streamed(n):
    for i in range(n):
        s = i % n_req
        engs[s].wait(); cls[s][...] = pr
        engs[s].iterate_host_async(imgs[i % 16][None], cls[s].view(CLUSTER_DTYPE).reshape(1, K), p, labs[s])
    for e in engs:
        e.wait()
blocking(20); t0 = time.perf_counter(); blocking(300); dt = time.perf_counter() - t0
print("FSLIC_GRAPH=%s blocking single image: %.3f ms  %.0f MP/s" % (os.environ.get("FSLIC_GRAPH", "0"), 1e3 * dt / 300, 300 * MP / dt))
streamed(32); t0 = time.perf_counter(); streamed(800); dt = time.perf_counter() - t0
print("FSLIC_GRAPH=%s 8 single-image requests in flight: %.3f ms/image  %.0f MP/s" % (os.environ.get("FSLIC_GRAPH", "0"), 1e3 * dt / 800, 800 * MP / dt))
