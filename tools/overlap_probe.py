"""How much do independent batches overlap on one GPU?  device-resident on 1..3 streams, host-streaming at depth 1..3."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from fast_slic_b200 import Engine, CLUSTER_DTYPE
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
from oracle import synthetic_image

H, W, K, B = 720, 1280, 1600, int(os.environ.get("B", 32))
STEPS = 60
base = np.stack([synthetic_image(H, W, seed=s, sigma=30.0) for s in range(8)])
pool = torch.from_numpy(base).cuda()
pool = torch.stack([torch.roll(pool[i % 8], shifts=(7 * i, 13 * i), dims=(0, 1)) for i in range(4 * B)]).view(4, B, H, W, 3)
MP = H * W / 1e6
for nctx in (1, 2, 3):
    engs = [Engine(H, W, K, B) for _ in range(nctx)]
    p = engs[0].params(10.0, 0.0, 3, True, 10)
    streams = [torch.cuda.Stream() for _ in range(nctx)]
    pr = engs[0].initialize_clusters(pool[0])
    cls = [pr.clone() for _ in range(nctx)]
    labs = [torch.empty((B, H, W), dtype=torch.int16, device="cuda") for _ in range(nctx)]
    def run(n):
        for i in range(n):
            s = i % nctx
            with torch.cuda.stream(streams[s]):
                cls[s].copy_(pr)
                engs[s].iterate(pool[i % 4], cls[s], p, labs[s])
    run(6); torch.cuda.synchronize()
    t0 = time.perf_counter(); run(STEPS); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("device-resident, %d stream(s): %.3f ms/step  %.0f MP/s" % (nctx, 1e3 * dt / STEPS, B * STEPS * MP / dt), flush=True)
    # host streaming at depth nctx
    himg = torch.empty((4, B, H, W, 3), dtype=torch.uint8).pin_memory(); himg.copy_(pool)
    hpr = torch.empty(pr.shape, dtype=torch.uint8).pin_memory(); hpr.copy_(pr)
    hw = [torch.empty(pr.shape, dtype=torch.uint8).pin_memory().numpy() for _ in range(nctx)]
    hl = [torch.empty((B, H, W), dtype=torch.int16).pin_memory().numpy() for _ in range(nctx)]
    def hrun(n):
        for i in range(n):
            s = i % nctx
            engs[s].wait()
            hw[s][...] = hpr.numpy()
            engs[s].iterate_host_async(himg.numpy()[i % 4], hw[s].view(CLUSTER_DTYPE).reshape(B, K), p, hl[s])
        for e in engs:
            e.wait()
    hrun(6)
    t0 = time.perf_counter(); hrun(STEPS); dt = time.perf_counter() - t0
    print("host streaming, depth %d:       %.3f ms/step  %.0f MP/s" % (nctx, 1e3 * dt / STEPS, B * STEPS * MP / dt), flush=True)
    for e in engs:
        e.close()
