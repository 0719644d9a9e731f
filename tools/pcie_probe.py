import time, torch
for mb in (3, 90):
    n = mb * 1000 * 1000
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    for _ in range(10):
        h.copy_(d, non_blocking=True)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("%d MB: H2D %.1f GB/s  D2H %.1f GB/s" % (mb, 10 * n / (t1 - t0) / 1e9, 10 * n / (t2 - t1) / 1e9))
