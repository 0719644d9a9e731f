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

# both directions at once (what the streaming host path does)
n = 90 * 1000 * 1000
h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
m = 60 * 1000 * 1000
h_out = torch.empty(m, dtype=torch.uint8).pin_memory(); d_out = torch.empty(m, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def both(k):
    for _ in range(k):
        with torch.cuda.stream(s1):
            d_in.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)
both(2); torch.cuda.synchronize()
t0 = time.perf_counter(); both(10); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("bidirectional 90 MB up + 60 MB down per round: %.3f ms/round  (H2D %.1f GB/s, D2H %.1f GB/s)" % (1e3 * dt / 10, 10 * n / dt / 1e9, 10 * m / dt / 1e9))
