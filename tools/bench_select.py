"""Times the libstdc++ partial_sort replay kernel alone on a synthetic area stream shaped like the 720p / K=1600
workload (170k components: specks of area 1..10 plus ~1600 superpixel bodies), and checks it against std::partial_sort."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from fast_slic_b200 import get_engine
from oracle.oracle import Port
rng = np.random.RandomState(3)
n, K = 170000, 1600
area = np.minimum(rng.geometric(0.55, n), 40).astype(np.int32)
big = rng.choice(n, 1650, replace=False)
area[big] = rng.randint(60, 900, len(big))
eng = get_engine(64, 64, 8, 1, 0)
a = torch.from_numpy(area).cuda()
for _ in range(3):
    kept = eng.debug_heap_select(a, K)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    kept = eng.debug_heap_select(a, K)
e1.record(); torch.cuda.synchronize()
print("select kernel, realistic stream: %.3f ms" % (e0.elapsed_time(e1) / 10))
flat = torch.ones(n, dtype=torch.int32, device="cuda")
for _ in range(3):
    eng.debug_heap_select(flat, K)
torch.cuda.synchronize()
e0.record()
for _ in range(10):
    eng.debug_heap_select(flat, K)
e1.record(); torch.cuda.synchronize()
print("select kernel, no entrants after the fill (streaming/filter overhead only): %.3f ms" % (e0.elapsed_time(e1) / 10))
want = Port().stl_partial_sort(area, K)
print("matches std::partial_sort:", bool((np.nonzero(kept.cpu().numpy())[0] == want).all()))
