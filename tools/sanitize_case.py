"""Small workloads for compute-sanitizer (memcheck / racecheck): a single image (fused prepare tail, graph replay),
batches with std::partial_sort replays (k_cca_select), the float-distance variants, preemptive, graph / density."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from cases import make_image
import fast_slic_b200 as fs
s = fs.Slic(num_components=60, min_size_factor=0.0)
img = make_image("syn", 96, 128, seed=1, sigma=30.0)
print(s.iterate(img).max(), s.iterate(img).max())
imgs = np.stack([make_image("noise" if b % 2 else "syn", 96, 128, seed=b, sigma=30.0) for b in range(5)])
print(s.iterate_batch(imgs).max(), s.iterate_batch(torch.from_numpy(imgs).cuda()).max().item())
eng = fs.base_slic.get_engine(96, 128, 60, 1, 0)
print("replays:", sum(fs.base_slic.get_engine(96, 128, 60, 8, 0).cca_counters(b)["need_sim"] for b in range(5)))
for cls in (fs.SlicRealDist, fs.SlicRealDistL2, fs.SlicRealDistNoQ):
    print(cls.__name__, cls(num_components=40).iterate(img).max())
p = fs.Slic(num_components=60, preemptive=True, preemptive_thres=0.1)
print("preemptive", p.iterate(img).max())
lab = s.iterate(img)
m = s.slic_model
print(len(m.get_connectivity(lab).tolist()), m.get_mask_density((img[..., 0] > 100).astype(np.uint8) * 255, lab).sum())
big = fs.Slic(num_components=400, min_size_factor=0.0)
print(big.iterate(make_image("noise", 240, 320, seed=3)).max())
