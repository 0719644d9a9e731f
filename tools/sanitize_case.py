"""Small workloads for compute-sanitizer (memcheck / racecheck): one single image and one batch with replays."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from cases import make_image
from fast_slic_b200 import Slic
s = Slic(num_components=60, min_size_factor=0.0)
img = make_image("syn", 96, 128, seed=1, sigma=30.0)
print(s.iterate(img).max())
imgs = np.stack([make_image("noise" if b % 2 else "syn", 96, 128, seed=b, sigma=30.0) for b in range(5)])
print(s.iterate_batch(imgs).max(), s.iterate_batch(torch.from_numpy(imgs).cuda()).max().item())
