"""Regenerates tests/golden/golden_v1.npz from the UNMODIFIED reference compiled by oracle/Makefile
(oracle/_ref/libfslic_ref.so).  Run in the container that has /root/reference:

    python tests/golden/make_golden.py

Inputs are seeded synthetic images (tests/cases.py), so only the reference's OUTPUTS are stored:
final labels (u16), the Cluster[K] records, and SHA-256 of the Lab quad image and of the pre-CCA labels.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from cases import PIPELINE_CASES, make_image, split_kwargs  # noqa: E402
from oracle.oracle import Ref  # noqa: E402

GOLDEN_CASES = [c for c in PIPELINE_CASES if c[2] * c[3] <= 80000]


def main():
    ref = Ref()
    out = {}
    for name, kind, H, W, K, kw in GOLDEN_CASES:
        sigma, a = split_kwargs(kw)
        img = make_image(kind, H, W, seed=7, sigma=sigma)
        cl = ref.initialize(img, K)
        out[name + "/init"] = cl.copy().view(np.uint8)
        lab, quad, pre = ref.iterate(img, cl, a["max_iter"], a["compactness"], a["min_size_factor"],
                                     a["subsample_stride"], a["convert_to_lab"], stages=True, arch="x64/avx2",
                                     num_threads=2)
        lab_std = ref.iterate(img, ref.initialize(img, K), a["max_iter"], a["compactness"], a["min_size_factor"],
                              a["subsample_stride"], a["convert_to_lab"], arch="standard", num_threads=1)
        assert (lab == lab_std).all(), name  # the reference's own scalar and AVX2 paths agree
        out[name + "/labels"] = lab
        out[name + "/clusters"] = cl.view(np.uint8)
        out[name + "/quad_sha"] = np.frombuffer(hashlib.sha256(quad.tobytes()).digest(), np.uint8)
        out[name + "/pre_sha"] = np.frombuffer(hashlib.sha256(pre.tobytes()).digest(), np.uint8)
    # connectivity enforcement alone: the 5x5 input of src/cpptest/test_cca.cpp:178-204 and random maps
    x = 9
    lab5 = np.array([[0, 0, 0, 0, 0], [1, 1, x, 0, 0], [1, x, 0, x, 4], [2, 2, x, x, 4], [2, 3, 3, 3, 3]], np.uint16)
    out["cca5/in"] = lab5
    out["cca5/out"] = ref.enforce_connectivity(lab5, 10, 0, 1)
    rng = np.random.RandomState(5)
    for t, (H, W, nlab, thres) in enumerate([(60, 80, 6, 0), (60, 80, 6, 5), (100, 33, 3, 12), (64, 64, 2, 1)]):
        small = rng.randint(0, nlab, (H // 3 + 1, W // 3 + 1))
        lab = np.kron(small, np.ones((3, 3), int))[:H, :W]
        noise = rng.rand(H, W) < 0.15
        lab[noise] = rng.randint(0, nlab, noise.sum())
        lab = np.ascontiguousarray(lab.astype(np.uint16))
        out["cca_rand%d/in" % t] = lab
        out["cca_rand%d/thres" % t] = np.array([thres, int(lab.max()) + 1])
        out["cca_rand%d/out" % t] = ref.enforce_connectivity(lab, int(lab.max()) + 1, thres, 1)
    np.savez_compressed(os.path.join(HERE, "golden_v1.npz"), **out)
    print("wrote", len(out), "arrays,", os.path.getsize(os.path.join(HERE, "golden_v1.npz")), "bytes")


if __name__ == "__main__":
    main()
