"""oracle/oracle.py -- TEST INFRASTRUCTURE ONLY (ctypes doors to the two CPU checkers).

* ``Port``  : oracle/liboracle.so   -- our plain-C restatement (oracle/slic_oracle.c)
* ``Ref``   : oracle/_ref/libfslic_ref.so -- the unmodified reference compiled from /root/reference
              (oracle/Makefile + oracle/ref_shim.cpp); exists wherever it was prebuilt.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product (fast_slic_b200) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

CLUSTER_DTYPE = np.dtype(
    [("y", "<f4"), ("x", "<f4"), ("r", "<f4"), ("g", "<f4"), ("b", "<f4"), ("a", "<f4"),
     ("number", "<u2"), ("is_active", "u1"), ("is_updatable", "u1"), ("num_members", "<u4")]
)
assert CLUSTER_DTYPE.itemsize == 32

_u8p = C.POINTER(C.c_uint8)
_u16p = C.POINTER(C.c_uint16)
_i32p = C.POINTER(C.c_int32)


def _p(arr, typ):
    return None if arr is None else arr.ctypes.data_as(typ)


def build(force=False):
    """Compile liboracle.so (always possible) and _ref (only where /root/reference exists)."""
    if force or not os.path.exists(os.path.join(_HERE, "liboracle.so")):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    if os.path.isdir("/root/reference/src") and (
        force or not os.path.exists(os.path.join(_HERE, "_ref", "libfslic_ref.so"))
    ):
        subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)


def synthetic_image(H, W, seed, sigma=12.0):
    """SURVEY.md section 8(d) synthetic input: smooth sinusoid field + N(0, sigma^2), uint8 RGB."""
    rng = np.random.RandomState(seed)
    y, x = np.mgrid[0:H, 0:W].astype(np.float32)
    base = np.stack([np.sin(x / 37 + y / 91), np.cos(y / 53 - x / 113), np.sin((x + y) / 71)], axis=-1)
    img = 127 + 100 * base + rng.normal(0, sigma, size=(H, W, 3))
    return np.ascontiguousarray(np.clip(img, 0, 255).astype(np.uint8))


class Port:
    """Plain-C restatement."""

    def __init__(self):
        build()
        self.lib = C.CDLL(os.path.join(_HERE, "liboracle.so"))
        assert self.lib.orc_sizeof_cluster() == 32

    def lab_tables(self):
        gamma = np.zeros(256, np.int32)
        lab = np.zeros(8193, np.int32)
        cb = np.zeros(9, np.int32)
        self.lib.orc_lab_tables(_p(gamma, _i32p), _p(lab, _i32p), _p(cb, _i32p))
        return gamma, lab, cb

    def rgb_to_quad(self, image, convert_to_lab=True):
        H, W, _ = image.shape
        quad = np.zeros((H, W, 4), np.uint8)
        self.lib.orc_rgb_to_quad(_p(image, _u8p), H, W, int(convert_to_lab), _p(quad, _u8p))
        return quad

    def initialize(self, image, K):
        H, W, _ = image.shape
        clusters = np.zeros(K, CLUSTER_DTYPE)
        self.lib.orc_initialize_clusters(H, W, K, _p(image, _u8p), clusters.ctypes.data_as(C.c_void_p))
        return clusters

    def spatial_lut(self, S, compactness, color_shift):
        lut = np.zeros(2 * S + 1, np.uint16)
        self.lib.orc_spatial_lut(S, C.c_float(compactness), color_shift, _p(lut, _u16p))
        return lut

    def iterate(self, image, clusters, max_iter=10, compactness=10.0, min_size_factor=0.25, stride=3,
                convert_to_lab=True, stages=False, preemptive=False, preemptive_thres=0.05):
        image = np.ascontiguousarray(image)
        H, W, _ = image.shape
        K = len(clusters)
        out = np.zeros((H, W), np.uint16)
        quad = np.zeros((H, W, 4), np.uint8) if stages else None
        pre = np.zeros((H, W), np.uint16) if stages else None
        self.lib.orc_iterate_preemptive(H, W, K, _p(image, _u8p), clusters.ctypes.data_as(C.c_void_p), _p(out, _u16p),
                                        max_iter, C.c_float(compactness), C.c_float(min_size_factor), stride,
                                        int(convert_to_lab), int(bool(preemptive)), C.c_float(preemptive_thres),
                                        _p(quad, _u8p), _p(pre, _u16p))
        return (out, quad, pre) if stages else out

    def enforce_connectivity(self, labels, K, thres):
        out = np.ascontiguousarray(labels.astype(np.uint16))
        H, W = out.shape
        self.lib.orc_enforce_connectivity(_p(out, _u16p), H, W, K, thres)
        return out

    def heap_select(self, area, middle):
        """Returns the sorted list of indices kept by __heap_select over comps = arange(len(area))."""
        area = np.ascontiguousarray(area, np.int32)
        comps = np.arange(len(area), dtype=np.int32)
        self.lib.orc_heap_select(_p(comps, _i32p), C.c_long(len(area)), C.c_long(middle), _p(area, _i32p))
        return np.sort(comps[:middle])

    def stl_partial_sort(self, area, middle):
        area = np.ascontiguousarray(area, np.int32)
        comps = np.arange(len(area), dtype=np.int32)
        self.lib.stl_partial_sort_by_area(_p(comps, _i32p), C.c_long(len(area)), C.c_long(middle), _p(area, _i32p))
        return np.sort(comps[:middle])

    # ---- consumers of the label map (fast-slic.cpp:16-168) ----
    def get_connectivity(self, labels, K):
        labels = np.ascontiguousarray(labels, np.uint16)
        H, W = labels.shape
        counts = np.zeros(K, np.int32)
        nb = np.zeros((K, 12), np.uint32)
        self.lib.orc_get_connectivity(H, W, K, _p(labels, _u16p), _p(counts, _i32p), nb.ctypes.data_as(C.c_void_p))
        return [nb[k, :counts[k]].tolist() for k in range(K)]


    def get_mask_density(self, clusters, labels, mask):
        labels = np.ascontiguousarray(labels, np.uint16)
        mask = np.ascontiguousarray(mask, np.uint8)
        H, W = labels.shape
        K = len(clusters)
        dens = np.zeros(K, np.uint8)
        self.lib.orc_get_mask_density(H, W, K, clusters.ctypes.data_as(C.c_void_p), _p(labels, _u16p), _p(mask, _u8p),
                                      _p(dens, _u8p))
        return dens

    def density_to_mask(self, K, labels, densities):
        labels = np.ascontiguousarray(labels, np.uint16)
        densities = np.ascontiguousarray(densities, np.uint8)
        H, W = labels.shape
        out = np.zeros((H, W), np.uint8)
        self.lib.orc_cluster_density_to_mask(H, W, K, _p(labels, _u16p), _p(densities, _u8p), _p(out, _u8p))
        return out

    def iterate_real(self, variant, image, clusters, max_iter=10, compactness=10.0, min_size_factor=0.25, stride=3,
                     convert_to_lab=True, stages=False):
        """Float-distance contexts (context.cpp:394-499): variant 0 standard, 1 l2, 2 noq."""
        image = np.ascontiguousarray(image)
        H, W, _ = image.shape
        out = np.zeros((H, W), np.uint16)
        pre = np.zeros((H, W), np.uint16)
        self.lib.orc_iterate_real(int(variant), H, W, len(clusters), _p(image, _u8p), clusters.ctypes.data_as(C.c_void_p),
                                  _p(out, _u16p), max_iter, C.c_float(compactness), C.c_float(min_size_factor), stride,
                                  int(convert_to_lab), _p(pre, _u16p))
        return (out, pre) if stages else out


class Ref:
    """The unmodified reference (standard or x64/avx2 arch), OpenMP threads = num_threads (-1: all)."""

    @staticmethod
    def available():
        return os.path.exists(os.path.join(_HERE, "_ref", "libfslic_ref.so")) or os.path.isdir("/root/reference/src")

    def __init__(self):
        build()
        self.lib = C.CDLL(os.path.join(_HERE, "_ref", "libfslic_ref.so"))
        assert self.lib.ref_sizeof_cluster() == 32

    def initialize(self, image, K):
        H, W, _ = image.shape
        clusters = np.zeros(K, CLUSTER_DTYPE)
        self.lib.ref_initialize(H, W, K, _p(image, _u8p), clusters.ctypes.data_as(C.c_void_p))
        return clusters

    def iterate(self, image, clusters, max_iter=10, compactness=10.0, min_size_factor=0.25, stride=3,
                convert_to_lab=True, stages=False, arch="x64/avx2", num_threads=-1, preemptive=False,
                preemptive_thres=0.05):
        image = np.ascontiguousarray(image)
        H, W, _ = image.shape
        K = len(clusters)
        out = np.zeros((H, W), np.uint16)
        quad = np.zeros((H, W, 4), np.uint8) if stages else None
        pre = np.zeros((H, W), np.uint16) if stages else None
        if preemptive:  # (no quad stage dump on this door; the Lab stage does not depend on the flag)
            self.lib.ref_iterate_preemptive(1 if arch == "x64/avx2" else 0, H, W, K, _p(image, _u8p),
                                            clusters.ctypes.data_as(C.c_void_p), _p(out, _u16p), max_iter,
                                            C.c_float(compactness), C.c_float(min_size_factor), stride,
                                            int(convert_to_lab), C.c_float(preemptive_thres), num_threads, _p(pre, _u16p))
            return (out, quad, pre) if stages else out
        self.lib.ref_iterate(1 if arch == "x64/avx2" else 0, H, W, K, _p(image, _u8p),
                             clusters.ctypes.data_as(C.c_void_p), _p(out, _u16p), max_iter, C.c_float(compactness),
                             C.c_float(min_size_factor), stride, int(convert_to_lab), num_threads,
                             _p(quad, _u8p), _p(pre, _u16p))
        return (out, quad, pre) if stages else out

    def enforce_connectivity(self, labels, K, thres, num_threads=-1):
        out = np.ascontiguousarray(labels.astype(np.uint16))
        H, W = out.shape
        self.lib.ref_enforce_connectivity(_p(out, _u16p), H, W, K, thres, num_threads)
        return out

    # ---- consumers of the label map: the reference's own fast-slic.o behind oracle/ref_shim.cpp ----
    def get_connectivity(self, labels, K):
        labels = np.ascontiguousarray(labels, np.uint16)
        H, W = labels.shape
        counts = np.zeros(K, np.int32)
        nb = np.zeros((K, 12), np.uint32)
        self.lib.ref_get_connectivity(H, W, K, _p(labels, _u16p), _p(counts, _i32p), nb.ctypes.data_as(C.c_void_p))
        return [nb[k, :counts[k]].tolist() for k in range(K)]


    def get_mask_density(self, clusters, labels, mask):
        labels = np.ascontiguousarray(labels, np.uint16)
        mask = np.ascontiguousarray(mask, np.uint8)
        H, W = labels.shape
        K = len(clusters)
        dens = np.zeros(K, np.uint8)
        self.lib.ref_get_mask_density(H, W, K, clusters.ctypes.data_as(C.c_void_p), _p(labels, _u16p), _p(mask, _u8p),
                                      _p(dens, _u8p))
        return dens

    def density_to_mask(self, K, labels, densities):
        labels = np.ascontiguousarray(labels, np.uint16)
        densities = np.ascontiguousarray(densities, np.uint8)
        H, W = labels.shape
        out = np.zeros((H, W), np.uint8)
        clusters = np.zeros(K, CLUSTER_DTYPE)
        self.lib.ref_cluster_density_to_mask(H, W, K, clusters.ctypes.data_as(C.c_void_p), _p(labels, _u16p),
                                             _p(densities, _u8p), _p(out, _u8p))
        return out

    def iterate_real(self, variant, image, clusters, max_iter=10, compactness=10.0, min_size_factor=0.25, stride=3,
                     convert_to_lab=True, stages=False, num_threads=2):
        image = np.ascontiguousarray(image)
        H, W, _ = image.shape
        out = np.zeros((H, W), np.uint16)
        pre = np.zeros((H, W), np.uint16)
        self.lib.ref_iterate_real(int(variant), H, W, len(clusters), _p(image, _u8p), clusters.ctypes.data_as(C.c_void_p),
                                  _p(out, _u16p), max_iter, C.c_float(compactness), C.c_float(min_size_factor), stride,
                                  int(convert_to_lab), num_threads, _p(pre, _u16p))
        return (out, pre) if stages else out
