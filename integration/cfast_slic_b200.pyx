# cython: language_level=3
# distutils: language = c++
"""The binding a maintainer of Algy/fast-slic would add next to cfast_slic.pyx: the `cuda/sm_100a` branch of
SlicModel.initialize / SlicModel.iterate (cfast_slic.pyx:124-147, 150-260) on top of the C ABI in
include/fslic_b200.h, with the reference's own memoryview signatures, GIL discipline and exception types.

Built and import-tested by tests/test_cpu.py::test_cython_stub_builds_and_binds (integration/build_stub.sh);
the product itself binds the same ABI through ctypes (fast_slic_b200/_lib.py)."""
from libc.stdint cimport uint8_t, uint16_t, uint32_t, int16_t
from libc.stdlib cimport malloc, free
from libc.string cimport memset
import numpy as np
cimport numpy as np

cdef extern from "fslic_b200.h":
    ctypedef struct fslic_cluster:
        float y, x, r, g, b, a
        uint16_t number
        uint8_t is_active, is_updatable
        uint32_t num_members
    ctypedef struct fslic_params:
        float compactness
        float min_size_factor
        int subsample_stride
        int convert_to_lab
        int max_iter
        int collect_timing
    ctypedef struct fslic_ctx:
        pass
    const char* fslic_b200_last_error() nogil
    int fslic_b200_create(int device, int H, int W, int K, int max_batch, fslic_ctx** out) nogil
    int fslic_b200_destroy(fslic_ctx* ctx) nogil
    int fslic_b200_initialize_clusters_host(fslic_ctx* ctx, const uint8_t* h_images, fslic_cluster* h_clusters, int batch) nogil
    int fslic_b200_iterate_host(fslic_ctx* ctx, const uint8_t* h_images, fslic_cluster* h_clusters, uint16_t* h_labels,
                                int batch, const fslic_params* params) nogil
    int fslic_b200_sizeof_cluster() nogil


cdef class SlicModelCuda:
    """== cfast_slic.SlicModel restricted to arch_name == "cuda/sm_100a"."""
    cdef fslic_cluster* _c_clusters
    cdef fslic_ctx* _ctx
    cdef int _ctx_H, _ctx_W
    cdef public int num_components
    cdef public object initialized
    cdef public object convert_to_lab
    cdef public int device

    def __cinit__(self, int num_components, int device=0):
        if num_components >= 65534:
            raise ValueError("num_components cannot exceed 65534")       # cfast_slic.pyx:24-25
        elif num_components <= 0:
            raise ValueError("num_components should be a non-negative integer")  # cfast_slic.pyx:26-27
        self.num_components = num_components
        self._c_clusters = <fslic_cluster*>malloc(sizeof(fslic_cluster) * num_components)   # cfast_slic.pyx:33
        memset(self._c_clusters, 0, sizeof(fslic_cluster) * num_components)
        self._ctx = NULL
        self.initialized = False
        self.convert_to_lab = False
        self.device = device

    cdef _context(self, int H, int W):
        cdef int rc
        cdef fslic_ctx* ctx = NULL
        if self._ctx != NULL and (self._ctx_H != H or self._ctx_W != W):
            fslic_b200_destroy(self._ctx)
            self._ctx = NULL
        if self._ctx == NULL:
            rc = fslic_b200_create(self.device, H, W, self.num_components, 1, &ctx)
            if rc != 0:
                raise RuntimeError(fslic_b200_last_error().decode("utf-8"))
            self._ctx = ctx
            self._ctx_H = H
            self._ctx_W = W

    cpdef void initialize(self, const uint8_t [:, :, ::1] image):
        if image.shape[2] != 3:
            raise ValueError("nchan != 3")                                # cfast_slic.pyx:125
        cdef int rc
        self._context(image.shape[0], image.shape[1])
        with nogil:                                                       # cfast_slic.pyx:143
            rc = fslic_b200_initialize_clusters_host(self._ctx, &image[0, 0, 0], self._c_clusters, 1)
        if rc != 0:
            raise RuntimeError(fslic_b200_last_error().decode("utf-8"))
        self.initialized = True

    cpdef iterate(self, const uint8_t [:, :, ::1] image, int max_iter, float compactness, float min_size_factor,
                  uint8_t subsample_stride):
        if not self.initialized:
            raise RuntimeError("Slic model is not initialized")           # cfast_slic.pyx:151
        if image.shape[2] != 3:
            raise ValueError("nchan != 3")                                # cfast_slic.pyx:153
        cdef int H = image.shape[0], W = image.shape[1], rc
        cdef np.ndarray[np.uint16_t, ndim=2, mode='c'] assignments = np.zeros([H, W], dtype=np.uint16)   # cfast_slic.pyx:160-161
        cdef fslic_params p
        p.compactness = compactness
        p.min_size_factor = min_size_factor
        p.subsample_stride = subsample_stride
        p.convert_to_lab = 1 if self.convert_to_lab else 0
        p.max_iter = max_iter
        p.collect_timing = 0
        self._context(H, W)
        with nogil:                                                       # cfast_slic.pyx:188
            rc = fslic_b200_iterate_host(self._ctx, &image[0, 0, 0], self._c_clusters, <uint16_t*>&assignments[0, 0], 1, &p)
        if rc == -1:
            raise ValueError(fslic_b200_last_error().decode("utf-8"))
        if rc != 0:
            raise RuntimeError(fslic_b200_last_error().decode("utf-8"))
        result = assignments.astype(np.int16)                             # cfast_slic.pyx:258-260
        result[result == 0xFFFF] = -1
        return result

    @property
    def clusters(self):                                                   # cfast_slic.pyx:51-66
        cdef fslic_cluster* c
        cdef int i
        result = []
        for i in range(self.num_components):
            c = self._c_clusters + i
            result.append(dict(number=c.number, yx=(c.y, c.x), color=(c.r, c.g, c.b), num_members=c.num_members))
        return result

    def __dealloc__(self):
        if self._ctx != NULL:
            fslic_b200_destroy(self._ctx)
        if self._c_clusters != NULL:
            free(self._c_clusters)


def sizeof_cluster():
    return fslic_b200_sizeof_cluster()
