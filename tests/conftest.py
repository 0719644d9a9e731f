import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def port():
    from oracle.oracle import Port
    return Port()


@pytest.fixture(scope="session")
def ref():
    from oracle.oracle import Ref
    if not Ref.available():
        pytest.skip("compiled reference (oracle/_ref) not available")
    return Ref()


def _usable_cores():
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = min(n, max(1, int(int(quota) / int(period))))
    except Exception:
        pass
    return max(1, n)


class Checker:
    """What the GPU parity tests compare against: the compiled, unmodified reference (oracle/_ref, SlicAvx2 path)
    wherever its prebuilt library exists -- it travels to the GPU box with the snapshot -- else the plain-C
    restatement.  Same call signature either way."""

    def __init__(self):
        from oracle.oracle import Port, Ref
        self.kind = "reference" if Ref.available() else "port"
        self._impl = Ref() if self.kind == "reference" else Port()
        # the GPU boxes show 128 logical CPUs under a 16-CPU quota: OpenMP must not be left at "all cores" there
        self._threads = min(8, _usable_cores())

    def initialize(self, image, K):
        return self._impl.initialize(image, K)

    def iterate(self, image, clusters, max_iter=10, compactness=10.0, min_size_factor=0.25, stride=3,
                convert_to_lab=True, stages=False, preemptive=False, preemptive_thres=0.05):
        if self.kind == "reference":
            return self._impl.iterate(image, clusters, max_iter, compactness, min_size_factor, stride, convert_to_lab,
                                      stages=stages, arch="x64/avx2", num_threads=self._threads, preemptive=preemptive,
                                      preemptive_thres=preemptive_thres)
        return self._impl.iterate(image, clusters, max_iter, compactness, min_size_factor, stride, convert_to_lab,
                                  stages=stages, preemptive=preemptive, preemptive_thres=preemptive_thres)

    def iterate_real(self, variant, image, clusters, max_iter=10, compactness=10.0, min_size_factor=0.25, stride=3,
                     convert_to_lab=True, stages=False):
        return self._impl.iterate_real(variant, image, clusters, max_iter, compactness, min_size_factor, stride,
                                       convert_to_lab, stages=stages)

    def get_connectivity(self, labels, K):
        return self._impl.get_connectivity(labels, K)

    def get_mask_density(self, clusters, labels, mask):
        return self._impl.get_mask_density(clusters, labels, mask)

    def density_to_mask(self, K, labels, densities):
        return self._impl.density_to_mask(K, labels, densities)

    def enforce_connectivity(self, labels, K, thres):
        if self.kind == "reference":
            return self._impl.enforce_connectivity(labels, K, thres, num_threads=self._threads)
        return self._impl.enforce_connectivity(labels, K, thres)


@pytest.fixture(scope="session")
def checker():
    return Checker()
