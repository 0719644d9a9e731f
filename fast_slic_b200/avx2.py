"""Import compatibility with /root/reference/fast_slic/avx2.py:1-14: code written against ``fast_slic.avx2.SlicAvx2`` (the
reference's fastest arch, the one BASELINE quotes) switches to this package by changing the import only.  Same engine as
``fast_slic_b200.Slic``: the results are bit-identical to the reference's AVX2 context."""
from .base_slic import LSC, Slic


class SlicAvx2(Slic):
    """== fast_slic.avx2.SlicAvx2 (avx2.py:10-11), on the CUDA engine."""


class LSCAvx2(LSC):
    """== fast_slic.avx2.LSCAvx2 (avx2.py:13-14): iterate() raises NotImplementedError like LSC."""
