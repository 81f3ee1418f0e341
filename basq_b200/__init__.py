"""basq_b200 - B200-native kernel recombination (RCHQ) hot path of BASQ.

Public surface mirrors the reference's ``_rchq.py``:
    from basq_b200 import recombination
plus the kernel descriptors (``KernelSpec``), device-side GP helpers (``gp``) and the sharded
driver (``sharded``).  Everything computes inside libbasq_b200.so (hand-written CUDA for sm_100a);
importing fails if the library has not been built and every call fails without a CUDA device.
"""
from . import _lib  # noqa: F401  (fails loudly when the CUDA library is missing)
from ._rchq import (Mod_Tchernychova_Lyons, Tchernychova_Lyons_CAR, ker_svd_sparsify, rc_kernel_svd,
                    recombination)
from .kernels import KernelSpec, describe_kernel

__all__ = ["recombination", "rc_kernel_svd", "ker_svd_sparsify", "Mod_Tchernychova_Lyons",
           "Tchernychova_Lyons_CAR", "KernelSpec", "describe_kernel"]
