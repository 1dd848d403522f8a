"""sigma_b200 -- B200-native SpMV + Krylov hot path of SiGMA (danshapero/sigma).

The product is the sm_100a CUDA library in ``csrc/`` behind the C-ABI of
``include/sigma_b200.h``; this package is its host-side mirror of the
reference's operator / solver interface plus deterministic input generators.
"""
from .api import *  # noqa: F401,F403
from . import generators  # noqa: F401
