"""lowthrustopt_b200 -- B200-native segment propagation for LowThrustOpt's multiple shooting.

Only the hot path lives here: the C-ABI library (csrc/ -> liblto_b200.so), its ctypes
binding (capi) and the host-side mirror of the reference's defectCalc / jacobianCalc
closures (direct, indirect).  Importing the package does not load CUDA; the first
propagation call does, and raises if the library or the device is missing.
"""
from .capi import MU, DU, TU, DAY  # noqa: F401

__all__ = ["MU", "DU", "TU", "DAY"]
