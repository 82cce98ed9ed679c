"""fedem_solvers_b200 -- B200-native (sm_100a) stress recovery for FEDEM superelements.

A from-scratch implementation of the ``fedem_stress`` / ``fedem_gage`` hot path of
SAP-archive/fedem-solvers: expansion of the reduced solution history through the reducer's
B / eigenvector matrices (FP64 tensor-core GEMM), per-element stress kernels with fused
von Mises envelopes, rainflow counting and damage.  All arithmetic runs in hand-written CUDA
kernels behind the C ABI of ``include/fedem_b200.h`` (``lib/libfedem_b200.so``); this package is
the host-side mirror of the reference's driver interface.  There is no CPU fallback.
"""
from ._lib import load_library, library_path, FsrError
from .model import SamData, ElementData, PartModel
from .recovery import StressRecovery, GroupRecovery, Comm, FatigueCounter, fatigue, split_elements
from .gage import Rosette, StrainGages

__all__ = ["load_library", "library_path", "FsrError", "SamData", "ElementData", "PartModel",
           "StressRecovery", "GroupRecovery", "Comm", "split_elements", "FatigueCounter", "fatigue", "Rosette", "StrainGages"]
