"""ddrl4nav_b200 -- B200-native (sm_100a) actor-learner hot path for DDRL4NAV.

Layout:  csrc/ (CUDA kernels + C ABI, built into libddrl_b200.so)   kernels.py (tensor-level bindings)
         nn/ agent/ server/ data/  (host-side mirror of the reference's interface for this path)
"""
from ._lib import DDRLError, LIB_PATH  # noqa: F401

__version__ = "0.1.0"
