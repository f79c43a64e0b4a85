"""ima2p_b200 -- B200-native engine for IMa2p's data-parallel hot path.

Only what the path needs: ``csrc/`` (hand-written sm_100a CUDA kernels + the C ABI of
include/ima2p_b200.h) and the host-side mirror of the reference's function seam (``engine``).
Importing the package does not load the CUDA library; the first Engine / LMode does, and fails loudly
when it is missing or no device is present (there is no CPU fallback).
"""
from .engine import (Engine, LMode, HEAT_LINEAR, HEAT_GEOMETRIC, HEAT_EVEN, MODEL_IS, MODEL_HKY, MODEL_SW)  # noqa: F401
from .capi import Ima2pError  # noqa: F401

__all__ = ["Engine", "LMode", "Ima2pError", "HEAT_LINEAR", "HEAT_GEOMETRIC", "HEAT_EVEN", "MODEL_IS", "MODEL_HKY",
           "MODEL_SW"]
