"""srcfd -- B200 (sm_100a) implementation of the fine-grid Navier-Stokes hot path of
bitseal02/SR-for-CFD behind the reference's own Python entry points.

    from srcfd import ldc   # mirrors PyCFD_ML_accelerated.py   (lid-driven cavity)
    from srcfd import bfs   # mirrors bfs_ml_accelerated.py     (backward-facing step)

Both expose CFDSolver, the configuration classes, the module-level kernel functions and the
workflow functions with the reference's names and signatures.  All field arithmetic runs in
hand-written CUDA kernels through the C ABI in include/srcfd.h; there is no CPU fallback.
"""
from . import _capi
from ._capi import SrcfdError, device_count

__all__ = ["_capi", "SrcfdError", "device_count"]
