"""Load the UNMODIFIED reference scripts from /root/reference for validation.

TEST INFRASTRUCTURE ONLY.  Works only where the reference tree is mounted (the
build container); the GPU box has no /root/reference, so nothing marked
`gpu`, smoke() or bench.py may call this.  It is used by
tests/test_oracle_vs_reference.py (skipped when the tree is absent) and by
tests/golden/make_golden.py, which turns reference outputs into the committed
fixtures.

The four scripts import matplotlib / h5py / tensorflow at module scope; none of
those is installed here, so empty stand-in modules are registered first.  The
numba kernels themselves run for real.  NUMBA_NUM_THREADS must be set before
numba is first imported: 1 gives the deterministic lexicographic Gauss-Seidel
order the oracle restates.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get("SRCFD_REFERENCE_ROOT", "/root/reference")

FILES = {
    "LDC": "PyCFD_ML_accelerated.py",
    "BFS": "bfs_ml_accelerated.py",
    "LDC_sir": "LDV PyCFD given by sir.py",
    "BFS_sir": "bfs code given by sir.py",
}


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, FILES["LDC"]))


def _stub(name: str, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _install_stubs():
    class _Model:  # base class needed at PyCFD_ML_accelerated.py:676
        def __init__(self, *a, **k):
            pass

    try:
        import matplotlib  # noqa: F401
    except Exception:
        mpl = _stub("matplotlib")
        plt = _stub("matplotlib.pyplot")
        mpl.pyplot = plt
        mpl.use = lambda *a, **k: None
    try:
        import h5py  # noqa: F401
    except Exception:
        _stub("h5py")
    try:
        import tensorflow  # noqa: F401
    except Exception:
        tf = _stub("tensorflow")
        keras = _stub("tensorflow.keras", Model=_Model)
        tf.keras = keras
        _stub("tensorflow.keras.models")


_cache: dict = {}


def load(which: str, threads: int | None = 1):
    """Return the reference module `which` in FILES.  `threads` sets NUMBA_NUM_THREADS
    if numba has not been imported yet (it cannot be raised afterwards)."""
    if which in _cache:
        return _cache[which]
    if not available():
        raise FileNotFoundError(f"reference tree not found at {REF_ROOT}")
    if threads is not None and "numba" not in sys.modules:
        os.environ.setdefault("NUMBA_NUM_THREADS", str(threads))
    _install_stubs()
    path = os.path.join(REF_ROOT, FILES[which])
    spec = importlib.util.spec_from_file_location(f"_srcfd_ref_{which}", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)  # the scripts guard their drivers with __main__
    if hasattr(mod, "CFDSolver"):
        mod.CFDSolver._save_results = lambda self, *a, **k: None
    _cache[which] = mod
    return mod
