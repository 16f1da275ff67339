"""No-GPU checks of the boundary: the shared object loads, exports every symbol include/srcfd.h declares, and
refuses to run without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from srcfd import _capi as capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "srcfd.h")).read()
    return sorted(set(re.findall(r"\b(srcfd_[A-Za-z0-9_]+)\s*\(", txt)))


def test_header_and_library_agree():
    names = _declared()
    assert len(names) >= 40
    L = C.CDLL(capi.LIB_PATH)
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, f"declared in include/srcfd.h but not exported: {missing}"
    assert sorted(set(capi.SYMBOLS)) == names, "srcfd/_capi.py SYMBOLS is out of sync with the header"
    assert L.srcfd_abi_version() == 2


def test_params_struct_layout_matches_header():
    # int32 x2, double x6, int32 + pad, int32[12], double[12], int32 + pad, double x3, int32 + pad, double x3, double,
    # int32 x4, int32[8]  -> 4-byte members packed with natural alignment
    assert C.sizeof(capi.Params) == 8 + 48 + 8 + 48 + 96 + 8 + 24 + 8 + 24 + 8 + 16 + 32
    assert capi.Params.sor_omega.offset == C.sizeof(capi.Params) - 32 and capi.Params.reserved.size == 24   # carved out of reserved[8]


@pytest.mark.skipif(capi.device_count() > 0, reason="a CUDA device is present")
def test_no_cpu_fallback():
    p = capi.Params()
    p.nx = p.ny = 8
    p.inner_max = 10
    with pytest.raises(capi.SrcfdError, match="no CUDA device"):
        capi.Handle(p)
    from srcfd import sr
    with pytest.raises(capi.SrcfdError, match="no CUDA device"):
        sr.synthetic_decoder(0).predict([[0.0] * 50])


def test_coarse_batch_boundary_without_a_gpu():
    """srcfd_coarse_result layout, the shared-memory sizing (a pure host function) and the size gate of the one-CTA path."""
    from srcfd import solver as S
    # int64, int32 x2, double[3], int64[3], int64, double[3], double[3], int32[3] + int32 pad
    assert C.sizeof(capi.CoarseResult) == 8 + 8 + 24 + 24 + 8 + 24 + 24 + 16
    P = 12 * 12
    ring = max((10 - 1 + 2 * 5 - 1) // 2 + 2, (10 - 1 + 3 * 4 - 1) // 3 + 2)
    assert capi.coarse_smem_bytes(10, 10) == 8 * (12 * P + ring * 10 * 5)
    assert S.fits_one_cta(10, 10) and S.fits_one_cta(30, 30) and S.fits_one_cta(1, 1)
    assert not S.fits_one_cta(48, 40) and not S.fits_one_cta(400, 400) and not S.fits_one_cta(64, 16)
    with pytest.raises(capi.SrcfdError):
        capi.coarse_smem_bytes(0, 10)


@pytest.mark.skipif(capi.device_count() > 0, reason="a CUDA device is present")
def test_coarse_batch_has_no_cpu_fallback():
    p = capi.Params()
    p.nx = p.ny = 10
    p.dx = p.dy = 0.1; p.volp = 0.01; p.dt = 1e-3; p.nu = 0.01; p.rho = 1.0
    p.inner_max, p.inner_tol = 10, 1e-6
    with pytest.raises(capi.SrcfdError, match="no CUDA device"):
        capi.coarse_solve_batch([p], 5, (1e-6,) * 3)
