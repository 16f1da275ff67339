"""The reference's module-level kernel functions, same names and positional signatures
(PyCFD_ML_accelerated.py:110-328, bfs_ml_accelerated.py:233-464), running on libsrcfd.

The reference's functions take raw host ndarrays and mutate them in place; so do these: every call
uploads the arrays it reads, launches the CUDA kernel through the C ABI and downloads the arrays it
writes.  (Inside CFDSolver.solve() nothing bounces through the host; these entry points exist because
the reference's workflow code calls copy_new_to_old / linear_interpolation directly, e.g.
PyCFD_ML_accelerated.py:946-948, and because they are the per-kernel parity surface.)

The inner solves return the number of sweeps executed (the reference returns None) and accept three
keyword extensions: sweep_order, tolerance, max_iter (defaults = the reference's hard-coded values).
"""
from __future__ import annotations

import os
from collections import OrderedDict

import numpy as np

from . import _capi as capi

_cache: "OrderedDict[tuple, capi.Handle]" = OrderedDict()
_CACHE_MAX = 4
default_device = 0


def _default_order():
    return os.environ.get("SRCFD_ORDER", "GS_LEX")


def _order(o):
    if o is None:
        o = _default_order()
    return capi.ORDERS[o.upper()] if isinstance(o, str) else int(o)


def _handle(Nx, Ny, **kw) -> capi.Handle:
    """A cached device context for an (Nx, Ny) grid, re-parameterised for this call."""
    p = capi.Params()
    p.nx, p.ny = int(Nx), int(Ny)
    p.dx, p.dy = float(kw.get("dx", 1.0)), float(kw.get("dy", 1.0))
    p.volp = float(kw.get("volp", p.dx * p.dy))
    p.dt, p.nu, p.rho = float(kw.get("dt", 1.0)), float(kw.get("nu", 1.0)), float(kw.get("rho", 1.0))
    p.scheme = int(kw.get("scheme", capi.SCHEME_UPWIND))
    bt, bv = kw.get("bc_types"), kw.get("bc_values")
    k = kw.get("k", 0)
    if bt is not None:
        for s in range(4):
            p.bc_types[k][s] = int(bt[s])
            p.bc_values[k][s] = float(bv[s])
    bfs = kw.get("bfs")
    if bfs is not None:
        p.bfs_enabled = 1
        p.bfs_step_h, p.bfs_h, p.bfs_Ub = (float(x) for x in bfs)
    p.inner_tol = float(kw.get("tolerance", 1e-6))
    p.inner_max = 1000
    want_max = int(kw.get("max_iter", 1000))
    p.sweep_order = _order(kw.get("sweep_order"))
    p.sor_omega = float(kw.get("sor_omega", 1.0) or 1.0)
    p.device = default_device
    key = (p.nx, p.ny, p.device, max(1000, want_max))
    h = _cache.get(key)
    if h is None:
        p.inner_max = max(1000, want_max)      # capacity of the per-sweep buffers
        h = capi.Handle(p)
        _cache[key] = h
        while len(_cache) > _CACHE_MAX:
            _cache.popitem(last=False)[1].close()
    else:
        _cache.move_to_end(key)
    p.inner_max = want_max
    h.set_params(p)
    h.reset_counters()
    return h


def _chk(a, planes, Nx, Ny, name):
    if not (isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
            and a.shape == (planes, Nx + 2, Ny + 2)):
        raise ValueError(f"{name} must be a C-contiguous float64 array of shape ({planes}, Nx+2, Ny+2)")


def copy_new_to_old(Var, VarOld, nVar, Nx, Ny):
    """LDC.py:110-115."""
    _chk(Var, 3, Nx, Ny, "Var"); _chk(VarOld, 3, Nx, Ny, "VarOld")
    h = _handle(Nx, Ny)
    if nVar != 3:                      # partial copy: keep the planes the reference would not touch
        h.upload(Var=Var, VarOld=VarOld)
        keep = VarOld[nVar:].copy()
        h.k_copy_new_to_old()
        h.download(VarOld=VarOld)
        VarOld[nVar:] = keep
        return
    h.upload(Var=Var)
    h.k_copy_new_to_old()
    h.download(VarOld=VarOld)


def apply_bc_configured(Var, k, Nx, Ny, bc_types, bc_values):
    """LDC.py:117-145."""
    _chk(Var, 3, Nx, Ny, "Var")
    h = _handle(Nx, Ny, k=k, bc_types=bc_types, bc_values=bc_values)
    h.upload(Var=Var)
    h.k_apply_bc_configured(k)
    h.download(Var=Var)


def apply_bfs_inlet(Var, k, Nx, Ny, dy, step_height, h_channel, Ub):
    """CFDSolver._apply_bfs_inlet (BFS.py:524-562) as a free function."""
    _chk(Var, 3, Nx, Ny, "Var")
    h = _handle(Nx, Ny, dy=dy, bfs=(step_height, h_channel, Ub))
    h.upload(Var=Var)
    h.k_apply_bfs_inlet(k)
    h.download(Var=Var)


def linear_interpolation(Var, Ff, Nx, Ny, dx, dy):
    """LDC.py:147-154."""
    _chk(Var, 3, Nx, Ny, "Var"); _chk(Ff, 4, Nx, Ny, "Ff")
    h = _handle(Nx, Ny, dx=dx, dy=dy)
    h.upload(Var=Var, Ff=Ff)
    h.k_linear_interpolation()
    h.download(Ff=Ff)


def update_flux(Var, Ff, dt, rho, Nx, Ny, dx, dy):
    """LDC.py:239-246."""
    _chk(Var, 3, Nx, Ny, "Var"); _chk(Ff, 4, Nx, Ny, "Ff")
    h = _handle(Nx, Ny, dx=dx, dy=dy, dt=dt, rho=rho)
    h.upload(Var=Var, Ff=Ff)
    h.k_update_flux()
    h.download(Ff=Ff)


def under_relax_field(Var, VarOld, k, Nx, Ny, alpha):
    """BFS.py:371-375."""
    _chk(Var, 3, Nx, Ny, "Var"); _chk(VarOld, 3, Nx, Ny, "VarOld")
    h = _handle(Nx, Ny)
    h.upload(Var=Var, VarOld=VarOld)
    h.k_under_relax(k, alpha)
    h.download(Var=Var)


def solve_pressure(Var, Ff, Nx, Ny, dx, dy, dt, rho, volp, *, sweep_order=None, tolerance=1e-6, max_iter=1000,
                   sor_omega=1.0):
    """LDC.py:292-314.  sor_omega (RED_BLACK order only): p += omega * R/ap, red-black SOR."""
    _chk(Var, 3, Nx, Ny, "Var"); _chk(Ff, 4, Nx, Ny, "Ff")
    h = _handle(Nx, Ny, dx=dx, dy=dy, dt=dt, rho=rho, volp=volp, sweep_order=sweep_order, tolerance=tolerance,
                max_iter=max_iter, sor_omega=sor_omega)
    h.upload(Var=Var, Ff=Ff)
    n, _ = h.k_solve_pressure()
    h.download(Var=Var)
    return n


def _momentum(scheme, Var, VarOld, Ff, k, Nx, Ny, dx, dy, dt, nu, volp, sweep_order, tolerance, max_iter):
    _chk(Var, 3, Nx, Ny, "Var"); _chk(VarOld, 3, Nx, Ny, "VarOld"); _chk(Ff, 4, Nx, Ny, "Ff")
    h = _handle(Nx, Ny, dx=dx, dy=dy, dt=dt, nu=nu, volp=volp, scheme=scheme, sweep_order=sweep_order,
                tolerance=tolerance, max_iter=max_iter)
    h.upload(Var=Var, VarOld=VarOld, Ff=Ff)
    n, _ = h.k_solve_momentum(k, scheme)
    h.download(Var=Var)
    return n


def solve_momentum_upwind(Var, VarOld, Ff, k, Nx, Ny, dx, dy, dt, nu, volp, *, sweep_order=None, tolerance=1e-6,
                          max_iter=1000):
    """LDC.py:270-290."""
    return _momentum(capi.SCHEME_UPWIND, Var, VarOld, Ff, k, Nx, Ny, dx, dy, dt, nu, volp, sweep_order, tolerance, max_iter)


def solve_momentum_quick(Var, VarOld, Ff, k, Nx, Ny, dx, dy, dt, nu, volp, *, sweep_order=None, tolerance=1e-6,
                         max_iter=1000):
    """LDC.py:248-268."""
    return _momentum(capi.SCHEME_QUICK, Var, VarOld, Ff, k, Nx, Ny, dx, dy, dt, nu, volp, sweep_order, tolerance, max_iter)


def correct_velocity(Var, VarOld, dt, rho, Nx, Ny, dx, dy, residual=None):
    """LDC.py:316-328 when `residual` (length-3 array, accumulated into) is given;
    BFS.py:445-464 when it is omitted: returns (res_u, res_v, res_p)."""
    _chk(Var, 3, Nx, Ny, "Var"); _chk(VarOld, 3, Nx, Ny, "VarOld")
    h = _handle(Nx, Ny, dx=dx, dy=dy, dt=dt, rho=rho)
    start = np.zeros(3) if residual is None else np.ascontiguousarray(residual, dtype=np.float64)
    h.upload(Var=Var, VarOld=VarOld, residual=start)
    out = h.k_correct_velocity()
    h.download(Var=Var)
    if residual is None:
        return float(out[0]), float(out[1]), float(out[2])
    residual[:] = out
