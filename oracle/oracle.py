"""ctypes front end of the CPU oracle (oracle/pycfd_oracle.c).

TEST INFRASTRUCTURE ONLY -- parity status: PINNED (see the header of
pycfd_oracle.c).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module; the product
package (sr-for-cfd_b200/) never does.

The module-level functions keep the reference's positional signatures
(PyCFD_ML_accelerated.py:110-328, bfs_ml_accelerated.py:233-464) so parity
tests read like calls into the reference.  The solve_* functions take three
extra keyword arguments (order, tolerance, max_iter) whose defaults are the
reference's hard-coded behaviour (in-place lexicographic sweep, 1e-6, 1000)
and RETURN the number of sweeps executed (the reference returns None).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libpycfd_oracle.so")

ORDER_GS_LEX, ORDER_JACOBI, ORDER_RB, ORDER_GS_OMP = 0, 1, 2, 3
ORDER_RB_JACOBI = 4      # composed solver only: momentum Jacobi, pressure red-black (SOR factor via set_sor_omega)
SCHEME_UPWIND, SCHEME_QUICK = 0, 1


def build(force: bool = False) -> str:
    """Compile the oracle with the recipe in oracle/Makefile."""
    src = os.path.join(_HERE, "pycfd_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "all"])
    return _LIB_PATH


_lib = None
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)


class OrcParams(C.Structure):
    _fields_ = [
        ("Nx", C.c_int32), ("Ny", C.c_int32),
        ("dx", C.c_double), ("dy", C.c_double), ("volp", C.c_double),
        ("dt", C.c_double), ("nu", C.c_double), ("rho", C.c_double),
        ("scheme", C.c_int32),
        ("bc_types", (C.c_int32 * 4) * 3),
        ("bc_values", (C.c_double * 4) * 3),
        ("bfs", C.c_int32),
        ("step_h", C.c_double), ("h", C.c_double), ("Ub", C.c_double),
        ("use_relax", C.c_int32),
        ("alpha", C.c_double * 3),
        ("order", C.c_int32),
        ("inner_tol", C.c_double),
        ("inner_max", C.c_int32),
    ]


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_solve_pressure.restype = C.c_int
        L.orc_solve_momentum_upwind.restype = C.c_int
        L.orc_solve_momentum_quick.restype = C.c_int
        L.orc_convergence_check.restype = C.c_int
        L.orc_solve.restype = C.c_int64
        L.orc_num_threads.restype = C.c_int
        _lib = L
    return _lib


def _d(a: np.ndarray):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"], "oracle wants C-contiguous float64"
    return a.ctypes.data_as(_dp)


def _chk(Var, Nx, Ny, planes=3):
    assert Var.shape == (planes, Nx + 2, Ny + 2), (Var.shape, Nx, Ny)


# --------------------------------------------------------------------------
# kernel-level functions (reference signatures)
# --------------------------------------------------------------------------
def copy_new_to_old(Var, VarOld, nVar, Nx, Ny):
    lib().orc_copy_new_to_old(_d(Var), _d(VarOld), C.c_int(nVar), C.c_int(Nx), C.c_int(Ny))


def apply_bc_configured(Var, k, Nx, Ny, bc_types, bc_values):
    t = np.ascontiguousarray(bc_types, dtype=np.int32)
    v = np.ascontiguousarray(bc_values, dtype=np.float64)
    lib().orc_apply_bc_configured(_d(Var), C.c_int(k), C.c_int(Nx), C.c_int(Ny), t.ctypes.data_as(_ip), _d(v))


def apply_bfs_inlet(Var, k, Nx, Ny, dy, step_h, h, Ub):
    lib().orc_apply_bfs_inlet(_d(Var), C.c_int(k), C.c_int(Nx), C.c_int(Ny), C.c_double(dy),
                              C.c_double(step_h), C.c_double(h), C.c_double(Ub))


def linear_interpolation(Var, Ff, Nx, Ny, dx, dy):
    _chk(Var, Nx, Ny); _chk(Ff, Nx, Ny, 4)
    lib().orc_linear_interpolation(_d(Var), _d(Ff), C.c_int(Nx), C.c_int(Ny), C.c_double(dx), C.c_double(dy))


def update_flux(Var, Ff, dt, rho, Nx, Ny, dx, dy):
    _chk(Var, Nx, Ny); _chk(Ff, Nx, Ny, 4)
    lib().orc_update_flux(_d(Var), _d(Ff), C.c_double(dt), C.c_double(rho), C.c_int(Nx), C.c_int(Ny),
                          C.c_double(dx), C.c_double(dy))


def under_relax_field(Var, VarOld, k, Nx, Ny, alpha):
    lib().orc_under_relax_field(_d(Var), _d(VarOld), C.c_int(k), C.c_int(Nx), C.c_int(Ny), C.c_double(alpha))


def _hist(max_iter, want):
    return np.zeros(max_iter, dtype=np.float64) if want else None


def solve_pressure(Var, Ff, Nx, Ny, dx, dy, dt, rho, volp, order=ORDER_GS_LEX, tolerance=1e-6,
                   max_iter=1000, rms_hist=False):
    _chk(Var, Nx, Ny); _chk(Ff, Nx, Ny, 4)
    last = C.c_double(0.0)
    h = _hist(max_iter, rms_hist)
    n = lib().orc_solve_pressure(_d(Var), _d(Ff), C.c_int(Nx), C.c_int(Ny), C.c_double(dx), C.c_double(dy),
                                 C.c_double(dt), C.c_double(rho), C.c_double(volp), C.c_int(order),
                                 C.c_double(tolerance), C.c_int(max_iter), C.byref(last),
                                 _d(h) if h is not None else None)
    return (n, h[:n]) if rms_hist else n


def _momentum(fn, Var, VarOld, Ff, k, Nx, Ny, dx, dy, dt, nu, volp, order, tolerance, max_iter, rms_hist):
    _chk(Var, Nx, Ny); _chk(VarOld, Nx, Ny); _chk(Ff, Nx, Ny, 4)
    last = C.c_double(0.0)
    h = _hist(max_iter, rms_hist)
    n = fn(_d(Var), _d(VarOld), _d(Ff), C.c_int(k), C.c_int(Nx), C.c_int(Ny), C.c_double(dx), C.c_double(dy),
           C.c_double(dt), C.c_double(nu), C.c_double(volp), C.c_int(order), C.c_double(tolerance),
           C.c_int(max_iter), C.byref(last), _d(h) if h is not None else None)
    return (n, h[:n]) if rms_hist else n


def solve_momentum_upwind(Var, VarOld, Ff, k, Nx, Ny, dx, dy, dt, nu, volp, order=ORDER_GS_LEX,
                          tolerance=1e-6, max_iter=1000, rms_hist=False):
    return _momentum(lib().orc_solve_momentum_upwind, Var, VarOld, Ff, k, Nx, Ny, dx, dy, dt, nu, volp,
                     order, tolerance, max_iter, rms_hist)


def solve_momentum_quick(Var, VarOld, Ff, k, Nx, Ny, dx, dy, dt, nu, volp, order=ORDER_GS_LEX,
                         tolerance=1e-6, max_iter=1000, rms_hist=False):
    return _momentum(lib().orc_solve_momentum_quick, Var, VarOld, Ff, k, Nx, Ny, dx, dy, dt, nu, volp,
                     order, tolerance, max_iter, rms_hist)


def correct_velocity(Var, VarOld, dt, rho, Nx, Ny, dx, dy, residual=None):
    """LDC form accumulates into `residual` (LDC.py:316-328); with residual=None it
    returns (res_u, res_v, res_p) like the BFS form (BFS.py:445-464)."""
    _chk(Var, Nx, Ny); _chk(VarOld, Nx, Ny)
    ret = residual is None
    if ret:
        residual = np.zeros(3)
    lib().orc_correct_velocity(_d(Var), _d(VarOld), C.c_double(dt), C.c_double(rho), C.c_int(Nx), C.c_int(Ny),
                               C.c_double(dx), C.c_double(dy), _d(residual))
    if ret:
        return float(residual[0]), float(residual[1]), float(residual[2])


# --------------------------------------------------------------------------
# composed solver
# --------------------------------------------------------------------------
@dataclass
class Case:
    """Everything CFDSolver.__init__ is given (LDC.py:333-349, BFS.py:473-496)."""
    nx: int
    ny: int
    lx: float = 1.0
    ly: float = 1.0
    Re: float = 100.0
    rho: float = 1.0
    dt: float = 1e-3
    scheme: str = "QUICK"
    # [k][left,right,top,bottom]; defaults = lid-driven cavity (LDC.py:47-67)
    bc_types: list = field(default_factory=lambda: [[0, 0, 0, 0], [0, 0, 0, 0], [1, 1, 1, 1]])
    bc_values: list = field(default_factory=lambda: [[0.0, 0.0, 1.0, 0.0], [0.0] * 4, [0.0] * 4])
    bfs: bool = False
    step_h: float = 1.0
    h: float = 2.0
    Ub: float = 1.0
    relax: tuple | None = None          # (alpha_u, alpha_v, alpha_p) or None (LDC: no relaxation calls)
    order: int = ORDER_GS_LEX
    inner_tol: float = 1e-6
    inner_max: int = 1000

    def params(self) -> OrcParams:
        p = OrcParams()
        p.Nx, p.Ny = self.nx, self.ny
        p.dx = self.lx / self.nx          # MeshParameters, LDC.py:76-78
        p.dy = self.ly / self.ny
        p.volp = p.dx * p.dy
        p.dt, p.rho = self.dt, self.rho
        p.nu = 1.0 / self.Re              # FluidProperties, LDC.py:86
        p.scheme = SCHEME_QUICK if self.scheme == "QUICK" else SCHEME_UPWIND
        for k in range(3):
            for s in range(4):
                p.bc_types[k][s] = int(self.bc_types[k][s])
                p.bc_values[k][s] = float(self.bc_values[k][s])
        p.bfs = int(self.bfs)
        p.step_h, p.h, p.Ub = self.step_h, self.h, self.Ub
        p.use_relax = int(self.relax is not None)
        a = self.relax if self.relax is not None else (1.0, 1.0, 1.0)
        for k in range(3):
            p.alpha[k] = a[k]
        p.order = self.order
        p.inner_tol, p.inner_max = self.inner_tol, self.inner_max
        return p


def bfs_case(nx, ny, Re=400.0, dt=2e-3, scheme="UPWIND", lx=10.0, ly=3.0, step_h=1.0, h=2.0, Ub=1.0,
             relax=(0.5, 0.5, 0.2), u_left=0.0, **kw) -> Case:
    """BFS boundary set of bfs_ml_accelerated.py:1789-1811 (the u_left value is a placeholder:
    the inlet override replaces the whole left ghost column)."""
    return Case(nx=nx, ny=ny, lx=lx, ly=ly, Re=Re, dt=dt, scheme=scheme,
                bc_types=[[0, 1, 0, 0], [0, 1, 0, 0], [1, 0, 1, 1]],
                bc_values=[[u_left, 0.0, 0.0, 0.0], [0.0] * 4, [0.0] * 4],
                bfs=True, step_h=step_h, h=h, Ub=Ub, relax=relax, **kw)


class OracleSolver:
    """State owner with the reference CFDSolver's array attributes."""

    def __init__(self, case: Case, initialize: bool = True, bfs_at_init: bool | None = None):
        self.case = case
        self.p = case.params()
        nx, ny = case.nx, case.ny
        self.Var = np.zeros((3, nx + 2, ny + 2))
        self.VarOld = np.zeros((3, nx + 2, ny + 2))
        self.Ff = np.zeros((4, nx + 2, ny + 2))
        self.residual = np.zeros(3)
        self.sweeps = np.zeros(3, dtype=np.int32)
        self.total_sweeps = np.zeros(3, dtype=np.int64)
        if initialize:
            # "bfs code given by sir.py":856-861 sets case_type only AFTER the constructor
            # ran _initialize_fields, i.e. the first BC pass there has no inlet override.
            if bfs_at_init is not None and bool(bfs_at_init) != bool(case.bfs):
                saved = self.p.bfs
                self.p.bfs = int(bfs_at_init)
                self.initialize_fields(True)
                self.p.bfs = saved
            else:
                self.initialize_fields(True)

    def initialize_fields(self, zero_first: bool):
        lib().orc_initialize_fields(C.byref(self.p), _d(self.Var), _d(self.VarOld), _d(self.Ff),
                                    C.c_int(int(zero_first)))

    def apply_bc(self, k: int):
        lib().orc_apply_bc_wrapper(C.byref(self.p), _d(self.Var), C.c_int(k))

    def set_interior(self, fields):
        """Warm-start injection, LDC.py:936-948: fields are (ny,nx) arrays u,v,p."""
        for k, name in enumerate(("u", "v", "p")):
            self.Var[k, 1:-1, 1:-1] = np.asarray(fields[name]).T
        self.initialize_fields(False)

    def implicit_solve(self):
        lib().orc_implicit_solve(C.byref(self.p), _d(self.Var), _d(self.VarOld), _d(self.Ff),
                                 _d(self.residual), self.sweeps.ctypes.data_as(_ip))
        return self.sweeps.copy()

    def convergence_check(self, crit=(1e-6, 1e-6, 1e-6)):
        rms = np.zeros(3)
        c = lib().orc_convergence_check(C.byref(self.p), _d(self.Var), _d(self.VarOld), _d(self.residual),
                                        _d(np.asarray(crit, dtype=np.float64)), _d(rms))
        if c < 0:
            raise ValueError("Solver failed: NaN/Inf in residuals")
        return bool(c), rms

    def solve(self, max_iterations: int, crit=(1e-6, 1e-6, 1e-6)):
        cap = max_iterations // 100 + 1
        hist = np.zeros((cap, 3))
        nh = C.c_int64(0)
        last = np.zeros(3)
        n = lib().orc_solve(C.byref(self.p), _d(self.Var), _d(self.VarOld), _d(self.Ff),
                            C.c_int64(max_iterations), _d(np.asarray(crit, dtype=np.float64)), _d(last),
                            _d(hist), C.c_int64(cap), C.byref(nh),
                            self.total_sweeps.ctypes.data_as(C.POINTER(C.c_int64)))
        if n < 0:
            raise ValueError("Solver failed: NaN/Inf in residuals")
        return int(n), last, hist[: nh.value]


def set_sor_omega(w: float):
    """Over-relaxation factor of the ORDER_RB pressure sweep (1.0 = plain red-black Gauss-Seidel)."""
    lib().orc_set_sor_omega(C.c_double(float(w)))


def num_threads() -> int:
    return lib().orc_num_threads()


def set_num_threads(n: int):
    lib().orc_set_num_threads(C.c_int(n))
