"""ctypes binding of libsrcfd.so (include/srcfd.h).  No CPU fallback: if the CUDA library is
missing or no CUDA device is present, the calls fail loudly."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# SRCFD_LIB: an alternative build of the same library (kernel experiments); the default is the in-tree build
LIB_PATH = os.environ.get("SRCFD_LIB") or os.path.join(_HERE, "_lib", "libsrcfd.so")

SCHEME_UPWIND, SCHEME_QUICK = 0, 1
ORDER_GS_LEX, ORDER_JACOBI, ORDER_RED_BLACK, ORDER_RB_JACOBI = 0, 1, 2, 3
ORDERS = {"GS_LEX": ORDER_GS_LEX, "REFERENCE": ORDER_GS_LEX, "JACOBI": ORDER_JACOBI, "RED_BLACK": ORDER_RED_BLACK,
          "RB": ORDER_RED_BLACK, "RB_JACOBI": ORDER_RB_JACOBI, "RB_SOR": ORDER_RB_JACOBI}
OK, ERR_ARG, ERR_CUDA, ERR_NAN, ERR_DEADLOCK = 0, 1, 2, 3, 4


class SrcfdError(RuntimeError):
    pass


class Params(C.Structure):
    _fields_ = [
        ("nx", C.c_int32), ("ny", C.c_int32),
        ("dx", C.c_double), ("dy", C.c_double), ("volp", C.c_double),
        ("dt", C.c_double),
        ("nu", C.c_double), ("rho", C.c_double),
        ("scheme", C.c_int32),
        ("bc_types", (C.c_int32 * 4) * 3),
        ("bc_values", (C.c_double * 4) * 3),
        ("bfs_enabled", C.c_int32),
        ("bfs_step_h", C.c_double), ("bfs_h", C.c_double), ("bfs_Ub", C.c_double),
        ("relax_enabled", C.c_int32),
        ("relax", C.c_double * 3),
        ("inner_tol", C.c_double),
        ("inner_max", C.c_int32),
        ("sweep_order", C.c_int32),
        ("device", C.c_int32),
        ("max_ctas", C.c_int32),
        ("sor_omega", C.c_double),
        ("reserved", C.c_int32 * 6),
    ]


class CoarseResult(C.Structure):
    """srcfd_coarse_result (include/srcfd.h)."""
    _fields_ = [
        ("iterations", C.c_int64), ("converged", C.c_int32), ("nan_flag", C.c_int32),
        ("rms", C.c_double * 3), ("total_sweeps", C.c_int64 * 3), ("n_hist", C.c_int64),
        ("last_inner_rms", C.c_double * 3), ("residual", C.c_double * 3), ("last_sweeps", C.c_int32 * 3),
        ("reserved_", C.c_int32),
    ]


_dp = C.POINTER(C.c_double)
_lib = None

# every symbol include/srcfd.h declares (tests check that the shared object exports all of them)
SYMBOLS = [
    "srcfd_abi_version", "srcfd_last_error", "srcfd_device_count", "srcfd_create", "srcfd_destroy",
    "srcfd_set_params", "srcfd_synchronize", "srcfd_stream", "srcfd_upload", "srcfd_download",
    "srcfd_device_ptrs", "srcfd_host_alloc", "srcfd_host_free", "srcfd_trace_read", "srcfd_debug_read",
    "srcfd_initialize_fields", "srcfd_set_fields", "srcfd_step", "srcfd_status",
    "srcfd_reset_counters", "srcfd_solve", "srcfd_k_copy_new_to_old", "srcfd_k_apply_bc",
    "srcfd_k_apply_bc_configured", "srcfd_k_apply_bfs_inlet",
    "srcfd_k_linear_interpolation", "srcfd_k_update_flux", "srcfd_k_under_relax", "srcfd_k_correct_velocity",
    "srcfd_k_solve_pressure", "srcfd_jacobi_pass_max", "srcfd_k_jacobi_pass", "srcfd_k_jacobi_commit", "srcfd_jacobi_sums_ptr", "srcfd_k_jacobi_snapshot", "srcfd_k_solve_momentum", "srcfd_k_implicit_solve", "srcfd_launch_count",
    "srcfd_coarse_smem_bytes", "srcfd_coarse_fits", "srcfd_coarse_solve_batch",
    "srcfd_slab_configure", "srcfd_slab_export", "srcfd_slab_attach_ipc", "srcfd_slab_attach_local", "srcfd_slab_info", "srcfd_slab_kernel_stats",
    "srcfd_slab_exchange", "srcfd_slab_solve_pressure", "srcfd_slab_solve_momentum", "srcfd_slab_step",
    "srcfd_timing_enable", "srcfd_timing_read", "srcfd_timer_start", "srcfd_timer_stop",
    "srcfd_sr_last_error", "srcfd_sr_create", "srcfd_sr_destroy", "srcfd_sr_set_encoder", "srcfd_sr_set_decoder",
    "srcfd_sr_encode", "srcfd_sr_decode", "srcfd_sr_predict", "srcfd_sr_super_resolve", "srcfd_sr_decode_device", "srcfd_sr_launch_count",
    "srcfd_sr_set_precision", "srcfd_sr_tc_error", "srcfd_sr_debug_convT_tc",
]


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C sr-for-cfd_b200/csrc`).  There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        L.srcfd_last_error.restype = C.c_char_p
        for name in SYMBOLS:
            getattr(L, name)  # AttributeError here = header/library mismatch
        _lib = L
    return _lib


def check(rc: int):
    if rc == OK:
        return
    msg = lib().srcfd_last_error().decode("utf8", "replace")
    if rc == ERR_NAN:
        raise ValueError("Solver failed: NaN/Inf in residuals")   # PyCFD_ML_accelerated.py:487
    raise SrcfdError(f"libsrcfd error {rc}: {msg}")


def device_count() -> int:
    n = C.c_int(0)
    rc = lib().srcfd_device_count(C.byref(n))
    return n.value if rc == OK else 0


def pinned_zeros(shape, dtype=np.float64) -> np.ndarray:
    """np.zeros in page-locked memory (srcfd_host_alloc): same array semantics, full-rate upload/download.
    The block is wrapped through ctypes (no Python-version-specific buffer protocol) and freed by a finalizer tied to
    the ctypes object, which every numpy view keeps alive through its .base chain."""
    import weakref
    count = int(np.prod(shape))
    n = max(count * np.dtype(dtype).itemsize, 1)
    p = C.c_void_p()
    check(lib().srcfd_host_alloc(C.c_uint64(n), C.byref(p)))
    raw = (C.c_char * n).from_address(p.value)
    weakref.finalize(raw, lib().srcfd_host_free, C.c_void_p(p.value))
    a = np.frombuffer(raw, dtype=dtype, count=count).reshape(shape)
    a[...] = 0
    return a


def _ptr(a):
    if a is None:
        return None
    assert isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags["C_CONTIGUOUS"], \
        "libsrcfd wants C-contiguous float64 arrays"
    return a.ctypes.data_as(_dp)


class Handle:
    """Owns one srcfd_handle (one flow case on one device)."""

    def __init__(self, params: Params):
        self._h = C.c_void_p()
        self.params = params
        check(lib().srcfd_create(C.byref(params), C.byref(self._h)))
        self.nx, self.ny = params.nx, params.ny

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            lib().srcfd_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- parameters / transfer
    def set_params(self, params: Params):
        check(lib().srcfd_set_params(self._h, C.byref(params)))
        self.params = params

    def upload(self, Var=None, VarOld=None, Ff=None, residual=None):
        check(lib().srcfd_upload(self._h, _ptr(Var), _ptr(VarOld), _ptr(Ff), _ptr(residual)))

    def download(self, Var=None, VarOld=None, Ff=None, residual=None):
        check(lib().srcfd_download(self._h, _ptr(Var), _ptr(VarOld), _ptr(Ff), _ptr(residual)))

    def synchronize(self):
        check(lib().srcfd_synchronize(self._h))

    def stream(self) -> int:
        s = C.c_uint64(0)
        check(lib().srcfd_stream(self._h, C.byref(s)))
        return s.value

    def device_ptrs(self):
        a, b, c = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
        check(lib().srcfd_device_ptrs(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    # ---- composed path
    def initialize_fields(self, zero_first=True):
        check(lib().srcfd_initialize_fields(self._h, C.c_int(int(zero_first))))

    def set_fields(self, fields: np.ndarray):
        f = np.ascontiguousarray(fields)
        assert f.shape == (3, self.ny, self.nx), (f.shape, (3, self.ny, self.nx))
        if f.dtype == np.float32:
            check(lib().srcfd_set_fields(self._h, f.ctypes.data_as(C.c_void_p), C.c_int(1)))
        else:
            f = np.ascontiguousarray(f, dtype=np.float64)
            check(lib().srcfd_set_fields(self._h, f.ctypes.data_as(C.c_void_p), C.c_int(0)))

    def step(self, n_outer: int, crit=(1e-6, 1e-6, 1e-6)):
        c = (C.c_double * 3)(*[float(x) for x in crit])
        check(lib().srcfd_step(self._h, C.c_int64(int(n_outer)), c))

    def status(self):
        it, conv = C.c_int64(0), C.c_int32(0)
        rms, ls, ts = (C.c_double * 3)(), (C.c_int32 * 3)(), (C.c_int64 * 3)()
        check(lib().srcfd_status(self._h, C.byref(it), C.byref(conv), rms, ls, ts))
        return dict(iterations=it.value, converged=bool(conv.value), rms=np.array(rms[:]),
                    last_sweeps=np.array(ls[:]), total_sweeps=np.array(ts[:]))

    def reset_counters(self):
        check(lib().srcfd_reset_counters(self._h))

    def solve(self, max_iterations: int, crit=(1e-6, 1e-6, 1e-6)):
        c = (C.c_double * 3)(*[float(x) for x in crit])
        cap = int(max_iterations) // 100 + 1
        hist = np.zeros((cap, 3))
        it, nh, sec = C.c_int64(0), C.c_int64(0), C.c_double(0.0)
        check(lib().srcfd_solve(self._h, C.c_int64(int(max_iterations)), c, C.byref(it), C.byref(sec), _ptr(hist),
                                C.c_int64(cap), C.byref(nh)))
        return it.value, sec.value, hist[: nh.value]

    # ---- kernel level
    def k_copy_new_to_old(self):
        check(lib().srcfd_k_copy_new_to_old(self._h))

    def k_apply_bc(self, k):
        check(lib().srcfd_k_apply_bc(self._h, C.c_int(int(k))))

    def k_apply_bc_configured(self, k):
        check(lib().srcfd_k_apply_bc_configured(self._h, C.c_int(int(k))))

    def k_apply_bfs_inlet(self, k):
        check(lib().srcfd_k_apply_bfs_inlet(self._h, C.c_int(int(k))))

    def k_linear_interpolation(self):
        check(lib().srcfd_k_linear_interpolation(self._h))

    def k_update_flux(self):
        check(lib().srcfd_k_update_flux(self._h))

    def k_under_relax(self, k, alpha):
        check(lib().srcfd_k_under_relax(self._h, C.c_int(int(k)), C.c_double(float(alpha))))

    def k_correct_velocity(self):
        r = (C.c_double * 3)()
        check(lib().srcfd_k_correct_velocity(self._h, r))
        return np.array(r[:])

    def k_solve_pressure(self):
        n, rms = C.c_int32(0), C.c_double(0.0)
        check(lib().srcfd_k_solve_pressure(self._h, C.byref(n), C.byref(rms)))
        return n.value, rms.value

    def jacobi_pass_max(self) -> int:
        H = C.c_int(0)
        check(lib().srcfd_jacobi_pass_max(self._h, C.byref(H)))
        return H.value

    def k_jacobi_pass(self, nsweeps: int, own_row0: int, own_row1: int, recompute_rhs: bool = False,
                      commit: bool = True) -> np.ndarray:
        """One temporally blocked Jacobi pass of the pressure relaxation; returns the per-sweep sums of R^2 over the rows.
        commit=False leaves the plane untouched until k_jacobi_commit()."""
        sums = np.zeros(nsweeps)
        check(lib().srcfd_k_jacobi_pass(self._h, C.c_int(nsweeps), C.c_int(own_row0), C.c_int(own_row1),
                                        C.c_int(int(recompute_rhs)), C.c_int(int(commit)), C.c_int(0), _ptr(sums)))
        return sums

    def k_jacobi_pass_device(self, nsweeps: int, own_row0: int, own_row1: int, recompute_rhs: bool = False,
                             commit: bool = False, slot: int = 0):
        """Same, but the sums stay on the device (slot `slot` of jacobi_sums_ptr) and the call does not synchronise."""
        check(lib().srcfd_k_jacobi_pass(self._h, C.c_int(nsweeps), C.c_int(own_row0), C.c_int(own_row1),
                                        C.c_int(int(recompute_rhs)), C.c_int(int(commit)), C.c_int(slot), None))

    def k_jacobi_snapshot(self, restore: bool = False):
        check(lib().srcfd_k_jacobi_snapshot(self._h, C.c_int(int(restore))))

    def jacobi_sums_ptr(self) -> int:
        p = C.c_uint64(0)
        check(lib().srcfd_jacobi_sums_ptr(self._h, C.byref(p)))
        return p.value

    def k_jacobi_commit(self):
        check(lib().srcfd_k_jacobi_commit(self._h))

    def k_solve_momentum(self, k, scheme):
        n, rms = C.c_int32(0), C.c_double(0.0)
        check(lib().srcfd_k_solve_momentum(self._h, C.c_int(int(k)), C.c_int(int(scheme)), C.byref(n), C.byref(rms)))
        return n.value, rms.value

    def k_implicit_solve(self):
        check(lib().srcfd_k_implicit_solve(self._h))

    # ---- introspection
    def launch_count(self) -> int:
        n = C.c_int64(0)
        check(lib().srcfd_launch_count(self._h, C.byref(n)))
        return n.value

    def timer_start(self):
        check(lib().srcfd_timer_start(self._h))

    def timer_stop(self) -> float:
        ms = C.c_double(0.0)
        check(lib().srcfd_timer_stop(self._h, C.byref(ms)))
        return ms.value

    def timing_enable(self, on=True):
        check(lib().srcfd_timing_enable(self._h, C.c_int(int(on))))

    def timing_read(self):
        pm, mm, pn, mn = C.c_double(0), C.c_double(0), C.c_int64(0), C.c_int64(0)
        check(lib().srcfd_timing_read(self._h, C.byref(pm), C.byref(pn), C.byref(mm), C.byref(mn)))
        return dict(pressure_ms=pm.value, pressure_launches=pn.value, momentum_ms=mm.value, momentum_launches=mn.value)


def coarse_smem_bytes(nx: int, ny: int) -> int:
    n = C.c_uint64(0)
    check(lib().srcfd_coarse_smem_bytes(int(nx), int(ny), C.byref(n)))
    return n.value


def coarse_fits(nx: int, ny: int, device: int = 0):
    """The library's own size gate of the one-CTA solver on `device`; None when no CUDA device is present."""
    f = C.c_int(0)
    rc = lib().srcfd_coarse_fits(int(nx), int(ny), int(device), C.byref(f))
    if rc == ERR_CUDA:
        return None
    check(rc)
    return bool(f.value)


def coarse_solve_batch(params, max_iterations: int, crit, hist_cap: int = 0, state=None):
    """srcfd_coarse_solve_batch: CFDSolver(...).solve() of len(params) small cases in one launch (one CTA per case).

    params: sequence of Params sharing nx, ny, device; crit: (n, 3) or (3,) outer criteria {u, v, p}.
    state: optional (Var, VarOld, Ff) arrays (n, 3|3|4, nx+2, ny+2) to resume from (updated in place).
    Returns dict(Var, VarOld, Ff, iterations, converged, nan, rms, total_sweeps, hist, ms)."""
    n = len(params)
    if n < 1:
        raise ValueError("no cases")
    arr = (Params * n)(*params)
    nx, ny = params[0].nx, params[0].ny
    crit = np.ascontiguousarray(np.broadcast_to(np.asarray(crit, dtype=np.float64), (n, 3)))
    if state is None:
        Var = np.zeros((n, 3, nx + 2, ny + 2)); VarOld = np.zeros_like(Var); Ff = np.zeros((n, 4, nx + 2, ny + 2))
    else:
        Var, VarOld, Ff = state
        assert Var.shape == (n, 3, nx + 2, ny + 2) and VarOld.shape == Var.shape and Ff.shape == (n, 4, nx + 2, ny + 2)
    res = (CoarseResult * n)()
    hist = np.zeros((n, max(hist_cap, 1), 3)) if hist_cap > 0 else None
    ms = C.c_double(0.0)
    check(lib().srcfd_coarse_solve_batch(arr, n, C.c_int64(int(max_iterations)), _ptr(crit), int(state is not None),
                                         _ptr(Var), _ptr(VarOld), _ptr(Ff), res, _ptr(hist), C.c_int64(int(hist_cap)),
                                         C.byref(ms)))
    return dict(Var=Var, VarOld=VarOld, Ff=Ff,
                iterations=np.array([r.iterations for r in res], dtype=np.int64),
                converged=np.array([bool(r.converged) for r in res]),
                nan=np.array([bool(r.nan_flag) for r in res]),
                rms=np.array([[r.rms[k] for k in range(3)] for r in res]),
                last_inner_rms=np.array([[r.last_inner_rms[k] for k in range(3)] for r in res]),
                total_sweeps=np.array([[r.total_sweeps[k] for k in range(3)] for r in res], dtype=np.int64),
                residual=np.array([[r.residual[k] for k in range(3)] for r in res]),
                last_sweeps=np.array([[r.last_sweeps[k] for k in range(3)] for r in res], dtype=np.int64),
                hist=[hist[i, :res[i].n_hist].copy() for i in range(n)] if hist is not None else None,
                ms=ms.value)
