"""Host-side mirror of the reference's solver classes, running on libsrcfd (CUDA, sm_100a).

Same names, constructor arguments, attributes and error behaviour as PyCFD_ML_accelerated.py:41-658
(LDC.py) and bfs_ml_accelerated.py:152-866 (BFS.py), so the reference's workflow functions -- and a
user's scripts -- run unchanged on top of it:

    solver = CFDSolver(mesh, fluid, solver_settings, bc[, step_height, h, Ub])
    solver.Var[k, 1:-1, 1:-1] = ...; solver._apply_bc_wrapper(k)
    count, seconds = solver.solve(output_name, verbose)
    solver.Var / solver.residual_history

Var, VarOld, Ff, residual stay host numpy arrays exactly as in the reference; each method moves what
it needs to the device and back.  `solve()` uploads once, iterates on the GPU with no host round trip
inside an iteration, and downloads once.
"""
from __future__ import annotations

import os
import time
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import numpy as np

from . import _capi as capi


# --------------------------------------------------------------------------------------------------
# configuration classes (LDC.py:41-105, BFS.py:152-226)
# --------------------------------------------------------------------------------------------------
@dataclass
class BoundaryCondition:
    """'dirichlet' or 'neumann' with a value (LDC.py:41-45)."""
    type: str
    value: float = 0.0


class BoundaryConditions:
    """Lid-driven-cavity defaults: u_top = 1, walls elsewhere, Neumann pressure (LDC.py:47-67)."""

    def __init__(self):
        self.u_boundaries = {
            'left': BoundaryCondition('dirichlet', 0.0), 'right': BoundaryCondition('dirichlet', 0.0),
            'top': BoundaryCondition('dirichlet', 1.0), 'bottom': BoundaryCondition('dirichlet', 0.0)}
        self.v_boundaries = {s: BoundaryCondition('dirichlet', 0.0) for s in ('left', 'right', 'top', 'bottom')}
        self.p_boundaries = {s: BoundaryCondition('neumann', 0.0) for s in ('left', 'right', 'top', 'bottom')}


class BFSBoundaryConditions(BoundaryConditions):
    """BFS.py's container defaults (BFS.py:158-178): u_left = 1 placeholder, everything else zero."""

    def __init__(self):
        super().__init__()
        self.u_boundaries['left'] = BoundaryCondition('dirichlet', 1.0)
        self.u_boundaries['top'] = BoundaryCondition('dirichlet', 0.0)


class MeshParameters:
    def __init__(self, nx: int = 100, ny: int = 100, lx: float = 1.0, ly: float = 1.0):
        self.nx, self.ny, self.lx, self.ly = nx, ny, lx, ly
        self.dx = lx / nx
        self.dy = ly / ny
        self.volp = self.dx * self.dy


class FluidProperties:
    def __init__(self, Re: float = 100.0, rho: float = 1.0):
        self.Re = Re
        self.rho = rho
        self.nu = 1.0 / Re


class SolverSettings:
    """LDC.py:88-105 plus BFS.py's relaxation_factors (BFS.py:198-226).

    Extensions (keyword-only, defaults = the reference's hard-coded behaviour):
      sweep_order  'GS_LEX' (reference order, bit-identical results) | 'JACOBI' | 'RED_BLACK'
      inner_tolerance / inner_max_iter   the constants at LDC.py:250-251
    """

    def __init__(self, dt: float = 0.001, max_iterations: int = 100000,
                 convergence_criteria: Dict[str, float] = None, scheme: str = 'QUICK',
                 relaxation_factors: Dict[str, float] = None, *, sweep_order: str = None,
                 inner_tolerance: float = 1e-6, inner_max_iter: int = 1000, sor_omega: float = 1.0):
        self.dt = dt
        self.sor_omega = sor_omega           # opt-in: over-relaxation of the RED_BLACK pressure sweep
        self.max_iterations = max_iterations
        self.scheme = scheme
        self.convergence_criteria = convergence_criteria if convergence_criteria is not None else {
            'u': 1e-6, 'v': 1e-6, 'p': 1e-6, 'continuity': 1e-6}
        self.relaxation_factors = relaxation_factors
        self.sweep_order = sweep_order or os.environ.get("SRCFD_ORDER", "GS_LEX")
        self.inner_tolerance = inner_tolerance
        self.inner_max_iter = inner_max_iter


class BFSSolverSettings(SolverSettings):
    """BFS.py:198-226: UPWIND default and relaxation 0.5/0.5/0.2 when none is given."""

    def __init__(self, dt: float = 0.001, max_iterations: int = 100000,
                 convergence_criteria: Dict[str, float] = None, scheme: str = 'UPWIND',
                 relaxation_factors: Dict[str, float] = None, **kw):
        super().__init__(dt, max_iterations, convergence_criteria, scheme, relaxation_factors, **kw)
        if relaxation_factors is None:
            self.relaxation_factors = {'u': 0.5, 'v': 0.5, 'p': 0.2}


_SIDES = ('left', 'right', 'top', 'bottom')


def get_bc_arrays(bc, k: int):
    """CFDSolver._get_bc_arrays (LDC.py:351-375) for a BoundaryConditions container."""
    bc_dict = (bc.u_boundaries, bc.v_boundaries, bc.p_boundaries)[min(k, 2)]
    bc_types = np.array([0 if bc_dict[s].type == 'dirichlet' else 1 for s in _SIDES], dtype=np.int32)
    bc_values = np.array([bc_dict[s].value for s in _SIDES], dtype=np.float64)
    return bc_types, bc_values


def make_params(mesh, fluid, settings, bc, case_type=None, step_height=1.0, h=2.0, Ub=1.0, relaxed=False,
                device: int = 0, max_ctas: int = 0) -> capi.Params:
    """The srcfd_params of one case (what CFDSolver.__init__ of LDC.py:333-349 / BFS.py:473-496 fixes)."""
    p = capi.Params()
    m, f, s = mesh, fluid, settings
    p.nx, p.ny = m.nx, m.ny
    p.dx, p.dy, p.volp = m.dx, m.dy, m.volp
    p.dt, p.nu, p.rho = s.dt, f.nu, f.rho
    p.scheme = capi.SCHEME_QUICK if s.scheme == 'QUICK' else capi.SCHEME_UPWIND
    for k in range(3):
        t, v = get_bc_arrays(bc, k)
        for i in range(4):
            p.bc_types[k][i] = int(t[i])
            p.bc_values[k][i] = float(v[i])
    p.bfs_enabled = int(case_type == 'BFS')
    p.bfs_step_h, p.bfs_h, p.bfs_Ub = float(step_height), float(h), float(Ub)
    rf = getattr(s, 'relaxation_factors', None)
    p.relax_enabled = int(bool(relaxed))
    rf = rf or {}
    p.relax[0], p.relax[1], p.relax[2] = rf.get('u', 0.5), rf.get('v', 0.5), rf.get('p', 0.2)
    p.inner_tol = getattr(s, 'inner_tolerance', 1e-6)
    p.inner_max = getattr(s, 'inner_max_iter', 1000)
    order = getattr(s, 'sweep_order', 'GS_LEX')
    p.sweep_order = capi.ORDERS[order.upper()] if isinstance(order, str) else int(order)
    p.device, p.max_ctas = device, max_ctas
    p.sor_omega = float(getattr(s, 'sor_omega', 1.0) or 1.0)
    return p


_ONE_CTA_SMEM = 227 * 1024 - 2048      # opt-in shared memory per CTA on sm_100a, less the kernel's static part
_ONE_CTA_THREADS = 640                # register-limited block size of k_coarse_solve


def fits_one_cta(nx: int, ny: int, device: int = 0) -> bool:
    """True when srcfd_coarse_solve_batch accepts the grid (state in shared memory, one thread per 2 cells of a row).
    The library answers from the device's and the kernel's own attributes (srcfd_coarse_fits); the constants above
    only serve where no CUDA device is present (documentation, CPU tests)."""
    ans = capi.coarse_fits(nx, ny, device)
    if ans is not None:
        return ans
    threads = ((nx * ((ny + 1) // 2) + 31) // 32) * 32 + 32
    return threads <= _ONE_CTA_THREADS and capi.coarse_smem_bytes(nx, ny) <= _ONE_CTA_SMEM


# --------------------------------------------------------------------------------------------------
# CFDSolver
# --------------------------------------------------------------------------------------------------
class CFDSolver:
    """LDC.py:331-501 / BFS.py:471-706 on the GPU.

    `case_type`, `step_height`, `h`, `Ub` may be assigned after construction, as
    "bfs code given by sir.py":856-861 does; parameters are re-read before every device call.
    `relaxed` selects BFS.py's _implicit_solve (under-relaxation calls present) over LDC.py's.
    """

    case_name = "lid driven cavity"

    def __init__(self, mesh: MeshParameters, fluid: FluidProperties, solver_settings: SolverSettings,
                 bc: BoundaryConditions, step_height: float = None, h: float = None, Ub: float = None,
                 *, device: int = 0, max_ctas: int = 0, relaxed: bool = None):
        self.mesh, self.fluid, self.settings, self.bc = mesh, fluid, solver_settings, bc
        bfs = step_height is not None or h is not None or Ub is not None
        self.case_type = 'BFS' if bfs else None
        self.step_height = 1.0 if step_height is None else step_height
        self.h = 2.0 if h is None else h
        self.Ub = 1.0 if Ub is None else Ub
        self.relaxed = bfs if relaxed is None else relaxed
        self.device, self.max_ctas = device, max_ctas
        self.nVar = 3
        # the reference's host arrays (LDC.py:340-345), page-locked so the per-call copies run at PCIe rate
        self.Var = capi.pinned_zeros((self.nVar, mesh.nx + 2, mesh.ny + 2))
        self.VarOld = capi.pinned_zeros((self.nVar, mesh.nx + 2, mesh.ny + 2))
        self.residual = capi.pinned_zeros(self.nVar)
        self.Ff = capi.pinned_zeros((4, mesh.nx + 2, mesh.ny + 2))
        self.residual_history = {'u': [], 'v': [], 'p': []}
        self.last_sweeps = np.zeros(3, dtype=np.int64)
        self.total_sweeps = np.zeros(3, dtype=np.int64)
        self.converged, self.last_rms = False, np.zeros(3)      # verdict and rms triplet of the last solve()
        self._handle = capi.Handle(self._params())
        self._initialize_fields()

    # ---- parameter marshalling ---------------------------------------------------------------
    def _get_bc_arrays(self, k: int):
        """LDC.py:351-375."""
        return get_bc_arrays(self.bc, k)

    def _params(self) -> capi.Params:
        return make_params(self.mesh, self.fluid, self.settings, self.bc, getattr(self, 'case_type', None),
                           self.step_height, self.h, self.Ub, self.relaxed, self.device, self.max_ctas)

    def _sync_params(self):
        self._handle.set_params(self._params())

    def _crit(self):
        c = self.settings.convergence_criteria
        return (c['u'], c['v'], c['p'])

    # ---- reference methods -------------------------------------------------------------------
    def _initialize_fields(self):
        """LDC.py:377-389."""
        self._sync_params()
        self._handle.initialize_fields(True)
        self._handle.download(self.Var, self.VarOld, self.Ff)

    def _apply_bc_wrapper(self, k: int):
        """LDC.py:391-394 / BFS.py:564-569 (inlet override included when case_type == 'BFS')."""
        self._sync_params()
        self._handle.upload(Var=self.Var)
        self._handle.k_apply_bc(k)
        self._handle.download(Var=self.Var)

    def _apply_bfs_inlet(self, k: int):
        """BFS.py:524-562 (the left-boundary wall/parabolic-inlet override) on the device."""
        if getattr(self, 'case_type', None) != 'BFS' or k not in (0, 1):
            return
        self._sync_params()
        self._handle.upload(Var=self.Var)
        self._handle.k_apply_bfs_inlet(k)
        self._handle.download(Var=self.Var)

    def _implicit_solve(self):
        """LDC.py:432-467 / BFS.py:622-673: one outer iteration on the device."""
        self._sync_params()
        self._handle.upload(self.Var, self.VarOld, self.Ff)
        self._handle.k_implicit_solve()
        self._handle.download(self.Var, None, self.Ff, self.residual)
        st = self._handle.status()
        self.last_sweeps = st['last_sweeps']

    def _convergence_check(self, print_residuals: bool = False) -> Tuple[bool, np.ndarray]:
        """LDC.py:469-501 (host scalars; the copy is the only array work)."""
        rms = np.zeros(self.nVar)
        for k in range(self.nVar):
            rms[k] = np.sqrt(self.residual[k] / (self.mesh.nx * self.mesh.ny))
            rms[k] = rms[k] / self.settings.dt
            if print_residuals:
                print(f"\t{rms[k]:.6e}", end="")
        if print_residuals:
            print()
        if np.isnan(rms).any() or np.isinf(rms).any():
            print("\n❌ ERROR: NaN or Inf detected in residuals!")
            print(f"   U-residual: {rms[0]:.6e}, V-residual: {rms[1]:.6e}, P-residual: {rms[2]:.6e}")
            raise ValueError("Solver failed: NaN/Inf in residuals")
        c = self.settings.convergence_criteria
        converged = not (rms[0] > c['u'] or rms[1] > c['v'] or rms[2] > c['p'])
        if not converged:
            self._handle.upload(Var=self.Var)
            self._handle.k_copy_new_to_old()
            self._handle.download(VarOld=self.VarOld)
        return converged, rms

    def solve(self, output_base_name: str = "output", verbose: bool = True, save: bool = True):
        """LDC.py:396-430.  Iterates on the GPU; the host only looks at the device's verdict every
        <= 100 iterations (the reference's own print/history cadence)."""
        start_time = time.time()
        if verbose:
            print(f"Starting simulation with Re={self.fluid.Re}, mesh={self.mesh.nx}x{self.mesh.ny}")
            print(f"Time step: {self.settings.dt}, Scheme: {self.settings.scheme}")
            print("\nIteration\tU-RMS\t\tV-RMS\t\tP-RMS")
            print("-" * 60)
        crit, max_it = self._crit(), int(self.settings.max_iterations)
        if self._fits_one_cta():
            count = self._solve_resident(crit, max_it, verbose)
            return self._finish_solve(count, start_time, output_base_name, verbose, save)
        self._sync_params()
        H = self._handle
        H.upload(self.Var, self.VarOld, self.Ff)
        H.reset_counters()
        count, converged = 0, False
        try:
            while not converged and count < max_it:
                chunk = min(100 - count % 100, max_it - count)
                H.step(chunk, crit)
                st = H.status()                      # raises ValueError on NaN/Inf like LDC.py:487
                count, converged = st['iterations'], st['converged']
                if count % 100 == 0 and count > 0:
                    if verbose:
                        print(f"{count}\t{st['rms'][0]:.6e}\t{st['rms'][1]:.6e}\t{st['rms'][2]:.6e}")
                    for k, n in enumerate('uvp'):
                        self.residual_history[n].append(st['rms'][k])
                self.last_sweeps, self.total_sweeps = st['last_sweeps'], st['total_sweeps']
                self.converged, self.last_rms = bool(converged), np.array(st['rms'], dtype=np.float64)
        finally:
            H.download(self.Var, self.VarOld, self.Ff, self.residual)
        return self._finish_solve(count, start_time, output_base_name, verbose, save)

    def _finish_solve(self, count, start_time, output_base_name, verbose, save):
        end_time = time.time()
        if verbose:
            print(f"\n\nSimulation completed in {end_time - start_time:.2f} seconds")
            print(f"Total iterations: {count}")
        if save:
            self._save_results(output_base_name)
        return count, end_time - start_time

    # ---- grids small enough for one CTA (the 10x10 coarse stage): the whole solve() in one launch ----
    resident_solve = True        # set False to force the whole-GPU kernels on a small grid

    def _fits_one_cta(self) -> bool:
        order = getattr(self.settings, 'sweep_order', 'GS_LEX')
        if not self.resident_solve or (order.upper() if isinstance(order, str) else order) not in ('GS_LEX', capi.ORDERS['GS_LEX']):
            return False
        return fits_one_cta(self.mesh.nx, self.mesh.ny, getattr(self, 'device', 0))

    def _solve_resident(self, crit, max_it, verbose) -> int:
        """solve() through srcfd_coarse_solve_batch (one case): same iterates, same history cadence; the
        every-100-iterations lines are printed once the launch is back."""
        st = (self.Var[None], self.VarOld[None], self.Ff[None])
        r = capi.coarse_solve_batch([self._params()], max_it, crit, hist_cap=max_it // 100 + 1, state=st)
        self.residual[:] = r['residual'][0]
        self.converged, self.last_rms = bool(r['converged'][0]), r['rms'][0].copy()
        self.last_sweeps, self.total_sweeps = r['last_sweeps'][0], r['total_sweeps'][0]
        for i, row in enumerate(r['hist'][0]):
            if verbose:
                print(f"{100 * (i + 1)}\t{row[0]:.6e}\t{row[1]:.6e}\t{row[2]:.6e}")
            for k, n in enumerate('uvp'):
                self.residual_history[n].append(row[k])
        if r['nan'][0]:
            rms = r['rms'][0]
            print("\n❌ ERROR: NaN or Inf detected in residuals!")
            print(f"   U-residual: {rms[0]:.6e}, V-residual: {rms[1]:.6e}, P-residual: {rms[2]:.6e}")
            raise ValueError("Solver failed: NaN/Inf in residuals")
        return int(r['iterations'][0])

    # ---- output (reference layout; plots need matplotlib and are skipped when it is absent) ---
    def _group_name(self):
        return f"Re{self.fluid.Re}_mesh{self.mesh.nx}x{self.mesh.ny}"     # LDC.py:511 (hazard H10)

    def _save_results(self, output_base_name: str):
        output_dir = os.path.dirname(output_base_name)
        if output_dir and not os.path.exists(output_dir):
            os.makedirs(output_dir)
        self._save_results_hdf5(f"{output_base_name}.h5", self._group_name())

    def _save_results_hdf5(self, filename: str, group_name: str):
        """Same group/dataset/attribute layout as LDC.py:517-544 / BFS.py:734-752."""
        from . import h5lite
        root = h5lite.read_h5(filename) if os.path.exists(filename) else h5lite.Group()
        grp = h5lite.Group()
        grp.attrs = {"case_name": self.case_name, "reynolds_number": self.fluid.Re, "nx": self.mesh.nx, "ny": self.mesh.ny}
        if getattr(self, 'case_type', None) == 'BFS':        # BFS.py:734-741 adds the domain and the step
            grp.attrs.update({"lx": self.mesh.lx, "ly": self.mesh.ly, "step_height": self.step_height})
        grp.attrs["total_points"] = self.mesh.nx * self.mesh.ny
        x = np.linspace(0, self.mesh.lx, self.mesh.nx)
        y = np.linspace(0, self.mesh.ly, self.mesh.ny)
        X, Y = np.meshgrid(x, y)
        grp["x"], grp["y"] = X.flatten(), Y.flatten()
        for k, n in enumerate("uvp"):
            grp[n] = self.Var[k, 1:-1, 1:-1].T.flatten()
        root[group_name] = grp
        h5lite.write_h5(filename, root)

    def _save_centerline_data(self, filename: str):
        """"bfs code given by sir.py":359-384 text layout."""
        u_vertical = self.Var[0, self.mesh.nx // 2, 1:-1]
        v_horizontal = self.Var[1, 1:-1, self.mesh.ny // 2]
        y = np.linspace(0, self.mesh.ly, self.mesh.ny)
        x = np.linspace(0, self.mesh.lx, self.mesh.nx)
        with open(filename, 'w') as f:
            f.write(f"# Reynolds number: {self.fluid.Re}\n")
            f.write(f"# Mesh: {self.mesh.nx}x{self.mesh.ny}\n")
            f.write("# Centerline data\n")
            f.write("# y\tu(x=0.5)\tx\tv(y=0.5)\n")
            for i in range(max(len(y), len(x))):
                f.write(f"{y[i]:.6f}\t{u_vertical[i]:.6f}\t" if i < len(y) else "\t\t")
                if i < len(x):
                    f.write(f"{x[i]:.6f}\t{v_horizontal[i]:.6f}")
                f.write("\n")


class BFSCFDSolver(CFDSolver):
    """BFS.py:471-496 signature: step_height/h/Ub default to 1/2/1 and the case is always 'BFS'."""

    case_name = "backward facing step"

    def __init__(self, mesh, fluid, solver_settings, bc, step_height: float = 1.0, h: float = 2.0,
                 Ub: float = 1.0, **kw):
        super().__init__(mesh, fluid, solver_settings, bc, step_height, h, Ub, **kw)
