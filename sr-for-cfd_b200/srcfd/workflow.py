"""Workflow functions of the two reference scripts, same names / arguments / return values:
PyCFD_ML_accelerated.py:696-1265 (LDC) and bfs_ml_accelerated.py:893-1604 (BFS).

coarse solve -> SR autoencoder (encoder_10 + decoder_400) -> inject as fine-grid initial guess ->
iterate.  Solver work runs on libsrcfd (CUDA); the SR networks run in srcfd.sr (CUDA).  Plotting is
not part of the hot path and is left out; result files keep the reference's HDF5 layout.
"""
from __future__ import annotations

import os
from datetime import datetime
from typing import Dict, Optional

import numpy as np

from . import kernels as K
from . import solver as S


def create_timestamped_output_dir(base_dir: str = "outputs") -> str:
    """PyCFD_ML_accelerated.py:21-34."""
    output_dir = os.path.join(base_dir, datetime.now().strftime("%d-%m-%Y-%H-%M-%S"))
    os.makedirs(output_dir, exist_ok=True)
    return output_dir


def standardize_with_stats(arr, mean, std):
    """PyCFD_ML_accelerated.py:665-668."""
    std = 1e-8 if std == 0 else std
    return (arr - mean) / std


def inverse_standardize(arr, mean, std):
    """PyCFD_ML_accelerated.py:671-673."""
    return arr * std + mean


def _say(verbose, *a):
    if verbose:
        print(*a)


def load_stats(stats_file: str, lr_dim: int, hr_dim: int):
    """Stats-file parser of PyCFD_ML_accelerated.py:786-825 (raises FileNotFoundError / KeyError alike)."""
    stats = {}
    with open(stats_file, "r") as f:
        for line in f:
            if line.strip().startswith('#') or not line.strip():
                continue
            parts = line.strip().split()
            if len(parts) == 2:
                stats[parts[0]] = float(parts[1])
    stats_lr = {c: (stats[f'mean{lr_dim}_{c}'], stats[f'std{lr_dim}_{c}']) for c in 'uvp'}
    stats_hr = {c: (stats[f'mean{hr_dim}_{c}'], stats[f'std{hr_dim}_{c}']) for c in 'uvp'}
    return stats_lr, stats_hr


class SuperResolutionAE:
    """PyCFD_ML_accelerated.py:676-689: encoder_lr followed by decoder_hr; `.predict` like Keras."""

    def __init__(self, encoder_lr, decoder_hr, **kwargs):
        self.encoder_lr = encoder_lr
        self.decoder_hr = decoder_hr

    def call(self, inputs, training=False):
        from . import sr
        if isinstance(self.encoder_lr, sr.Encoder) and isinstance(self.decoder_hr, sr.Decoder):
            return sr.predict(self.encoder_lr, self.decoder_hr, inputs)      # latents stay on the device
        return self.decoder_hr(self.encoder_lr(inputs))

    __call__ = call

    def predict(self, x, verbose=0):
        return self.call(np.asarray(x, dtype=np.float32))


class _Workflow:
    """One instance per mirrored script; `bfs` switches defaults and the BFS-only arguments."""

    def __init__(self, bfs: bool):
        self.bfs = bfs
        self.verbose = True

    # expose the helpers under the reference's names
    create_timestamped_output_dir = staticmethod(create_timestamped_output_dir)
    standardize_with_stats = staticmethod(standardize_with_stats)
    inverse_standardize = staticmethod(inverse_standardize)
    SuperResolutionAE = SuperResolutionAE

    # ---- construction helpers ----------------------------------------------------------------
    def _default_bc(self):
        if not self.bfs:
            return S.BoundaryConditions()
        bc = S.BFSBoundaryConditions()                       # bfs_ml_accelerated.py:948-953
        bc.u_boundaries['right'] = S.BoundaryCondition('neumann', 0.0)
        bc.v_boundaries['right'] = S.BoundaryCondition('neumann', 0.0)
        bc.p_boundaries['right'] = S.BoundaryCondition('dirichlet', 0.0)
        return bc

    def _make_solver(self, Re, nx, ny, dt, scheme, convergence_criteria, max_iterations, bc, step_height, h, Ub,
                     lx, ly, relaxation_factors, **solver_kw):
        mesh = S.MeshParameters(nx=nx, ny=ny, lx=lx, ly=ly)
        fluid = S.FluidProperties(Re=Re, rho=1.0)
        if convergence_criteria is None:
            convergence_criteria = {'u': 1e-6, 'v': 1e-6, 'p': 1e-6, 'continuity': 1e-6}
        if bc is None:
            bc = self._default_bc()
        if self.bfs:
            st = S.BFSSolverSettings(dt=dt, scheme=scheme, max_iterations=max_iterations,
                                     convergence_criteria=convergence_criteria, relaxation_factors=relaxation_factors)
            return S.BFSCFDSolver(mesh, fluid, st, bc, step_height=step_height, h=h, Ub=Ub, **solver_kw)
        st = S.SolverSettings(dt=dt, scheme=scheme, max_iterations=max_iterations,
                              convergence_criteria=convergence_criteria)
        return S.CFDSolver(mesh, fluid, st, bc, **solver_kw)

    def _defaults(self, dt, scheme, lx, ly):
        if self.bfs:
            return (0.002 if dt is None else dt, 'UPWIND' if scheme is None else scheme,
                    10.0 if lx is None else lx, 3.0 if ly is None else ly)
        return (0.001 if dt is None else dt, 'QUICK' if scheme is None else scheme, 1.0, 1.0)

    # ---- step 1: coarse solve (LDC.py:696-761 / BFS.py:893-977) -------------------------------
    def run_coarse_simulation(self, Re, lr_dim=10, dt=None, scheme=None, convergence_criteria=None,
                              max_iterations=100000, output_dir=None, bc=None, step_height=1.0, h=2.0, Ub=1.0,
                              lx=None, ly=None, relaxation_factors=None, save=True):
        dt, scheme, lx, ly = self._defaults(dt, scheme, lx, ly)
        _say(self.verbose, f"\n{'=' * 70}\nSTEP 1: Running Coarse Simulation (Re={Re}, mesh={lr_dim}x{lr_dim})\n{'=' * 70}")
        solver = self._make_solver(Re, lr_dim, lr_dim, dt, scheme, convergence_criteria, max_iterations, bc,
                                   step_height, h, Ub, lx, ly, relaxation_factors)
        if output_dir is None and save:
            output_dir = create_timestamped_output_dir()
        prefix = "bfs_coarse" if self.bfs else "coarse"
        output_name = os.path.join(output_dir or ".", f"{prefix}_Re{Re}_{lr_dim}x{lr_dim}_{max_iterations}_coarse_iterations")
        iterations, time_elapsed = solver.solve(output_name, verbose=self.verbose, save=save)
        _say(self.verbose, f"Coarse simulation completed in {iterations} iterations ({time_elapsed:.2f} seconds)")
        return {n: solver.Var[k, 1:-1, 1:-1].T.copy() for k, n in enumerate('uvp')}

    # ---- step 2: SR (LDC.py:764-879 / BFS.py:979-1137) ------------------------------------------
    def ml_super_resolution(self, coarse_fields, lr_dim, hr_dim, stats_file, encoder_file, decoder_file,
                            use_aspect_ratio_correction=False, lx=1.0, ly=1.0, use_adaptive_normalization=None,
                            blend_factor=0.3):
        return self.ml_super_resolution_batch([coarse_fields], lr_dim, hr_dim, stats_file, encoder_file, decoder_file,
                                              use_aspect_ratio_correction, lx, ly, use_adaptive_normalization,
                                              blend_factor)[0]

    def ml_super_resolution_batch(self, coarse_list, lr_dim, hr_dim, stats_file, encoder_file, decoder_file,
                                  use_aspect_ratio_correction=False, lx=1.0, ly=1.0, use_adaptive_normalization=None,
                                  blend_factor=0.3):
        """ml_super_resolution (LDC.py:764-879 / BFS.py:979-1137) for a LIST of cases: the u, v, p fields of every case go
        through ONE library call (3 x cases fields; the reference runs three batch-1 predicts per case) in which the
        statistics blend, the standardisation, both networks, the inverse standardisation and the NaN/Inf guard all run
        on the device (srcfd_sr_super_resolve).  Only the aspect-ratio spline resampling stays on the host (scipy)."""
        from . import sr
        if use_adaptive_normalization is None:
            use_adaptive_normalization = self.bfs        # BFS.py:984 defaults it on, LDC.py has none
        reshape = self.bfs and use_aspect_ratio_correction and (lx != ly)
        stats_lr, stats_hr = load_stats(stats_file, lr_dim, hr_dim)
        enc, dec = sr.load_model(encoder_file), sr.load_model(decoder_file)
        adaptive = bool(self.bfs and use_adaptive_normalization)
        xs, st = [], []
        for coarse_fields in coarse_list:
            f = reshape_rectangular_to_square(coarse_fields, lr_dim, lr_dim, lx, ly) if reshape else coarse_fields
            for c in 'uvp':
                xs.append(np.asarray(f[c]).astype(np.float32))
                st.append([stats_lr[c][0], stats_lr[c][1], stats_hr[c][0], stats_hr[c][1]])
        if isinstance(enc, sr.Encoder) and isinstance(dec, sr.Decoder):
            pred = sr.super_resolve(enc, dec, np.stack(xs), np.asarray(st, dtype=np.float64), adaptive, blend_factor)
        else:                                            # foreign model objects: the reference's statement sequence on the host
            model = SuperResolutionAE(enc, dec)
            pred = np.empty((len(xs), hr_dim, hr_dim), dtype=np.float32)
            for i, (x_lr_raw, (mean_lr, std_lr, mean_hr, std_hr)) in enumerate(zip(xs, st)):
                if adaptive:                             # BFS.py:1090-1100
                    input_mean, input_std = np.mean(x_lr_raw), np.std(x_lr_raw)
                    mean_lr = (1 - blend_factor) * mean_lr + blend_factor * input_mean
                    std_lr = (1 - blend_factor) * std_lr + blend_factor * max(input_std, 1e-8)
                x = np.expand_dims(standardize_with_stats(x_lr_raw, mean_lr, std_lr), axis=(0, -1))
                p = inverse_standardize(model.predict(x, verbose=0)[0, ..., 0], mean_hr, std_hr)
                if np.isnan(p).any() or np.isinf(p).any():   # LDC.py:869-876
                    p = np.nan_to_num(p, nan=0.0, posinf=0.0, neginf=0.0)
                pred[i] = p
        out = []
        for n in range(len(coarse_list)):
            hr_fields = {c: pred[3 * n + k] for k, c in enumerate('uvp')}
            if reshape:
                hr_fields = reshape_square_to_rectangular(hr_fields, hr_dim, hr_dim, lx, ly)
            out.append(hr_fields)
        return out

    # ---- step 3: fine solve from the SR field (LDC.py:882-959 / BFS.py:1140-1234) ---------------
    def run_fine_simulation_with_ml_init(self, Re, nx, ny, ml_initial_fields, dt=None, scheme=None,
                                         convergence_criteria=None, max_iterations=100000, output_name=None,
                                         bc=None, step_height=1.0, h=2.0, Ub=1.0, lx=None, ly=None,
                                         relaxation_factors=None, save=True):
        dt, scheme, lx, ly = self._defaults(dt, scheme, lx, ly)
        if output_name is None:
            output_name = "bfs_accelerated" if self.bfs else "cavity_accelerated"
        solver = self._make_solver(Re, nx, ny, dt, scheme, convergence_criteria, max_iterations, bc, step_height,
                                   h, Ub, lx, ly, relaxation_factors)
        # same statement sequence as the reference (it pokes the host arrays, then calls two kernels)
        solver.Var[0, 1:-1, 1:-1] = ml_initial_fields['u'].T
        solver.Var[1, 1:-1, 1:-1] = ml_initial_fields['v'].T
        solver.Var[2, 1:-1, 1:-1] = ml_initial_fields['p'].T
        for k in range(solver.nVar):
            solver._apply_bc_wrapper(k)
        K.copy_new_to_old(solver.Var, solver.VarOld, solver.nVar, solver.mesh.nx, solver.mesh.ny)
        K.linear_interpolation(solver.Var, solver.Ff, solver.mesh.nx, solver.mesh.ny, solver.mesh.dx, solver.mesh.dy)
        if not output_name.endswith("_accelerated"):
            output_name = f"{output_name}_accelerated"
        iterations, time_elapsed = solver.solve(output_name, verbose=self.verbose, save=save)
        return solver, iterations, time_elapsed

    def run_normal_simulation(self, Re, nx, ny, dt=None, scheme=None, convergence_criteria=None,
                              max_iterations=100000, output_name=None, bc=None, step_height=1.0, h=2.0, Ub=1.0,
                              lx=None, ly=None, relaxation_factors=None, save=True):
        """LDC.py:1126-1180 / BFS.py:1237-1307."""
        dt, scheme, lx, ly = self._defaults(dt, scheme, lx, ly)
        if output_name is None:
            output_name = "bfs_normal" if self.bfs else "cavity_normal"
        solver = self._make_solver(Re, nx, ny, dt, scheme, convergence_criteria, max_iterations, bc, step_height,
                                   h, Ub, lx, ly, relaxation_factors)
        if not output_name.endswith("_normal"):
            output_name = f"{output_name}_normal"
        iterations, time_elapsed = solver.solve(output_name, verbose=self.verbose, save=save)
        return solver, iterations, time_elapsed

    def generate_coarse_mesh_solution(self, Re, lr_dim=10, dt=None, scheme=None, convergence_criteria=None,
                                      max_iterations_coarse=100000, output_dir=None, bc=None, **bfs_kw):
        """LDC.py:966-1021 / BFS.py:1310-1381: returns (coarse_fields, output_dir)."""
        if output_dir is None:
            output_dir = create_timestamped_output_dir()
        fields = self.run_coarse_simulation(Re=Re, lr_dim=lr_dim, dt=dt, scheme=scheme,
                                            convergence_criteria=convergence_criteria,
                                            max_iterations=max_iterations_coarse, output_dir=output_dir, bc=bc,
                                            **bfs_kw)
        return fields, output_dir

    def run_ml_accelerated_fine_simulation(self, coarse_fields, Re, nx, ny, lr_dim=10, dt=None, scheme=None,
                                           convergence_criteria=None, max_iterations_fine=100000, output_name=None,
                                           stats_file=None, encoder_file=None, decoder_file=None, bc=None,
                                           step_height=1.0, h=2.0, Ub=1.0, lx=None, ly=None, relaxation_factors=None,
                                           use_aspect_ratio_correction=False, use_adaptive_normalization=None,
                                           blend_factor=0.3, save=True):
        """LDC.py:1024-1119 / BFS.py:1384-1517."""
        dt, scheme, lx, ly = self._defaults(dt, scheme, lx, ly)
        if stats_file is None:
            stats_file = f"standardization_stats_{lr_dim}to{nx}.txt"
        if encoder_file is None:
            encoder_file = f"vanilla_encoder{lr_dim}_to_{nx}.h5"
        if decoder_file is None:
            decoder_file = f"vanilla_decoder{nx}_from_{lr_dim}.h5"
        if output_name is None:
            output_name = f"{'bfs' if self.bfs else 'cavity'}_Re{Re}_{nx}x{ny}"
        for fname, desc in [(stats_file, "Stats file"), (encoder_file, "Encoder model"), (decoder_file, "Decoder model")]:
            if not (isinstance(fname, str) and os.path.exists(fname)) and isinstance(fname, str):
                raise FileNotFoundError(f"{desc} not found: {fname}")
        hr_fields = self.ml_super_resolution(coarse_fields=coarse_fields, lr_dim=lr_dim, hr_dim=nx,
                                             stats_file=stats_file, encoder_file=encoder_file,
                                             decoder_file=decoder_file,
                                             use_aspect_ratio_correction=use_aspect_ratio_correction, lx=lx, ly=ly,
                                             use_adaptive_normalization=use_adaptive_normalization,
                                             blend_factor=blend_factor)
        return self.run_fine_simulation_with_ml_init(Re=Re, nx=nx, ny=ny, ml_initial_fields=hr_fields, dt=dt,
                                                     scheme=scheme, convergence_criteria=convergence_criteria,
                                                     max_iterations=max_iterations_fine, output_name=output_name,
                                                     bc=bc, step_height=step_height, h=h, Ub=Ub, lx=lx, ly=ly,
                                                     relaxation_factors=relaxation_factors, save=save)

    @staticmethod
    def extract_centerlines(solver, nx: int, ny: int, lx: float = 1.0, ly: float = 1.0):
        """LDC.py:1236-1270 / BFS.py:1569-1603."""
        x = np.linspace(0, lx, nx)
        y = np.linspace(0, ly, ny)
        u_field = solver.Var[0, 1:-1, 1:-1].T.copy()
        v_field = solver.Var[1, 1:-1, 1:-1].T.copy()
        return {'u_vertical': {'y': y, 'values': u_field[:, nx // 2]},
                'v_horizontal': {'x': x, 'values': v_field[ny // 2, :]}}


def reshape_rectangular_to_square(fields: Dict[str, np.ndarray], nx_rect: int, ny_rect: int, lx: float, ly: float):
    """bfs_ml_accelerated.py:59-101: cubic-spline resample of the (ny,nx) fields onto a square frame.
    Host-side pre-step (scipy), outside the CUDA hot path."""
    from scipy import interpolate
    x_rect, y_rect = np.linspace(0, lx, nx_rect), np.linspace(0, ly, ny_rect)
    L = max(lx, ly)
    x_sq, y_sq = np.linspace(0, L, nx_rect), np.linspace(0, L, nx_rect)
    return {c: interpolate.RectBivariateSpline(y_rect, x_rect, fields[c], kx=3, ky=3)(y_sq, x_sq) for c in 'uvp'}


def reshape_square_to_rectangular(fields: Dict[str, np.ndarray], nx_rect: int, ny_rect: int, lx: float, ly: float):
    """bfs_ml_accelerated.py:104-145."""
    from scipy import interpolate
    n_sq = fields['u'].shape[0]
    L = max(lx, ly)
    x_sq, y_sq = np.linspace(0, L, n_sq), np.linspace(0, L, n_sq)
    x_rect, y_rect = np.linspace(0, lx, nx_rect), np.linspace(0, ly, ny_rect)
    return {c: interpolate.RectBivariateSpline(y_sq, x_sq, fields[c], kx=3, ky=3)(y_rect, x_rect) for c in 'uvp'}


ldc_workflow = _Workflow(bfs=False)
bfs_workflow = _Workflow(bfs=True)
bfs_workflow.reshape_rectangular_to_square = reshape_rectangular_to_square
bfs_workflow.reshape_square_to_rectangular = reshape_square_to_rectangular
