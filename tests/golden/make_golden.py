"""Regenerate the committed golden fixtures from the reference tree.

Run in the build container only (needs /root/reference and numba):

    NUMBA_NUM_THREADS=1 python tests/golden/make_golden.py

Writes, next to this file:
  reference_outputs.npz   numbers the reference itself committed: the table in
                          outputs/bfs_Re400_centerline.dat and the u/v/p vectors of every
                          coarse 10x10 result .h5 under outputs/ (read with h5lite).
  numba_kernels.npz       seeded random inputs and the outputs of the reference's own numba
                          kernels (1 thread => deterministic lexicographic Gauss-Seidel).
  numba_solves.npz        final Var/VarOld/Ff of short CFDSolver.solve() runs of the reference.
  encoder10_multiBC.h5, stats_10to400_multiBC.txt
                          byte copies of the committed Keras encoder weights and
                          standardisation statistics (data artefacts the SR warm start loads;
                          the decoder weights are absent from the reference tree).
"""
import glob
import os
import shutil
import sys

os.environ.setdefault("NUMBA_NUM_THREADS", "1")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "sr-for-cfd_b200"))

import numpy as np  # noqa: E402

from oracle import ref_loader as RL  # noqa: E402
from srcfd import h5lite  # noqa: E402

REF = RL.REF_ROOT


def reference_outputs():
    out = {"centerline_dat": np.loadtxt(os.path.join(REF, "outputs", "bfs_Re400_centerline.dat"))}
    names = []
    for path in sorted(glob.glob(os.path.join(REF, "outputs", "*", "*coarse*10x10*.h5"))):
        g = h5lite.read_h5(path)
        gname = list(g.keys())[0]
        grp = g[gname]
        tag = os.path.basename(os.path.dirname(path)).split(" ")[0] + "|" + os.path.basename(path) + "|" + gname
        names.append(tag)
        out[f"coarse_{len(names) - 1}"] = np.stack([grp[k].data for k in ("u", "v", "p", "x", "y")])
    out["coarse_names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "reference_outputs.npz"), **out)
    print("reference_outputs.npz:", len(names), "coarse files")


def numba_kernels():
    ldc, bfs = RL.load("LDC"), RL.load("BFS")
    rng = np.random.default_rng(20251018)
    Nx, Ny = 13, 9
    lx, ly = 1.0, 0.7
    dx, dy = lx / Nx, ly / Ny
    volp, dt, nu, rho = dx * dy, 1e-3, 1.0 / 100.0, 1.0
    Var = rng.uniform(-1, 1, (3, Nx + 2, Ny + 2))
    VarOld = Var + 0.01 * rng.uniform(-1, 1, Var.shape)
    Ff = 0.05 * rng.uniform(-1, 1, (4, Nx + 2, Ny + 2))
    out = dict(Nx=Nx, Ny=Ny, dx=dx, dy=dy, volp=volp, dt=dt, nu=nu, rho=rho, Var=Var, VarOld=VarOld, Ff=Ff)

    A = Ff.copy(); ldc.linear_interpolation(Var, A, Nx, Ny, dx, dy); out["linear_interpolation"] = A
    A = Ff.copy(); ldc.update_flux(Var, A, dt, rho, Nx, Ny, dx, dy); out["update_flux"] = A
    bt = np.array([[0, 1, 0, 1], [1, 0, 0, 0], [1, 0, 1, 1]], dtype=np.int32)
    bv = np.array([[0.3, -0.2, 1.0, 0.5], [0.0, 0.1, 0.0, -0.4], [0.0, 0.25, 0.0, 0.0]])
    out["bc_types"], out["bc_values"] = bt, bv
    A = Var.copy()
    for k in range(3):
        ldc.apply_bc_configured(A, k, Nx, Ny, bt[k], bv[k])
    out["apply_bc"] = A
    A = Var.copy(); bfs.under_relax_field(A, VarOld, 1, Nx, Ny, 0.5); out["under_relax_k1_a05"] = A
    A = Var.copy(); r = np.zeros(3); ldc.correct_velocity(A, VarOld, dt, rho, Nx, Ny, dx, dy, r)
    out["correct_velocity"], out["correct_velocity_residual"] = A, r
    A = Var.copy(); ldc.solve_pressure(A, Ff, Nx, Ny, dx, dy, dt, rho, volp); out["solve_pressure"] = A
    for k in (0, 1):
        A = Var.copy(); ldc.solve_momentum_upwind(A, VarOld, Ff, k, Nx, Ny, dx, dy, dt, nu, volp)
        out[f"solve_momentum_upwind_k{k}"] = A
        A = Var.copy(); ldc.solve_momentum_quick(A, VarOld, Ff, k, Nx, Ny, dx, dy, dt, nu, volp)
        out[f"solve_momentum_quick_k{k}"] = A
    # BFS inlet override through the reference class method (bfs_ml_accelerated.py:524-562)
    mesh = bfs.MeshParameters(nx=Nx, ny=Ny, lx=10.0, ly=3.0)
    s = bfs.CFDSolver(mesh, bfs.FluidProperties(Re=400.0), bfs.SolverSettings(), bfs.BoundaryConditions(),
                      step_height=1.0, h=2.0, Ub=1.0)
    s.Var[:] = Var
    s._apply_bfs_inlet(0); s._apply_bfs_inlet(1)
    out["bfs_inlet"] = s.Var.copy()
    np.savez_compressed(os.path.join(HERE, "numba_kernels.npz"), **out)
    print("numba_kernels.npz written")


def numba_solves():
    ldc, bfs = RL.load("LDC"), RL.load("BFS")
    out = {}
    s = ldc.CFDSolver(ldc.MeshParameters(nx=24, ny=20, lx=1.0, ly=1.0), ldc.FluidProperties(Re=100.0),
                      ldc.SolverSettings(dt=1e-3, scheme="QUICK", max_iterations=30), ldc.BoundaryConditions())
    s.solve("x", verbose=False)
    out["ldc_24x20_quick_30_Var"], out["ldc_24x20_quick_30_VarOld"], out["ldc_24x20_quick_30_Ff"] = s.Var, s.VarOld, s.Ff

    s = ldc.CFDSolver(ldc.MeshParameters(nx=16, ny=16, lx=1.0, ly=1.0), ldc.FluidProperties(Re=400.0),
                      ldc.SolverSettings(dt=1e-3, scheme="UPWIND", max_iterations=40), ldc.BoundaryConditions())
    s.solve("x", verbose=False)
    out["ldc_16x16_upwind_40_Var"] = s.Var

    def bfs_bc():
        bc = bfs.BoundaryConditions()
        bc.u_boundaries["left"] = bfs.BoundaryCondition("dirichlet", 0.0)
        bc.u_boundaries["right"] = bfs.BoundaryCondition("neumann", 0.0)
        bc.v_boundaries["right"] = bfs.BoundaryCondition("neumann", 0.0)
        bc.p_boundaries["right"] = bfs.BoundaryCondition("dirichlet", 0.0)
        return bc

    mesh = bfs.MeshParameters(nx=20, ny=16, lx=10.0, ly=3.0)
    s = bfs.CFDSolver(mesh, bfs.FluidProperties(Re=400.0),
                      bfs.SolverSettings(dt=2e-3, scheme="UPWIND", max_iterations=250), bfs_bc(),
                      step_height=1.0, h=2.0, Ub=1.0)
    s.solve("x", verbose=False)
    out["bfs_20x16_upwind_250_Var"], out["bfs_20x16_upwind_250_Ff"] = s.Var, s.Ff
    out["bfs_20x16_upwind_250_hist"] = np.array([s.residual_history[k] for k in "uvp"]).T
    s = bfs.CFDSolver(mesh, bfs.FluidProperties(Re=400.0),
                      bfs.SolverSettings(dt=2e-3, scheme="QUICK", max_iterations=40), bfs_bc(),
                      step_height=1.0, h=2.0, Ub=1.0)
    s.solve("x", verbose=False)
    out["bfs_20x16_quick_40_Var"] = s.Var
    # warm start from a given field (PyCFD_ML_accelerated.py:928-948) then 5 iterations
    rng = np.random.default_rng(7)
    fields = {k: 0.1 * rng.standard_normal((20, 24)).astype(np.float32) for k in "uvp"}
    s = ldc.CFDSolver(ldc.MeshParameters(nx=24, ny=20, lx=1.0, ly=1.0), ldc.FluidProperties(Re=100.0),
                      ldc.SolverSettings(dt=1e-3, scheme="QUICK", max_iterations=5), ldc.BoundaryConditions())
    for k, n in enumerate("uvp"):
        s.Var[k, 1:-1, 1:-1] = fields[n].T
    for k in range(3):
        s._apply_bc_wrapper(k)
    ldc.copy_new_to_old(s.Var, s.VarOld, 3, 24, 20)
    ldc.linear_interpolation(s.Var, s.Ff, 24, 20, s.mesh.dx, s.mesh.dy)
    s.solve("x", verbose=False)
    out["warm_fields"] = np.stack([fields[k] for k in "uvp"])
    out["warm_ldc_24x20_quick_5_Var"] = s.Var
    np.savez_compressed(os.path.join(HERE, "numba_solves.npz"), **out)
    print("numba_solves.npz written")


def data_files():
    shutil.copyfile(os.path.join(REF, "vanilla_encoder10_to_400_swish_trained_upto_700_multiBC.h5"),
                    os.path.join(HERE, "encoder10_multiBC.h5"))
    shutil.copyfile(os.path.join(REF, "standardization_stats_10to400_swish_trained_upto_700_multiBC.txt"),
                    os.path.join(HERE, "stats_10to400_multiBC.txt"))
    # output-format fixtures: one committed coarse result file (h5py layout) and the centerline text file
    src = sorted(glob.glob(os.path.join(REF, "outputs", "*", "bfs_coarse_Re400_10x10_100000_coarse_iterations.h5")))[0]
    shutil.copyfile(src, os.path.join(HERE, "ref_bfs_coarse_Re400_10x10.h5"))
    shutil.copyfile(os.path.join(REF, "outputs", "bfs_Re400_centerline.dat"), os.path.join(HERE, "ref_bfs_Re400_centerline.dat"))
    print("encoder + stats + output-format fixtures copied")


if __name__ == "__main__":
    if not RL.available():
        sys.exit("reference tree not mounted; golden fixtures are regenerated only in the build container")
    reference_outputs()
    numba_kernels()
    numba_solves()
    data_files()
