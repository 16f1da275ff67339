"""Live check of the oracle against the UNMODIFIED reference (numba, 1 thread).

Runs only where /root/reference is mounted (the build container); skipped elsewhere.  The same
comparisons are frozen into tests/golden/*.npz so they also hold on the GPU box.
"""
import os
import subprocess
import sys

import pytest

from oracle import ref_loader as RL

pytestmark = pytest.mark.skipif(not RL.available(), reason="reference tree not mounted")

_SCRIPT = r"""
import os, sys
os.environ["NUMBA_NUM_THREADS"] = "1"
sys.path.insert(0, sys.argv[1])
import numpy as np
from oracle import oracle as O, ref_loader as RL
ldc, bfs = RL.load("LDC"), RL.load("BFS")
rng = np.random.default_rng(int(sys.argv[2]))
Nx, Ny = int(rng.integers(3, 20)), int(rng.integers(3, 20))
dx, dy = 1.3 / Nx, 0.9 / Ny
volp, dt, nu, rho = dx * dy, 2e-3, 1 / 250.0, 1.0
Var = rng.uniform(-1, 1, (3, Nx + 2, Ny + 2)); VarOld = Var + 0.01 * rng.uniform(-1, 1, Var.shape)
Ff = 0.05 * rng.uniform(-1, 1, (4, Nx + 2, Ny + 2))
def same(a, b, what): assert np.array_equal(a, b), what
A, B = Ff.copy(), Ff.copy(); ldc.linear_interpolation(Var, A, Nx, Ny, dx, dy); O.linear_interpolation(Var, B, Nx, Ny, dx, dy); same(A, B, "interp")
A, B = Ff.copy(), Ff.copy(); ldc.update_flux(Var, A, dt, rho, Nx, Ny, dx, dy); O.update_flux(Var, B, dt, rho, Nx, Ny, dx, dy); same(A, B, "flux")
A, B = Var.copy(), Var.copy(); ra, rb = np.zeros(3), np.zeros(3)
ldc.correct_velocity(A, VarOld, dt, rho, Nx, Ny, dx, dy, ra); O.correct_velocity(B, VarOld, dt, rho, Nx, Ny, dx, dy, rb); same(A, B, "correct"); same(ra, rb, "res")
A, B = Var.copy(), Var.copy(); ldc.solve_pressure(A, Ff, Nx, Ny, dx, dy, dt, rho, volp); O.solve_pressure(B, Ff, Nx, Ny, dx, dy, dt, rho, volp); same(A, B, "pressure")
for k in (0, 1):
    A, B = Var.copy(), Var.copy(); ldc.solve_momentum_upwind(A, VarOld, Ff, k, Nx, Ny, dx, dy, dt, nu, volp); O.solve_momentum_upwind(B, VarOld, Ff, k, Nx, Ny, dx, dy, dt, nu, volp); same(A, B, "upwind")
    A, B = Var.copy(), Var.copy(); ldc.solve_momentum_quick(A, VarOld, Ff, k, Nx, Ny, dx, dy, dt, nu, volp); O.solve_momentum_quick(B, VarOld, Ff, k, Nx, Ny, dx, dy, dt, nu, volp); same(A, B, "quick")
# composed: BFS class, a few outer iterations
mesh = bfs.MeshParameters(nx=Nx + 4, ny=Ny + 4, lx=10.0, ly=3.0)
bc = bfs.BoundaryConditions()
bc.u_boundaries["left"] = bfs.BoundaryCondition("dirichlet", 0.0)
for d in (bc.u_boundaries, bc.v_boundaries): d["right"] = bfs.BoundaryCondition("neumann", 0.0)
bc.p_boundaries["right"] = bfs.BoundaryCondition("dirichlet", 0.0)
s = bfs.CFDSolver(mesh, bfs.FluidProperties(Re=400.0), bfs.SolverSettings(dt=2e-3, scheme="UPWIND", max_iterations=25), bc)
s.solve("x", verbose=False)
o = O.OracleSolver(O.bfs_case(Nx + 4, Ny + 4)); o.solve(25)
same(s.Var, o.Var, "bfs solve"); same(s.Ff, o.Ff, "bfs Ff")
print("OK")
"""


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_oracle_matches_numba_reference(seed):
    # own interpreter: NUMBA_NUM_THREADS must be fixed before numba is imported
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", _SCRIPT, root, str(seed)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout + r.stderr
