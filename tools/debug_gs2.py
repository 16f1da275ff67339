import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sr-for-cfd_b200"))
import numpy as np
from srcfd import kernels as K
from oracle import oracle as O
np.set_printoptions(linewidth=200, precision=5)
for (Nx, Ny) in [(1, 1), (2, 3), (13, 9), (100, 40)]:
    rng = np.random.default_rng(5)
    Var = rng.uniform(-1, 1, (3, Nx + 2, Ny + 2)); VarOld = Var + 0.01 * rng.uniform(-1, 1, Var.shape)
    Ff = 0.05 * rng.uniform(-1, 1, (4, Nx + 2, Ny + 2))
    dx, dy = 1.3 / Nx, 0.9 / Ny; volp, dt, nu, rho = dx * dy, 2e-3, 1 / 250.0, 1.0
    for cap in (1, 2, 3, 8, 9, 17):
        A, B = Var.copy(), Var.copy()
        n = K.solve_pressure(A, Ff, Nx, Ny, dx, dy, dt, rho, volp, tolerance=0.0, max_iter=cap)
        m = O.solve_pressure(B, Ff, Nx, Ny, dx, dy, dt, rho, volp, tolerance=0.0, max_iter=cap)
        d = np.abs(A - B)
        print(f"pressure {Nx}x{Ny} cap={cap}: n={n} m={m} maxdiff={d.max():.3e} nbad={int((d>0).sum())}", flush=True)
        if d.max() > 0 and Nx <= 2:
            print("gpu", A[2]); print("ref", B[2])
    for cap in (1, 2, 5):
        A, B = Var.copy(), Var.copy()
        n = K.solve_momentum_quick(A, VarOld, Ff, 0, Nx, Ny, dx, dy, dt, nu, volp, tolerance=0.0, max_iter=cap)
        m = O.solve_momentum_quick(B, VarOld, Ff, 0, Nx, Ny, dx, dy, dt, nu, volp, tolerance=0.0, max_iter=cap)
        d = np.abs(A - B)
        print(f"quick {Nx}x{Ny} cap={cap}: n={n} m={m} maxdiff={d.max():.3e} nbad={int((d>0).sum())}", flush=True)
