"""Time to the 1e-6 criterion, double-lid cavity Re=1050 (the reference's published case), reference sweep order vs
RB_JACOBI (Jacobi momentum + red-black pressure), and the distance between the two converged fields."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sr-for-cfd_b200"))
import numpy as np
from srcfd import ldc
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
orders = sys.argv[2].split(",") if len(sys.argv) > 2 else ["GS_LEX", "RB_JACOBI"]
out, fields = {}, {}
for order in orders:
    bc = ldc.BoundaryConditions()
    bc.u_boundaries['bottom'] = ldc.BoundaryCondition('dirichlet', 1.0)
    s = ldc.CFDSolver(ldc.MeshParameters(nx=n, ny=n), ldc.FluidProperties(Re=1050.0),
                      ldc.SolverSettings(dt=1e-3, scheme="QUICK", max_iterations=150000, sweep_order=order), bc)
    t0 = time.perf_counter()
    its, _ = s.solve("x", verbose=False, save=False)
    dt = time.perf_counter() - t0
    fields[order] = s.Var.copy()
    out[order] = dict(iterations=int(its), seconds=dt, ms_per_iteration=1e3 * dt / its, converged=bool(s.converged),
                      sweeps=[int(x) for x in s.total_sweeps])
    print(order, json.dumps(out[order]), flush=True)
if len(orders) == 2:
    a, b = fields[orders[0]], fields[orders[1]]
    rel = []
    for k in range(3):
        x, y = a[k, 1:-1, 1:-1], b[k, 1:-1, 1:-1]
        if k == 2: x, y = x - x.mean(), y - y.mean()
        rel.append(float(np.linalg.norm(x - y) / np.linalg.norm(x)))
    print("relL2 u,v,p(demeaned):", rel)
