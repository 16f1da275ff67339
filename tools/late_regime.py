"""Where an outer iteration's time goes in the converging regime (few sweeps per inner solve): advance the published
double-lid case by N iterations, then time M iterations with the per-solve stopwatch.  python tools/late_regime.py [n] [N] [M]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sr-for-cfd_b200"))
import numpy as np  # noqa: E402
from srcfd import ldc  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 400
N = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
M = int(sys.argv[3]) if len(sys.argv) > 3 else 500
bc = ldc.BoundaryConditions()
bc.u_boundaries['bottom'] = ldc.BoundaryCondition('dirichlet', 1.0)
s = ldc.CFDSolver(ldc.MeshParameters(nx=n, ny=n), ldc.FluidProperties(Re=1050.0),
                  ldc.SolverSettings(dt=1e-3, scheme="QUICK", max_iterations=N), bc)
s.solve("x", verbose=False, save=False)
H = s._handle
H.upload(s.Var, s.VarOld, s.Ff)
crit = (1e-6,) * 3
H.step(10, crit); H.synchronize()
st0, l0 = H.status(), H.launch_count()
H.timing_enable(True)
t0 = time.perf_counter()
H.step(M, crit); H.synchronize()
wall = time.perf_counter() - t0
tr, st1 = H.timing_read(), H.status()
H.timing_enable(False)
sw = (st1["total_sweeps"] - st0["total_sweeps"]) / M
out = dict(grid=n, after_iterations=N, timed_iterations=M, ms_per_iteration=1e3 * wall / M, sweeps_per_iteration=sw.tolist(),
           pressure_ms=tr["pressure_ms"] / M, pressure_launches_per_iteration=tr["pressure_launches"] / M,
           momentum_ms=tr["momentum_ms"] / M, momentum_launches_per_iteration=tr["momentum_launches"] / M,
           launches_per_iteration=(H.launch_count() - l0) / M)
out["rest_ms"] = out["ms_per_iteration"] - out["pressure_ms"] - out["momentum_ms"]
print(json.dumps(out))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"late_regime_{n}.json"), "w"), indent=1)
