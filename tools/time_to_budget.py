"""Time to a fixed outer-iteration budget on the BASELINE workloads (the reference itself does not reach its 1e-6
criterion on these grids within any practical budget): wall time of CFDSolver.solve() on the device
loop, residual norms at the end, inner sweep totals.  python tools/time_to_budget.py [iterations]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sr-for-cfd_b200"))
import numpy as np
import bench
from srcfd import ldc

its = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
out = []
# BASELINE configs[1]: BFS Re=400 400x400, SR warm start
fields, init = bench.warm_start_fields(400.0)
s = bench.make_solver(400.0)
s.settings.max_iterations = its
s._handle.set_fields(fields)
s._handle.download(s.Var, s.VarOld, s.Ff)
t0 = time.perf_counter(); n, _ = s.solve("x", verbose=False, save=False); dt = time.perf_counter() - t0
rms = np.sqrt(s.residual / (400 * 400)) / s.settings.dt
out.append({"case": "BFS Re=400 400x400 UPWIND, SR warm start", "outer_iterations": int(n), "seconds": dt, "ms_per_iteration": 1e3 * dt / n,
            "rms_u_v_p": rms.tolist(), "inner_sweeps": s.total_sweeps.tolist(), "glups": 160000 * float(s.total_sweeps.sum()) / dt / 1e9})
# BASELINE configs[0]: LDC Re=100 400x400 QUICK, zero start
s = ldc.CFDSolver(ldc.MeshParameters(nx=400, ny=400), ldc.FluidProperties(Re=100.0),
                  ldc.SolverSettings(dt=1e-3, scheme="QUICK", max_iterations=its), ldc.BoundaryConditions())
t0 = time.perf_counter(); n, _ = s.solve("x", verbose=False, save=False); dt = time.perf_counter() - t0
rms = np.sqrt(s.residual / (400 * 400)) / s.settings.dt
out.append({"case": "LDC Re=100 400x400 QUICK, zero start", "outer_iterations": int(n), "seconds": dt, "ms_per_iteration": 1e3 * dt / n,
            "rms_u_v_p": rms.tolist(), "inner_sweeps": s.total_sweeps.tolist(), "glups": 160000 * float(s.total_sweeps.sum()) / dt / 1e9})
print(json.dumps(out))
