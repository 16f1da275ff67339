"""Wall time of a lid-driven-cavity solve on the device loop (not the bench contract): python tools/ldc_run.py [n] [its] [scheme]"""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sr-for-cfd_b200"))
import numpy as np
from srcfd import ldc
n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
its = int(sys.argv[2]) if len(sys.argv) > 2 else 300
scheme = sys.argv[3] if len(sys.argv) > 3 else "QUICK"
for rep in range(2):
    s = ldc.CFDSolver(ldc.MeshParameters(nx=n, ny=n), ldc.FluidProperties(Re=100.0),
                      ldc.SolverSettings(dt=1e-3, scheme=scheme, max_iterations=its), ldc.BoundaryConditions())
    t0 = time.perf_counter()
    k, _ = s.solve("x", verbose=False, save=False)
    dt = time.perf_counter() - t0
sw = s.total_sweeps
print(f"LDC {n}x{n} {scheme}: {k} outer iterations in {dt*1e3:.1f} ms ({dt/k*1e3:.3f} ms/it); sweeps u,v,p = {sw.tolist()}; "
      f"{n*n*sw.sum()/dt/1e9:.1f} GLUP/s")
