"""Per-outer-iteration timing of the 4096^2 cavity from rest on one GPU (which pressure kernel ran, how many IEEE trips)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sr-for-cfd_b200"))
import numpy as np
from srcfd import slab, _capi as capi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
its = int(sys.argv[2]) if len(sys.argv) > 2 else 10
p = capi.Params()
p.nx = p.ny = n; p.dx = p.dy = 1.0 / n; p.volp = p.dx * p.dy
p.dt, p.nu, p.rho = 1e-3, 1e-3, 1.0
p.scheme = capi.SCHEME_QUICK
for k in range(3):
    for s in range(4):
        p.bc_types[k][s] = 1 if k == 2 else 0
p.bc_values[0][2] = 1.0
p.inner_tol, p.inner_max, p.sweep_order, p.device = 1e-6, 1000, capi.ORDER_JACOBI, 0
s = slab.GpuSlab(p, 1, 0)
s.h.initialize_fields(True)
prev = s.h.status()["total_sweeps"].copy()
for it in range(its):
    s.h.timer_start()
    slab.step([s], 1, (0.0, 0.0, 0.0))
    ms = s.h.timer_stop()
    st = s.h.status(); ks = s.kernel_stats()
    sw = st["total_sweeps"] - prev; prev = st["total_sweeps"].copy()
    print(json.dumps({"it": it, "ms": round(ms, 2), "sweeps": sw.tolist(), "stats": ks, "replays": s.info()["replays"]}), flush=True)
