"""One rank's compute of a thin slab on one GPU: nx x ny plane, JACOBI pressure relaxation, 1000 sweeps (no exchange)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sr-for-cfd_b200"))
import numpy as np
from srcfd import slab, _capi as capi
nx, ny = int(sys.argv[1]), int(sys.argv[2])
p = capi.Params()
p.nx, p.ny = nx, ny
p.dx, p.dy = 1.0 / 4096, 1.0 / ny
p.volp = p.dx * p.dy
p.dt, p.nu, p.rho = 1e-3, 1e-3, 1.0
for k in range(3):
    for s in range(4):
        p.bc_types[k][s] = 1 if k == 2 else 0
p.inner_tol, p.inner_max, p.sweep_order, p.device = 0.0, 1000, capi.ORDER_JACOBI, 0
s = slab.GpuSlab(p, 1, 0)
rng = np.random.default_rng(0)
Var = np.zeros((3, nx + 2, ny + 2)); Var[2] = rng.uniform(-1, 1, (nx + 2, ny + 2))
Ff = 1e-3 * rng.uniform(-1, 1, (4, nx + 2, ny + 2))
s.h.upload(Var=Var, Ff=Ff)
slab.solve_pressure([s])
s.h.timer_start()
n, _ = slab.solve_pressure([s])
ms = s.h.timer_stop()
print(json.dumps({"nx": nx, "ny": ny, "ms": ms, "glups": nx * ny * n / ms / 1e6, "us_per_pass": 1e3 * ms / (n / 4), "stats": s.kernel_stats()}))
