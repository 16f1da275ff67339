"""Scratch timing of the inner solves (not the bench contract): pressure/momentum at a given size and order."""
import sys, time, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sr-for-cfd_b200"))
import numpy as np
from srcfd import _capi as capi

def run(n, order, sweeps=1000, reps=3, scheme=0):
    p = capi.Params()
    p.nx = p.ny = n
    p.dx = p.dy = 1.0 / n; p.volp = p.dx * p.dy; p.dt = 1e-3; p.nu = 1e-2; p.rho = 1.0
    p.scheme = scheme; p.inner_tol = 0.0; p.inner_max = sweeps; p.sweep_order = order
    for k in range(3):
        for s in range(4):
            p.bc_types[k][s] = 1 if k == 2 else 0
    h = capi.Handle(p)
    rng = np.random.default_rng(0)
    Var = rng.uniform(-1, 1, (3, n + 2, n + 2)); Ff = 1e-3 * rng.uniform(-1, 1, (4, n + 2, n + 2))
    h.upload(Var, Var, Ff)
    h.timing_enable(True)
    out = {}
    for what in ("pressure", "momentum"):
        ts = []
        for r in range(reps):
            h.upload(Var, Var, Ff); h.reset_counters()
            t0 = time.perf_counter()
            nsw, rms = h.k_solve_pressure() if what == "pressure" else h.k_solve_momentum(0, scheme)
            ts.append(time.perf_counter() - t0)
        t = min(ts)
        out[what] = (nsw, t * 1e3, n * n * nsw / t / 1e9)
    tr = h.timing_read()
    h.close()
    return out, tr

if __name__ == "__main__":
    sizes = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [400]
    sweeps = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    for n in sizes:
        for order, name in ((0, "GS_LEX"), (1, "JACOBI"), (2, "RED_BLACK")):
            out, tr = run(n, order, sweeps)
            print(f"n={n} {name:9s} pressure: {out['pressure'][0]} sweeps {out['pressure'][1]:.3f} ms {out['pressure'][2]:.2f} GLUP/s | "
                  f"momentum(upwind): {out['momentum'][0]} sweeps {out['momentum'][1]:.3f} ms {out['momentum'][2]:.2f} GLUP/s | events {tr}", flush=True)
