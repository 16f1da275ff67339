"""SRCFD_TRACE=1 python tools/trace_gs3.py [n] [sweeps]: per-group timeline of one full-height pressure solve."""
import sys, os, time, ctypes as C
os.environ["SRCFD_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sr-for-cfd_b200"))
import numpy as np
from srcfd import _capi as capi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 400
sweeps = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
K = int(os.environ.get("SRCFD_K3", "3"))
p = capi.Params(); p.nx = p.ny = n; p.dx = p.dy = 1.0 / n; p.volp = p.dx * p.dy; p.dt = 1e-3; p.nu = 1e-2; p.rho = 1.0
p.inner_tol = 0.0; p.inner_max = sweeps
for k in range(3):
    for s in range(4): p.bc_types[k][s] = 1 if k == 2 else 0
h = capi.Handle(p)
rng = np.random.default_rng(0); Var = rng.uniform(-1, 1, (3, n + 2, n + 2)); Ff = 1e-3 * rng.uniform(-1, 1, (4, n + 2, n + 2))
p2 = capi.Params(); p2.nx = p2.ny = 40; p2.dx = p2.dy = 1.0 / 40; p2.volp = p2.dx * p2.dy; p2.dt = 1e-3; p2.nu = 1e-2; p2.rho = 1.0
p2.inner_tol = 0.0; p2.inner_max = int(os.environ.get("WARM_CODE", "0") or 0) or 30
for k in range(3):
    for s_ in range(4): p2.bc_types[k][s_] = 1 if k == 2 else 0
h2 = capi.Handle(p2)
V2 = rng.uniform(-1, 1, (3, 42, 42)); h2.upload(V2, V2, 1e-3 * rng.uniform(-1, 1, (4, 42, 42)))
flush = None
if os.environ.get("FLUSH"):
    import torch
    flush = torch.empty(int(os.environ.get("FLUSH_MB", "256")) << 20, dtype=torch.uint8, device="cuda")
for _ in range(3):
    h.upload(Var, Var, Ff); h.synchronize()
    if flush is not None:
        if os.environ["FLUSH"] == "2": flush.sum()           # read-only sweep of 256 MiB: evicts without leaving dirty lines
        else: flush.zero_()
        if os.environ.get("FLUSH_BUSY"):                      # keep the SMs busy between the flush and the solve
            xb = torch.randn(4096, 4096, device="cuda")
            for _ in range(int(os.environ["FLUSH_BUSY"])): xb = (xb @ xb) * 1e-3
        torch.cuda.synchronize()
        if os.environ.get("FLUSH_SLEEP"): time.sleep(float(os.environ["FLUSH_SLEEP"]))
    if os.environ.get("WARM_CODE"):                         # a tiny solve on another handle: pulls the kernel's code back into L2
        h2.k_solve_pressure()
    h.k_solve_pressure()
buf = np.zeros(8 * 1024 + 64, dtype=np.int64)
capi.lib().srcfd_trace_read(h._h, buf.ctypes.data_as(C.POINTER(C.c_longlong)), C.c_int64(buf.size))
G = (sweeps + K - 1) // K
tr = buf[: 8 * G].reshape(G, 8); kt = buf[8 * 1024: 8 * 1024 + 4]
t0 = kt[0]
start, mid, end = [(tr[:, i] - t0) / 1e3 for i in range(3)]
polls = tr[:, 4] - tr[:, 3]
print(f"K={K} groups={G}: kernel prologue (rhs re-lay + grid sync) {(kt[1]-kt[0])/1e3:.1f} us, run end {(kt[2]-kt[0])/1e3:.1f} us, writeback end {(kt[3]-kt[0])/1e3:.1f} us")
print(f"group duration: mean {(end-start).mean():.1f} us (first {end[0]-start[0]:.1f}); steps {2*n+2*K+1} -> {(end-start).mean()/(2*n+2*K+1)*1e3:.0f} ns/step avg")
print(f"second half (mid..end): {((end-mid).mean()):.1f} us for {n+2*K+1} steps -> {(end-mid).mean()/(n+2*K+1)*1e3:.0f} ns/step")
if G > 1:
    lag_mid = np.diff(mid); lag_end = np.diff(end)
    print(f"group-to-group lag at mid: mean {lag_mid.mean():.2f} us (min {lag_mid.min():.2f} max {lag_mid.max():.2f}); at end: mean {lag_end.mean():.2f}")
    print(f"polls per group (all threads): mean {polls.mean():.0f}, first groups {polls[:6]}")
    print("lag at mid by group decile:", np.round([lag_mid[i * len(lag_mid) // 10:(i + 1) * len(lag_mid) // 10].mean() for i in range(10)], 2))
    print("start of groups 0..5:", np.round(start[:6], 1), " ends:", np.round(end[:6], 1))
    print("group 0: start/mid/end", np.round([start[0], mid[0], end[0]], 1), " group 10:", np.round([start[10], mid[10], end[10]], 1), " group 100:", np.round([start[100], mid[100], end[100]], 1))
    mhz = (tr[:, 6] - tr[:, 5]) / np.maximum(tr[:, 2] - tr[:, 0], 1) * 1e3
    print("SM clock seen by groups 0, 10, 100, 200, 300 (MHz):", np.round(mhz[[0, 10, 100, min(200, G - 1), min(300, G - 1)]], 0))
    c = min(148, G - 1)
    print(f"group {c}: start {start[c]:.1f} (group {c-148 if c>=148 else 0} ended {end[max(c-148,0)]:.1f})")
