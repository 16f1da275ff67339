"""SRCFD_TRACE=1 python tools/trace_gs2.py [n] [sweeps]: per-task timeline of one K-sweep pressure solve."""
import sys, os, ctypes as C
os.environ["SRCFD_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sr-for-cfd_b200"))
import numpy as np
from srcfd import _capi as capi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 400
sweeps = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
p = capi.Params(); p.nx = p.ny = n; p.dx = p.dy = 1.0 / n; p.volp = p.dx * p.dy; p.dt = 1e-3; p.nu = 1e-2; p.rho = 1.0
p.inner_tol = 0.0; p.inner_max = sweeps
for k in range(3):
    for s in range(4): p.bc_types[k][s] = 1 if k == 2 else 0
h = capi.Handle(p)
rng = np.random.default_rng(0); Var = rng.uniform(-1, 1, (3, n + 2, n + 2)); Ff = 1e-3 * rng.uniform(-1, 1, (4, n + 2, n + 2))
h.upload(Var, Var, Ff)
MOM = os.environ.get("TRACE_OP", "pressure") != "pressure"      # TRACE_OP=momentum: the u solve (upwind) instead
for _ in range(2):
    h.upload(Var, Var, Ff); h.reset_counters()
    print("sweeps, rms:", h.k_solve_momentum(0, 0) if MOM else h.k_solve_pressure())
K = 4 if MOM else 8
buf = np.zeros(((sweeps // 4 + 2) * 64 * 8), dtype=np.int64)
capi.lib().srcfd_trace_read(h._h, buf.ctypes.data_as(C.POINTER(C.c_longlong)), C.c_int64(buf.size))
tr = buf.reshape(-1, 8); tr = tr[tr[:, 0] > 0]
ngroups = (sweeps + K - 1) // K; B = len(tr) // ngroups
tr = tr[: ngroups * B].reshape(ngroups, B, 8)
t0 = tr[:, :, 0].min()
start, first, end, wait, steps = [tr[:, :, i] for i in range(5)]
print(f"bands={B} groups={ngroups} total={(end.max() - t0) / 1e3:.1f} us")
print(f"publishes per task: mean {tr[:, :, 6].mean():.1f}; max step gap between publishes: mean {tr[:, :, 7].mean():.1f} max {tr[:, :, 7].max()}")
run = (end - first) / 1e3
print(f"task run time (first step..end): mean {run.mean():.1f} us; waited inside: mean {wait.mean() / 1e3:.1f} us; steps {steps[0, 0]}")
print(f"=> busy step time {(run.mean() - wait.mean() / 1e3) / steps[0, 0] * 1e3:.0f} ns/step; idle before first step: mean {((first - start) / 1e3).mean():.1f} us")
g = np.diff(first[:, 0]) / 1e3
print(f"group-to-group start lag (band 0): mean {g.mean():.1f} us, min {g.min():.1f}, max {g.max():.1f}")
t53 = tr[:, :, 5]
for b in range(1, B):
    hop = (first[:, b] - t53[:, b - 1]) / 1e3
    print(f"  hop: band {b} first step - band {b-1} reached the step that allows it: mean {hop.mean():.1f} us  median {np.median(hop):.1f} min {hop.min():.1f}")
    d = (first[:, b] - first[:, b - 1]) / 1e3
    print(f"band {b} starts {d.mean():.1f} us after band {b - 1}")
print("first 3 groups, band 0: start/first/end (us):", [(round((start[i, 0] - t0) / 1e3, 1), round((first[i, 0] - t0) / 1e3, 1), round((end[i, 0] - t0) / 1e3, 1)) for i in range(3)])
if len(sys.argv) > 3:
    g0 = int(sys.argv[3])
    for g_ in range(g0, min(g0 + 3, ngroups)):
        for b_ in range(B):
            print(f"grp {g_} band {b_}: picked {(start[g_, b_] - t0) / 1e3:8.1f}  first {(first[g_, b_] - t0) / 1e3:8.1f}  end {(end[g_, b_] - t0) / 1e3:8.1f}  waited {wait[g_, b_] / 1e3:6.1f}  sm {tr[g_, b_, 5]}")
