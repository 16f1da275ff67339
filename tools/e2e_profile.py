"""Where the host-array path spends its time (BFS Re=400 400x400): python tools/e2e_profile.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sr-for-cfd_b200"))
import numpy as np
from srcfd import bfs
import bench
s = bench.make_solver(400.0)
s._sync_params()
H = s._handle
for _ in range(3): s._implicit_solve(); s._convergence_check()
T = {}
def tm(name, f):
    t0 = time.perf_counter(); r = f(); H.synchronize(); T[name] = T.get(name, 0.0) + time.perf_counter() - t0; return r
N = 10
t_all = time.perf_counter()
for _ in range(N):
    tm("sync_params", s._sync_params)
    tm("upload", lambda: H.upload(s.Var, s.VarOld, s.Ff))
    tm("k_implicit_solve", H.k_implicit_solve)
    tm("download", lambda: H.download(s.Var, None, s.Ff, s.residual))
    tm("status", H.status)
    tm("convergence_check", s._convergence_check)
t_all = time.perf_counter() - t_all
for k, v in T.items(): print(f"{k:20s} {v / N * 1e3:7.3f} ms")
print(f"{'sum (serialised)':20s} {t_all / N * 1e3:7.3f} ms")
t0 = time.perf_counter()
for _ in range(N): s._implicit_solve(); s._convergence_check()
print(f"{'as called':20s} {(time.perf_counter() - t0) / N * 1e3:7.3f} ms")
