"""One batched coarse launch (the multiBC sweep's 31 cases, N outer iterations) -- the command profiled with ncu."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "sr-for-cfd_b200"))
from srcfd import ensemble as E  # noqa: E402

E.coarse_stage(E.multibc_sweep(), max_iterations=int(sys.argv[1]) if len(sys.argv) > 1 else 200)
print("ok")
