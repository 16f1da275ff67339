"""Coarse stage (10x10 solves of the multiBC sweep) timing: one launch for the whole sweep (one CTA per case)
against the per-case whole-GPU kernel path.  Writes gpurun_out/coarse_bench.json."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "sr-for-cfd_b200"))

from srcfd import bfs, ensemble as E, ldc  # noqa: E402


def main():
    its = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    cases = E.multibc_sweep()
    out = dict(cases=len(cases), lr_dim=10, max_iterations=its)
    E.coarse_stage(cases[:2], max_iterations=10)                  # module load, context
    t0 = time.perf_counter(); got = E.coarse_stage(cases, max_iterations=its); out["batched_wall_s"] = time.perf_counter() - t0
    from srcfd import _capi as capi, solver as S
    # device time and sweep totals of the same launch
    t0 = time.perf_counter()
    n_k = 4
    ref = []
    for spec in cases[:n_k]:
        wf = (bfs if spec.kind == "bfs" else ldc)._wf
        wf.verbose = False
        S.CFDSolver.resident_solve = False
        ref.append(wf.run_coarse_simulation(Re=spec.Re, lr_dim=10, max_iterations=its, bc=E._case_bc(spec), save=False))
        S.CFDSolver.resident_solve = True
    out["kernel_path_wall_s_per_case"] = (time.perf_counter() - t0) / n_k
    out["identical_to_kernel_path"] = all(np.array_equal(got[i][n], ref[i][n]) for i in range(n_k) for n in "uvp")
    out["speedup_vs_kernel_path"] = out["kernel_path_wall_s_per_case"] * len(cases) / out["batched_wall_s"]
    # the reference's own budget (100 000 iterations, LDC.py:1372), whole sweep in one launch
    if "--full" in sys.argv:
        t0 = time.perf_counter(); E.coarse_stage(cases, max_iterations=100000); out["batched_100k_wall_s"] = time.perf_counter() - t0
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/coarse_bench.json", "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
