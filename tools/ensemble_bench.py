"""BASELINE config 3 on one GPU: a slice of the multiBC sweep (SR-warm-started from ONE batched coarse launch),
fine solves run `concurrency` at a time, aggregate cell-updates/s.  Writes gpurun_out/ensemble_bench.json.

usage: ensemble_bench.py [n_cases] [outer_iterations] [concurrency,...]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "sr-for-cfd_b200"))
from srcfd import ensemble as E, sr  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def main():
    n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    its = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    concs = [int(c) for c in sys.argv[3].split(",")] if len(sys.argv) > 3 else [1, 2, 4]
    sweep = E.multibc_sweep(max_iterations=its)
    cases = (sweep[::4] + sweep[-3:])[:n_cases]               # both lids over the Re range, plus the BFS cases
    from srcfd import bfs, ldc
    bfs._wf.verbose = ldc._wf.verbose = False
    files = dict(stats=os.path.join(GOLDEN, "stats_10to400_multiBC.txt"), encoder=sr.load_model(os.path.join(GOLDEN, "encoder10_multiBC.h5")),
                 decoder=sr.synthetic_decoder(seed=0), coarse_iterations=2000)
    E.run_local(cases[:1], concurrency=1, sr_files=files, keep_fields=False)        # load, context, code warm-up
    out = dict(cases=[c.label() for c in cases], outer_iterations=its, runs=[])
    t0 = time.perf_counter()
    warm = E.warm_stage(cases, files)
    out["warm_stage_s"] = time.perf_counter() - t0            # batched coarse launch + SR passes, all cases
    base = None
    for conc in concs:
        t0 = time.perf_counter()
        res = E.run_local(cases, concurrency=conc, sr_files=files, keep_fields=True, warm_fields=warm)
        wall = time.perf_counter() - t0
        lups = sum(int(np.sum(r.total_sweeps)) * 400 * 400 for r in res)
        if base is None:
            base = [r.fields for r in res]
        same = all(np.array_equal(a, r.fields) for a, r in zip(base, res))
        out["runs"].append(dict(concurrency=conc, fine_stage_wall_s=wall, glups=lups / wall / 1e9,
                                identical_to_first_run=same, iterations=[r.iterations for r in res],
                                case_seconds=[round(r.seconds, 3) for r in res],
                                phase_means={k: round(float(np.mean([r.phases[k] for r in res])), 4) for k in ("construct", "warm", "solve")},
                                case_sweeps=[r.total_sweeps for r in res]))
        print(out["runs"][-1], flush=True)
    print("warm stage", out["warm_stage_s"])
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "ensemble_bench.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
