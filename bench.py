#!/usr/bin/env python
"""bench.py -- headline benchmark of the fine-grid Navier-Stokes hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): backward-facing step, Re=400, 400x400 cells on a 10x3 domain,
UPWIND, dt=2e-3, under-relaxation .5/.5/.2, fine field warm-started by the SR autoencoder
(real committed encoder weights; the decoder weights are absent from the reference tree, so a
seed-0 synthetic decoder is used -- stated in `config`).  One "step" = one outer iteration
(_implicit_solve + _convergence_check: two momentum solves, the <=1000-sweep pressure solve, velocity
correction, BCs, flux update).  Metric: fine-grid cell-updates/s = interior cells x inner relaxation
sweeps executed / device time, in GLUP/s.  N>1: one independent case per GPU (Re-sweep ensemble,
no data-path collective), weak scaling, time = max over ranks.

`value`   inputs resident in HBM, CUDA-event time on the library's stream, L2 flushed between steps.
`e2e`     the same metric through the reference-facing API (CFDSolver._implicit_solve /
          _convergence_check on HOST numpy arrays): H2D of Var/VarOld/Ff and D2H of Var/VarOld/Ff/
          residual inside the timed region every step.
`roofline` dominant kernel (pressure-Poisson inner solve): algorithmic 24 B per cell-update x
          updates per launch / mean launch duration (CUDA events around every launch).
`cpu_baseline` / --impl reference: the CPU oracle port of the reference kernels (C, OpenMP over rows =
          numba prange's chunked in-place sweep) on the box's host cores, bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "sr-for-cfd_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

os.environ["NCCL_DEBUG"] = os.environ.get("SRCFD_NCCL_DEBUG", "WARN")   # keep stdout to the one JSON line

NX = NY = 400
LX, LY = 10.0, 3.0
DT = 2e-3
RELAX = {'u': 0.5, 'v': 0.5, 'p': 0.2}
ENSEMBLE_RE = [400.0, 100.0, 200.0, 300.0, 500.0, 600.0, 700.0, 50.0]
BYTES_PER_LUP_PRESSURE = 24.0     # read p, read rhs, write p
GOLDEN = os.path.join(ROOT, "tests", "golden")
WARM_DESC = "coarse 10x10 solve (2000 its) -> encoder_10 (committed weights) + decoder_400 (synthetic seed-0 weights)"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per pressure-solve launch from the committed ncu --set full capture."""
    for name in ("ncu_pressure_r01c.json", "ncu_pressure_r01.json"):     # latest capture of the kernel in use first
        path = os.path.join(ROOT, "profiles", name)
        if os.path.exists(path):
            with open(path) as f:
                return json.load(f).get("dram_bytes_per_launch")
    return None


def warm_start_fields(Re):
    """Coarse 10x10 BFS solve -> SR autoencoder -> (3, ny, nx) float32 initial guess, all on the GPU path
    (bfs_ml_accelerated.py:1310-1517).  Returns (fields, description)."""
    from srcfd import bfs, sr
    bfs._wf.verbose = False
    cache = os.path.join(ROOT, "gpurun_out", f"warm_Re{Re:g}.npy")
    if os.environ.get("SRCFD_BENCH_WARM_CACHE") and os.path.exists(cache):     # profiling runs: skip the 10x10 coarse solve
        return np.load(cache), WARM_DESC + " [field cached by the preceding plain run]"
    coarse = bfs.run_coarse_simulation(Re=Re, lr_dim=10, dt=DT, scheme='UPWIND', max_iterations=2000,
                                       relaxation_factors=RELAX, save=False)
    dec = sr.synthetic_decoder(seed=0)
    hr = bfs.ml_super_resolution(coarse, 10, 400, os.path.join(GOLDEN, "stats_10to400_multiBC.txt"),
                                 os.path.join(GOLDEN, "encoder10_multiBC.h5"), dec,
                                 use_aspect_ratio_correction=True, lx=LX, ly=LY, blend_factor=0.3)
    f = np.stack([np.asarray(hr[c], dtype=np.float32) for c in "uvp"])
    return f, WARM_DESC


def make_solver(Re, device=0):
    from srcfd import bfs
    bc = bfs.BoundaryConditions()
    bc.u_boundaries['left'] = bfs.BoundaryCondition('dirichlet', 0.0)
    bc.u_boundaries['right'] = bfs.BoundaryCondition('neumann', 0.0)
    bc.v_boundaries['right'] = bfs.BoundaryCondition('neumann', 0.0)
    bc.p_boundaries['right'] = bfs.BoundaryCondition('dirichlet', 0.0)
    st = bfs.SolverSettings(dt=DT, scheme='UPWIND', max_iterations=10 ** 9, relaxation_factors=dict(RELAX))
    return bfs.CFDSolver(bfs.MeshParameters(nx=NX, ny=NY, lx=LX, ly=LY), bfs.FluidProperties(Re=Re), st, bc,
                         step_height=1.0, h=2.0, Ub=1.0, device=device)


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace('.', '', 1).isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace('.', '', 1).isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def cpu_baseline(Re, fields, steps, threads=None):
    """The oracle port on the host cores: `steps` outer iterations of the SAME workload (bounded sample)."""
    from oracle import oracle as O
    # torchrun exports OMP_NUM_THREADS=1; the CPU arm is meant to use every host core it can
    O.set_num_threads(threads or len(os.sched_getaffinity(0)))
    cores = O.num_threads()
    case = O.bfs_case(NX, NY, Re=Re, dt=DT, scheme="UPWIND", lx=LX, ly=LY, relax=(0.5, 0.5, 0.2),
                      order=O.ORDER_GS_OMP if cores > 1 else O.ORDER_GS_LEX)
    o = O.OracleSolver(case)
    if fields is not None:
        o.set_interior({"u": fields[0], "v": fields[1], "p": fields[2]})
    o.implicit_solve(); o.convergence_check()                    # warm caches / page in
    per_step, sweeps = [], 0
    for _ in range(steps):
        t0 = time.perf_counter()
        sw = o.implicit_solve()
        o.convergence_check()
        per_step.append(time.perf_counter() - t0)
        sweeps += int(sw.sum())
    t = float(np.sum(per_step))
    return {"value": NX * NY * sweeps / t / 1e9, "unit": "GLUP/s", "cores": cores, "kind": "port",
            "sample": f"{steps} outer iterations of the same BFS Re={Re:g} {NX}x{NY} workload, {sweeps} inner sweeps, "
                      f"{t:.2f} s; C/OpenMP port of the reference numba kernels (rows split over threads like prange)",
            "ms_per_step": 1e3 * t / steps}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    fields, init = warm_start_fields_cpu_safe(ENSEMBLE_RE[0])
    steps = max(1, args.steps)
    cb = cpu_baseline(ENSEMBLE_RE[0], fields, min(steps, 40))
    line = {"metric": "fine-grid cell-updates/s", "value": cb["value"], "unit": "GLUP/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": args.warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": workload_config(init, args.gpus), "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "GLUP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def warm_start_fields_cpu_safe(Re):
    """The reference arm may run where CUDA is busy/absent: use the cached warm-start field if the ours-arm
    wrote it, else fall back to a zero field (the CPU arm times sweeps, which are data-independent in cost)."""
    cache = os.path.join(ROOT, "gpurun_out", f"warm_Re{Re:g}.npy")
    if os.path.exists(cache):
        return np.load(cache), "cached SR warm-start field"
    try:
        from srcfd import _capi
        if _capi.device_count() > 0:
            f, d = warm_start_fields(Re)
            return f, d
    except Exception:
        pass
    return None, "zero field"


def workload_config(init, n):
    return {"workload": f"BFS Re=400 {NX}x{NY} ({LX:g}x{LY:g} domain), UPWIND, dt={DT}, relax .5/.5/.2, "
                        "reference sweep order (in-place lexicographic Gauss-Seidel), inner tol 1e-6 / cap 1000",
            "init": init,
            "step": "one outer iteration: 2 momentum solves + pressure solve + correction + BCs + flux update + convergence check",
            "ensemble": f"{n} independent case(s), one per GPU, Re in {ENSEMBLE_RE[:n]}" if n > 1 else "single case",
            "l2": "flushed between timed steps (256 MiB write); the 12.9 MB solver state is L2-resident within a step by design",
            "parallelism": f"ensemble x{n} (no data-path collective)" if n > 1 else "1 GPU"}


def run_ours(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    Re = ENSEMBLE_RE[rank % len(ENSEMBLE_RE)]
    fields, init = warm_start_fields(Re)
    if rank == 0 and fields is not None:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        np.save(os.path.join(ROOT, "gpurun_out", f"warm_Re{Re:g}.npy"), fields)
    solver = make_solver(Re, device=local)
    H = solver._handle
    solver._sync_params()
    if fields is not None:
        H.set_fields(fields)
    crit = (0.0, 0.0, 0.0)                      # never "converged": every timed step does full work
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=f"cuda:{local}")
    cells = NX * NY

    # ---- device-resident timing --------------------------------------------------------------
    H.reset_counters()
    for _ in range(args.warmup):
        H.step(1, crit)
    H.synchronize()
    st0 = H.status()
    H.timing_enable(True)
    launches0 = H.launch_count()
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    step_ms = []
    for _ in range(args.steps):
        if not os.environ.get("SRCFD_BENCH_NOFLUSH"):      # experiments only: the contract run always flushes
            flush.zero_()
        torch.cuda.synchronize()
        H.timer_start()
        H.step(1, crit)
        step_ms.append(H.timer_stop())
    barrier()
    clocks = sampler.stop()
    st1 = H.status()
    tr = H.timing_read()
    H.timing_enable(False)
    launches = H.launch_count() - launches0
    sweeps = (st1["total_sweeps"] - st0["total_sweeps"]).astype(np.int64)
    t_ms = float(np.sum(step_ms))

    # ---- end to end through the host-array API -------------------------------------------------
    h2d = solver.Var.nbytes + solver.VarOld.nbytes + solver.Ff.nbytes
    d2h = solver.Var.nbytes + solver.Ff.nbytes + solver.residual.nbytes + solver.VarOld.nbytes
    H.download(solver.Var, solver.VarOld, solver.Ff)
    e2e_steps = max(2, min(args.steps, 10))
    solver._implicit_solve(); solver._convergence_check()
    e2e_sweeps = 0
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        solver._implicit_solve()
        e2e_sweeps += int(np.sum(solver.last_sweeps))
        solver._convergence_check()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0

    # ---- reduce over ranks ---------------------------------------------------------------------
    tot_lup = float(cells * sweeps.sum())
    e2e_lup = float(cells * e2e_sweeps)
    if world > 1:
        v = torch.tensor([t_ms, e2e_s], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(v, op=dist.ReduceOp.MAX)
        w = torch.tensor([tot_lup, e2e_lup, float(launches)], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(w, op=dist.ReduceOp.SUM)
        t_ms, e2e_s = v.tolist()
        tot_lup, e2e_lup, launches = w.tolist()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_src = peaks()
    p_lups = cells * float(sweeps[2])
    p_ms = tr["pressure_ms"]
    achieved = BYTES_PER_LUP_PRESSURE * p_lups / (p_ms * 1e-3) / 1e9 if p_ms > 0 else 0.0
    cb = cpu_baseline(Re, fields, 3) if world == 1 and not args.no_cpu else None
    line = {
        "metric": "fine-grid cell-updates/s", "value": tot_lup / (t_ms * 1e-3) / 1e9, "unit": "GLUP/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(init, world),
        "inner_sweeps_per_step": {"u": float(sweeps[0]) / args.steps, "v": float(sweeps[1]) / args.steps,
                                  "p": float(sweeps[2]) / args.steps},
        "e2e": {"value": e2e_lup / e2e_s / 1e9, "unit": "GLUP/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "ms_per_step": 1e3 * e2e_s / e2e_steps,
                "api": "CFDSolver._implicit_solve() + _convergence_check() on host numpy arrays"},
        "gpu_launches": int(launches),
        "step_breakdown_ms": {"pressure_solve": p_ms / args.steps, "momentum_solves": tr["momentum_ms"] / args.steps,
                              "bc_flux_correction_and_gaps": (t_ms - p_ms - tr["momentum_ms"]) / args.steps},
        "roofline": {"bound": "hbm", "kernel": "k_solve_gs3 (solve_pressure inner loop: full-height sweep groups)",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "peak_source": peak_src, "traffic": ncu_traffic(),
                     "algorithmic_bytes_per_launch": BYTES_PER_LUP_PRESSURE * p_lups / max(1, tr["pressure_launches"]),
                     "launches": tr["pressure_launches"], "avg_launch_ms": p_ms / max(1, tr["pressure_launches"]),
                     "share_of_step": p_ms / t_ms if t_ms else None,
                     "note": "24 B per cell-update x updates per launch; the 400^2 planes stay in L2/SMEM across sweeps "
                             "(temporal reuse), so this is an algorithmic-bytes rate, not DRAM traffic"},
        "clocks": clocks,
    }
    if cb:
        line["cpu_baseline"] = cb
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
