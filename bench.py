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
          numba prange's chunked in-place sweep) on the box's host cores, bounded sample.  That arm never imports the
          GPU package.
Sub-records on the same JSON line (each event-timed on the library's stream):
`large_grid` (N=1)  4096x4096 cavity, JACOBI order: temporally blocked pressure relaxation and upwind / QUICK momentum
          sweeps -- the HBM-bound regime; its `roofline` is the one the >= 0.70 target is judged on.
`slab`    BASELINE configs[3]: the same 4096x4096 pressure relaxation and 10 whole outer iterations (QUICK) split into
          row slabs over the N ranks (strong scaling), halo rows and residual sums pushed through cudaIpc-mapped peer
          memory by the kernels; `slab_parity`: every rank's owned rows bit-equal to the single-domain result computed
          in the same run on the same GPU.
`ensemble` BASELINE configs[2]: the 31-case multiBC sweep shared by the ranks (strong scaling), SR-warm-started.
`time_to_converged` (N=1)  double-lid cavity Re=1050 100x100 to the reference's 1e-6 criterion (published: 212.41 s).
`decoder` (N=1)  BASELINE configs[4]: decoder_400, batch 1024.
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

# Host-side library thread pools only get in the way of the worker threads that drive concurrent cases (measured on the
# ensemble record: 66 -> 87 GLUP/s); the CPU baseline sets its own thread count explicitly (oracle.set_num_threads).
os.environ.setdefault("OMP_NUM_THREADS", "1")

import numpy as np  # noqa: E402

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
            "config": dict(workload_config(init, args.gpus),
                           reference_scope="ONE case on this host's cores; for --gpus N > 1 the GPU arm runs N cases on N GPUs, "
                                           "so the driver's ratio there reads 'N GPUs vs one host' (a rate: running the N-case "
                                           "share back to back on the same cores gives the same GLUP/s)"),
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "GLUP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def warm_start_fields_cpu_safe(Re):
    """The reference arm never touches the GPU package: it starts from the SR field the ours-arm cached in gpurun_out/
    when that file exists, else from a zero field -- the arm reports a RATE (cell updates per second), and the cost of
    a sweep does not depend on the field values."""
    cache = os.path.join(ROOT, "gpurun_out", f"warm_Re{Re:g}.npy")
    if os.path.exists(cache):
        return np.load(cache), "cached SR warm-start field (written by the GPU arm)"
    return None, "zero field (the rate does not depend on the field values)"


def workload_config(init, n):
    return {"workload": f"BFS Re=400 {NX}x{NY} ({LX:g}x{LY:g} domain), UPWIND, dt={DT}, relax .5/.5/.2, "
                        "reference sweep order (in-place lexicographic Gauss-Seidel), inner tol 1e-6 / cap 1000",
            "init": init,
            "step": "one outer iteration: 2 momentum solves + pressure solve + correction + BCs + flux update + convergence check",
            "ensemble": f"{n} independent case(s), one per GPU, Re in {ENSEMBLE_RE[:n]}" if n > 1 else "single case",
            "l2": "flushed between timed steps (256 MiB write); the 12.9 MB solver state is L2-resident within a step by design",
            "parallelism": f"ensemble x{n} (no data-path collective)" if n > 1 else "1 GPU"}


# ---------------------------------------------------------------------------------------------------------------------
# sub-records
# ---------------------------------------------------------------------------------------------------------------------
LG_N = 4096
BYTES_PER_LUP_MOMENTUM = 40.0     # phi, phi_old, 2 face fluxes, write phi


def _ldc_params(n, device, inner_max, tol, scheme_quick=True, Re=1000.0, nx=None):
    from srcfd import _capi as capi
    p = capi.Params()
    p.nx, p.ny = (nx or n), n
    p.dx = p.dy = 1.0 / n
    p.volp = p.dx * p.dy
    p.dt, p.nu, p.rho = 1e-3, 1.0 / Re, 1.0
    p.scheme = capi.SCHEME_QUICK if scheme_quick else capi.SCHEME_UPWIND
    for k in range(3):
        for s in range(4):
            p.bc_types[k][s] = 1 if k == 2 else 0
    p.bc_values[0][2] = 1.0                                   # the lid (PyCFD_ML_accelerated.py:47-67)
    p.inner_tol, p.inner_max, p.sweep_order, p.device = tol, inner_max, capi.ORDER_JACOBI, device
    return p


def _synthetic_rows(n, g0, g1, nx=None):
    """Rows g0..g1 of the synthetic nx x n state (nx = n unless given; seeded per GLOBAL row, so every world size sees the
    same field): p ~ U(-1,1), face fluxes ~ 1e-3 U(-1,1) with exactly zero boundary-face fluxes along i (wall)."""
    Var = np.zeros((3, g1 - g0 + 1, n + 2)); Ff = np.zeros((4, g1 - g0 + 1, n + 2))
    for r in range(g0, g1 + 1):
        rr = np.random.default_rng(1000 + r)
        Var[2, r - g0] = rr.uniform(-1, 1, n + 2)
        Var[:2, r - g0] = 0.1 * rr.uniform(-1, 1, (2, n + 2))
        Ff[:, r - g0] = 1e-3 * rr.uniform(-1, 1, (4, n + 2))
        if r == 1: Ff[2, r - g0] = 0.0
        if r == (nx or n): Ff[0, r - g0] = 0.0
    return Var, Ff


def ncu_value(name, key):
    path = os.path.join(ROOT, "profiles", name)
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f).get(key)
    return None


def large_grid_record(device):
    """N=1: the HBM-bound regime.  4096^2, JACOBI order, inputs resident in HBM (the planes are 134 MB each: nothing
    is L2-resident), CUDA events on the library's stream."""
    from srcfd import slab, _capi as capi
    peak, peak_src = peaks()
    n = LG_N
    out = {"grid": [n, n], "order": "JACOBI (every cell from the previous iterate; bit-identical to the oracle's restatement)",
           "l2": "planes of 134 MB each exceed the 126 MB L2"}
    s = slab.GpuSlab(_ldc_params(n, device, 1000, 0.0), 1, 0)
    Var, Ff = _synthetic_rows(n, 0, n + 1)
    s.h.upload(Var=Var, VarOld=Var, Ff=Ff)
    del Var, Ff
    slab.solve_pressure([s])                                  # warm-up
    s.h.synchronize()
    s.h.timer_start()
    sw, _ = slab.solve_pressure([s])
    ms = s.h.timer_stop()
    lups = float(n) * n * sw
    ach = BYTES_PER_LUP_PRESSURE * lups / (ms * 1e-3) / 1e9
    traffic = ncu_value("ncu_jtb_4096_r02.json", "dram_bytes_per_launch")
    out["pressure"] = {"kernel": "k_jtb2_pass<4, 2> (solve_pressure, warp-streaming, 4 sweeps per pass over HBM)", "sweeps": sw, "ms": ms,
                       "value": lups / (ms * 1e-3) / 1e9, "unit": "GLUP/s",
                       "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                                    "peak_source": peak_src, "traffic": traffic,
                                    "algorithmic_bytes_per_launch": BYTES_PER_LUP_PRESSURE * float(n) * n * 4,
                                    "note": "24 B per cell-update x 4 sweeps x 4096^2 cells per launch (one pass); temporal blocking "
                                            "makes DRAM traffic per launch smaller than the algorithmic bytes"}}
    s.close()
    for scheme, name in ((False, "upwind"), (True, "quick")):
        s = slab.GpuSlab(_ldc_params(n, device, 32, 0.0, scheme_quick=scheme), 1, 0)
        Var, Ff = _synthetic_rows(n, 0, n + 1)
        s.h.upload(Var=Var, VarOld=Var)
        del Var, Ff
        for k in range(3):                                    # face fluxes as the solver's own kernels leave them: an outer
            s.h.k_apply_bc(k)                                 # iteration's momentum solves read what the previous iteration's
        s.h.k_linear_interpolation(); s.h.k_update_flux()     # linear_interpolation + update_flux produced
        sc = capi.SCHEME_QUICK if scheme else capi.SCHEME_UPWIND
        slab.solve_momentum([s], 0, sc)
        s.h.synchronize()
        s.h.timer_start()
        sw, _ = slab.solve_momentum([s], 0, sc)
        ms = s.h.timer_stop()
        lups = float(n) * n * sw
        ach = BYTES_PER_LUP_MOMENTUM * lups / (ms * 1e-3) / 1e9
        e = os.environ.get("SRCFD_SLAB_SWEEP2")                # library default: two sweeps per pass for upwind, one per launch for QUICK
        two = (not scheme) if e is None else (e != "0")
        kern = (f"k_slab_sweep2<{name}, paired fluxes> (two JACOBI sweeps per pass over HBM: register windows per warp, the west flux is the "
                "east flux of the row above; 24 B of DRAM traffic per cell update against 40 algorithmic)") if two else \
               (f"k_slab_sweep<{name}, paired fluxes> (one JACOBI sweep per launch; the west flux is the east flux of the "
                "row above, 48 B of DRAM traffic per cell update against 40 algorithmic)")
        out["momentum_" + name] = {"kernel": kern, "sweeps": sw, "ms": ms,
                                   "value": lups / (ms * 1e-3) / 1e9, "unit": "GLUP/s",
                                   "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak}}
        s.close()
    return out


def slab_record(rank, world, local, dist, torch):
    """BASELINE configs[3]: 4096^2 cavity split into row slabs over the ranks (strong scaling)."""
    from srcfd import slab
    peak, _ = peaks()
    n = LG_N
    halo = int(os.environ.get("SRCFD_BENCH_HALO", "16"))

    def maxtime(ms):
        if world > 1:
            t = torch.tensor([ms], device=f"cuda:{local}", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    def allok(flag):
        if world > 1:
            t = torch.tensor([1.0 if flag else 0.0], device=f"cuda:{local}", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            return bool(t.item() > 0.5)
        return bool(flag)

    out = {"grid": [n, n], "n_gpus": world, "halo_rows": halo if world > 1 else 0, "scaling": "strong",
           "data_plane": "halo rows and residual sums stored by the kernels into cudaIpc-mapped peer mailboxes with "
                         "sequence flags (csrc/slab.cuh); no NCCL call and no host copy between sweeps"}
    # ---- pressure relaxation, 1000 sweeps (the reference's inner cap), random plane
    s = slab.GpuSlab(_ldc_params(n, local, 1000, 0.0), world, rank, halo=halo)
    slab.attach_distributed(s)
    g0, g1 = s.part.global_rows()
    Var, Ff = _synthetic_rows(n, g0, g1)
    s.h.upload(Var=Var, VarOld=Var, Ff=Ff)
    slab.solve_pressure([s])                                  # warm-up
    s.h.upload(Var=Var)                                       # same start for the timed solve and for the parity reference
    s.h.synchronize()
    if world > 1:
        dist.barrier()
    i0 = s.info()
    s.h.timer_start()
    sw, rms = slab.solve_pressure([s])
    ms = maxtime(s.h.timer_stop())
    i1 = s.info()
    lups = float(n) * n * sw
    ach = BYTES_PER_LUP_PRESSURE * lups / (ms * 1e-3) / 1e9
    out["pressure"] = {"sweeps": sw, "ms": ms, "value": lups / (ms * 1e-3) / 1e9, "unit": "GLUP/s", "last_rms": rms,
                       "roofline": {"bound": "hbm", "achieved": ach, "peak": peak * world, "unit": "GB/s", "frac": ach / (peak * world),
                                    "note": "peak = measured single-GPU copy bandwidth x n_gpus"},
                       "exchanges": i1["exchanges"] - i0["exchanges"],
                       "halo_bytes_pushed_per_rank": i1["halo_bytes"] - i0["halo_bytes"]}
    parity = None
    if world > 1:                                             # every rank: the undivided solve on its own GPU, same kernels
        own = s.owned_rows()
        del Var, Ff
        s.close()
        one = slab.GpuSlab(_ldc_params(n, local, 1000, 0.0), 1, 0)
        V1, F1 = _synthetic_rows(n, 0, n + 1)
        one.h.upload(Var=V1, VarOld=V1, Ff=F1)
        slab.solve_pressure([one])                            # warm-up (and the kernel-choice probe), as for the slabs
        one.h.upload(Var=V1)
        del V1, F1
        one.h.synchronize()
        one.h.timer_start()
        sw1, _ = slab.solve_pressure([one])
        ms1 = one.h.timer_stop()
        ref = one.owned_rows()
        one.close()
        a = s.part.own0 - 1
        parity = allok(sw1 == sw and np.array_equal(own, ref[a:a + s.part.n_own]))
        out["pressure"]["single_gpu_same_run"] = {"ms": ms1, "value": lups / (ms1 * 1e-3) / 1e9, "unit": "GLUP/s",
                                                  "note": "the undivided plane on this rank's GPU with the kernel a single GPU would use "
                                                          "(streaming kernel on a roomy plane); the slabs pick tiles or streaming by their height"}
        out["pressure"]["speedup_vs_single_gpu_same_run"] = ms1 / ms
        out["pressure"]["strong_scaling_efficiency_same_run"] = ms1 / ms / world
    else:
        del Var, Ff
        s.close()
    out["pressure"]["slab_parity"] = parity
    # ---- the same relaxation with the work per GPU held fixed: a (4096 x N) x 4096 plane, 4096 rows per rank (weak scaling).
    # 4096^2 over 8 ranks leaves 2.2 M cells = ~35 us of work per rank and pass, the size of the launch / hand-off latencies;
    # this record shows the exchange itself does not cost throughput when a rank has a GPU's worth of rows.
    if world > 1:
        s = slab.GpuSlab(_ldc_params(n, local, 1000, 0.0, nx=n * world), world, rank, halo=halo)
        slab.attach_distributed(s)
        g0, g1 = s.part.global_rows()
        Var, Ff = _synthetic_rows(n, g0, g1, nx=n * world)
        s.h.upload(Var=Var, VarOld=Var, Ff=Ff)
        del Var, Ff
        slab.solve_pressure([s])
        s.h.synchronize()
        dist.barrier()
        s.h.timer_start()
        sww, _ = slab.solve_pressure([s])
        msw = maxtime(s.h.timer_stop())
        s.close()
        vw = float(n) * n * world * sww / (msw * 1e-3) / 1e9
        out["pressure_weak"] = {"grid": [n * world, n], "rows_per_gpu": n, "sweeps": sww, "ms": msw, "value": vw, "unit": "GLUP/s",
                                "scaling": "weak", "per_gpu": vw / world,
                                "roofline": {"bound": "hbm", "achieved": BYTES_PER_LUP_PRESSURE * vw, "peak": peak * world, "unit": "GB/s",
                                             "frac": BYTES_PER_LUP_PRESSURE * vw / (peak * world)}}
    # ---- configs[3] proper: 10 outer iterations of the Re=1000 cavity (QUICK, zero start, inner tol 1e-6 / cap 1000)
    its = 10
    s = slab.GpuSlab(_ldc_params(n, local, 1000, 1e-6), world, rank, halo=halo)
    slab.attach_distributed(s)
    s.h.initialize_fields(True)
    slab.step([s], 1, (0.0, 0.0, 0.0))                        # warm-up iteration (also pages everything in)
    s.h.initialize_fields(True)
    s.h.reset_counters()
    s.h.synchronize()
    if world > 1:
        dist.barrier()
    s.h.timer_start()
    slab.step([s], its, (0.0, 0.0, 0.0))
    ms = maxtime(s.h.timer_stop())
    st = s.h.status()
    sweeps = [int(x) for x in st["total_sweeps"]]
    out["outer_iterations"] = {"case": "lid-driven cavity Re=1000, QUICK, dt=1e-3, zero start, inner tol 1e-6 / cap 1000",
                               "iterations": its, "ms_per_iteration": ms / its, "inner_sweeps_u_v_p": sweeps,
                               "value": float(n) * n * float(sum(sweeps)) / (ms * 1e-3) / 1e9, "unit": "GLUP/s",
                               "rms_u_v_p": [float(x) for x in st["rms"]], "replays": s.info()["replays"]}
    own = s.owned()[0] if world > 1 else None                # (the parity check below compares the zero-start result)
    # ---- the same iteration on a field without a front (every cell O(1), as after an SR warm start: no band of denormal values,
    # so the pressure solves run on the streaming kernel and into the 1000-sweep cap): whole outer iterations in the HBM-bound regime
    try:
        g0, g1 = s.part.global_rows()
        xi = (np.arange(g0, g1 + 1, dtype=np.float64) / (n + 1))[:, None]
        yj = (np.arange(n + 2, dtype=np.float64) / (n + 1))[None, :]
        Var = np.empty((3, g1 - g0 + 1, n + 2))
        Var[0] = 0.1 * np.sin(np.pi * xi) * np.cos(2 * np.pi * yj) * np.sin(np.pi * yj)      # a smooth divergence-light swirl
        Var[1] = -0.1 * np.cos(np.pi * xi) * np.sin(np.pi * xi) * np.sin(2 * np.pi * yj)
        Var[2] = 0.05 * np.cos(2 * np.pi * xi) * np.cos(2 * np.pi * yj)
        s.h.upload(Var=Var, VarOld=Var)
        del Var, xi, yj
        for k in range(3):
            s.h.k_apply_bc(k)
        s.h.k_linear_interpolation()
        slab.step([s], 2, (0.0, 0.0, 0.0))                    # warm-up: the first pressure solve is the tile-kernel probe
        s.h.reset_counters()
        s.h.synchronize()
        if world > 1:
            dist.barrier()
        its_w = 5
        s.h.timer_start()
        slab.step([s], its_w, (0.0, 0.0, 0.0))
        msw2 = maxtime(s.h.timer_stop())
        stw = s.h.status()
        sww2 = [int(x) for x in stw["total_sweeps"]]
        ks = s.kernel_stats()
        out["outer_iterations_developed"] = {
            "case": "as outer_iterations, started from a smooth analytic field without a front (a swirl of amplitude 0.1, p of amplitude 0.05; a function of the global cell index)",
            "iterations": its_w, "ms_per_iteration": msw2 / its_w, "inner_sweeps_u_v_p": sww2,
            "value": float(n) * n * float(sum(sww2)) / (msw2 * 1e-3) / 1e9, "unit": "GLUP/s",
            "pressure_kernel_solves": {"streaming": int(ks["stream_solves"]), "tiles": int(ks["tile_solves"])},
            "note": "throughput record; bitwise parity of the decomposed iteration is asserted on the zero-start record above"}
    except Exception as e:                                    # a sub-record never takes the bench line down
        out["outer_iterations_developed"] = {"error": repr(e)}
    parity2 = None
    if world > 1:
        s.close()
        one = slab.GpuSlab(_ldc_params(n, local, 1000, 1e-6), 1, 0)
        one.h.initialize_fields(True)
        slab.step([one], 1, (0.0, 0.0, 0.0))
        one.h.initialize_fields(True)
        one.h.reset_counters()
        one.h.timer_start()
        slab.step([one], its, (0.0, 0.0, 0.0))
        ms1 = one.h.timer_stop()
        st1 = one.h.status()
        ref = one.owned()[0]
        one.close()
        a = s.part.own0 - 1
        parity2 = allok([int(x) for x in st1["total_sweeps"]] == sweeps and np.array_equal(own, ref[:, a:a + s.part.n_own]))
        out["outer_iterations"]["single_gpu_same_run_ms_per_iteration"] = ms1 / its
        out["outer_iterations"]["speedup_vs_single_gpu_same_run"] = ms1 / ms
    else:
        s.close()
    out["outer_iterations"]["slab_parity"] = parity2
    out["slab_parity"] = None if world == 1 else bool(parity and parity2)
    return out


def ensemble_record(rank, world, local, dist, torch):
    """BASELINE configs[2]: the multiBC Re sweep (14 Re x {single, double lid} + 3 BFS = 31 cases, 400x400, SR-warm-started)
    dealt round-robin to the ranks, fixed budget of outer iterations per case; strong scaling, no data-path collective."""
    from srcfd import ensemble as E, sr, bfs, ldc
    its = int(os.environ.get("SRCFD_BENCH_ENSEMBLE_ITS", "60"))
    bfs._wf.verbose = ldc._wf.verbose = False
    cases = E.multibc_sweep(max_iterations=its)
    mine = E.shard_cases(cases, rank, world)
    sr.default_device = local
    files = dict(stats=os.path.join(GOLDEN, "stats_10to400_multiBC.txt"), encoder=sr.load_model(os.path.join(GOLDEN, "encoder10_multiBC.h5")),
                 decoder=sr.synthetic_decoder(seed=0), coarse_iterations=2000)
    E.run_local(mine[:1], device=local, concurrency=1, sr_files=files, keep_fields=False)     # context / code warm-up
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    warm = E.warm_stage(mine, files, device=local)
    t1 = time.perf_counter()
    res = E.run_local(mine, device=local, concurrency=4, sr_files=files, keep_fields=False, warm_fields=warm)
    t2 = time.perf_counter()
    lups = float(sum(int(np.sum(r.total_sweeps)) * r_nx * r_ny for r, (r_nx, r_ny) in zip(res, [(c.nx, c.ny) for c in mine])))
    v = torch.tensor([t1 - t0, t2 - t1, t2 - t0], dtype=torch.float64, device=f"cuda:{local}")
    w = torch.tensor([lups, float(len(res))], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(v, op=dist.ReduceOp.MAX)
        dist.all_reduce(w, op=dist.ReduceOp.SUM)
    warm_s, fine_s, tot_s = v.tolist()
    lups, ncase = w.tolist()
    return {"cases": int(ncase), "grid": [NX, NY], "outer_iterations_per_case": its, "scaling": "strong",
            "concurrent_cases_per_gpu": 4, "warm_stage_s": warm_s, "fine_stage_s": fine_s,
            "warm_stage": "one k_coarse_solve launch for the rank's 10x10 coarse solves (2000 its each) + one batched SR call per case family",
            "value": lups / fine_s / 1e9, "unit": "GLUP/s", "value_incl_warm_stage": lups / tot_s / 1e9,
            "timing": "host wall clock, barrier before, max over ranks"}


def time_to_converged_record(device):
    """The second half of BASELINE's metric on the smaller of the two cases whose timings the reference publishes
    (stored stdout of sr-simulation-data-creation.ipynb), in two sweep orders: the reference's
    own (results bit-identical to the single-thread reference) and RB_JACOBI (Jacobi momentum + red-black pressure, the
    orders north_star names: same criterion, same fixed point, no wavefront latency)."""
    from srcfd import ldc
    n = 100
    runs, fields = {}, {}
    for order in ("GS_LEX", "RB_JACOBI"):
        bc = ldc.BoundaryConditions()
        bc.u_boundaries['bottom'] = ldc.BoundaryCondition('dirichlet', 1.0)          # double lid: the notebook's default
        s = ldc.CFDSolver(ldc.MeshParameters(nx=n, ny=n), ldc.FluidProperties(Re=1050.0),
                          ldc.SolverSettings(dt=1e-3, scheme="QUICK", max_iterations=150000, sweep_order=order), bc, device=device)
        t0 = time.perf_counter()
        its, _ = s.solve("x", verbose=False, save=False)
        dt = time.perf_counter() - t0
        fields[order] = s.Var.copy()
        runs[order] = {"converged": bool(s.converged), "outer_iterations": int(its), "seconds": dt, "ms_per_iteration": 1e3 * dt / its,
                       "inner_sweeps_u_v_p": [int(x) for x in s.total_sweeps],
                       "value": n * n * float(np.sum(s.total_sweeps)) / dt / 1e9, "unit": "GLUP/s", "speedup_vs_published": 212.41 / dt}
    rel = []
    for k in range(3):
        a, b = fields["GS_LEX"][k, 1:-1, 1:-1], fields["RB_JACOBI"][k, 1:-1, 1:-1]
        if k == 2:
            a, b = a - a.mean(), b - b.mean()                  # pressure is defined up to a constant (all-Neumann cavity)
        rel.append(float(np.linalg.norm(a - b) / np.linalg.norm(a)))
    return {"case": f"double-lid cavity Re=1050 {n}x{n}, QUICK, dt=1e-3, zero start, criterion 1e-6 on u, v, p",
            "timing": "host wall clock around CFDSolver.solve()",
            "reference_order": runs["GS_LEX"], "rb_jacobi_order": runs["RB_JACOBI"],
            "seconds": runs["GS_LEX"]["seconds"], "speedup_vs_published": runs["GS_LEX"]["speedup_vs_published"],
            "relL2_between_orders_u_v_p": rel,
            "relL2_note": "both runs stop on the same 1e-6 criterion; the distance between the two stopped iterates is of the size "
                          "of the reference's own thread-count scatter at that criterion (3-4e-5, SURVEY hazard H1)",
            "reference_published": {"outer_iterations": 80012, "seconds": 212.41, "hardware": "Kaggle CPU notebook (core count not recorded)",
                                    "source": "sr-simulation-data-creation.ipynb raw 8420",
                                    "note": "multi-threaded numba: racy sweep order, iteration count not reproducible"}}


DEC_FLOP_PER_SAMPLE = 2 * (50 * 36864 + 144 * 256 * 9 * 128 + 625 * 128 * 256 + 2500 * 64 * 128 + 10000 * 32 * 64 + 40000 * 16 * 32 + 160000 * 72)


def decoder_record(device, torch):
    """BASELINE configs[4]: decoder_400 on 1024 latents resident in HBM -> 1024 x (400,400,1) fp32 fields."""
    from srcfd import sr
    B = 1024
    tf_peak = None
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            tf_peak = json.load(f).get("bf16_tflops_sustained")
    tf_peak = tf_peak or 1400.0
    hbm, _ = peaks()
    dec = sr.synthetic_decoder(0)
    z = torch.randn(B, 50, device=f"cuda:{device}", dtype=torch.float32, generator=torch.Generator(f"cuda:{device}").manual_seed(0))
    o = torch.empty(B, 400, 400, 1, device=f"cuda:{device}", dtype=torch.float32)
    res = {}
    for mode in ("fp32", "bf16x3", "bf16"):
        sr.set_precision(mode)
        for _ in range(2):
            sr.decode_device(dec, z.data_ptr(), 64, o.data_ptr())
        ms = min(sr.decode_device(dec, z.data_ptr(), B, o.data_ptr()) for _ in range(3))
        res[mode] = {"ms": ms, "samples_per_s": B / (ms * 1e-3), "tflops": B * DEC_FLOP_PER_SAMPLE / (ms * 1e-3) / 1e12,
                     "output_write_gbs": B * 640000 / (ms * 1e-3) / 1e9}
    sr.set_precision("bf16x3")
    torch.cuda.synchronize()
    tf = res["bf16x3"]["tflops"]
    return {"metric": "SR decoder inference throughput", "batch": B, "unit": "samples/s",
            "value": res["bf16x3"]["samples_per_s"],
            "value_path": "bf16x3: every ConvT on tcgen05 as three bf16 MMAs per K-step (a_hi*w_hi + a_hi*w_lo + a_lo*w_hi), fp32 activations, "
                          "epilogues and final conv; the last ConvT, its swish and the final 3x3 conv are ONE kernel (the 400x400x8 "
                          "activation between them stays in shared memory) -- the library's default; held to the fp32 path's tolerance "
                          "(rtol 1e-4 / atol 5e-5 against the numpy restatement, tests/test_gpu_sr.py)",
            "paths": res, "tc_error": bool(sr.tc_error()),
            "bf16_note": "bf16 operands AND activations on tcgen05: within 3e-2 of the output range of the fp32 restatement -- narrower "
                         "arithmetic than the reference's fp32, reported beside the value, not as the value",
            "parity": "unpinned (no decoder weights / TensorFlow / input-output pair in the reference tree): synthetic seed-0 weights",
            "roofline": {"bound": "tensor", "achieved": tf, "peak": tf_peak, "unit": "TFLOP/s", "frac": tf / tf_peak,
                         "note": "useful FLOPs (276 MFLOP per sample) / time; the split path issues 3x the MMAs",
                         "hbm_output_frac": res["bf16x3"]["output_write_gbs"] / hbm}}


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
    from srcfd import sr as _sr, kernels as _kernels
    _sr.default_device = _kernels.default_device = local     # every library object of this rank lives on its own GPU

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
    e2e_steps = max(2, min(args.steps, 100))
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

    # ---- sub-records (every rank takes part in the slab one) ---------------------------------------
    del flush
    torch.cuda.empty_cache()
    extras = {}
    if not args.no_extras:
        slab_rec = slab_record(rank, world, local, dist, torch)
        ens_rec = ensemble_record(rank, world, local, dist, torch)
        if rank == 0:
            extras["slab"] = slab_rec
            extras["ensemble"] = ens_rec
        if world == 1:
            extras["large_grid"] = large_grid_record(local)
            extras["decoder"] = decoder_record(local, torch)
            if not args.no_ttc:
                extras["time_to_converged"] = time_to_converged_record(local)

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
    # latency model of the full-height wavefront kernel: groups of K sweeps follow each other by a
    # fixed lag; a launch = (groups - 1) x lag + one crossing of the plane
    n_launch = max(1, tr["pressure_launches"])
    sw_launch = float(sweeps[2]) / n_launch
    K3, lag_us, step_us = 3, 4.1, 0.48
    groups = int(np.ceil(sw_launch / K3))
    steps = NX + NY + 2 * K3 + 1
    model_ms = ((groups - 1) * lag_us + steps * step_us) * 1e-3
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
        "roofline": {"bound": "latency", "kernel": "k_solve_gs3 (solve_pressure inner loop: full-height sweep groups, reference sweep order)",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "peak_source": peak_src, "traffic": ncu_traffic(),
                     "algorithmic_bytes_per_launch": BYTES_PER_LUP_PRESSURE * p_lups / n_launch,
                     "launches": tr["pressure_launches"], "avg_launch_ms": p_ms / n_launch,
                     "share_of_step": p_ms / t_ms if t_ms else None,
                     "latency_model": {"groups_per_launch": groups, "sweeps_per_group": K3, "lag_us_per_group": lag_us,
                                       "steps_per_crossing": steps, "us_per_step": step_us, "model_ms": model_ms,
                                       "measured_over_model": (p_ms / n_launch) / model_ms if model_ms else None,
                                       "source": "per-task device timestamps, tools/trace_gs3.py (DESIGN.md section 4.1)"},
                     "note": "the 400^2 planes (1.3 MB) never leave L2/SMEM, so HBM does not bound this kernel: `achieved`/`frac` are an "
                             "algorithmic-bytes rate (24 B per cell-update), the binding limit is the dependency chain of the in-place "
                             "sweep order (latency_model).  The HBM-bound regime is large_grid.pressure.roofline (4096^2); the >= 0.70 "
                             "target is judged there",
                     "hbm_frac_large_grid": (extras.get("large_grid", {}).get("pressure", {}).get("roofline", {}).get("frac"))},
        "clocks": clocks,
    }
    line.update(extras)
    if cb:
        line["cpu_baseline"] = cb
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extras", action="store_true", help="skip the sub-records (large_grid, slab, decoder, time_to_converged)")
    ap.add_argument("--no-ttc", action="store_true", help="skip the time_to_converged sub-record (about 25 s)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
