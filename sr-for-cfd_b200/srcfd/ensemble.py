"""Re/BC parameter-sweep ensembles (BASELINE.json configs[2]): independent cases sharded over GPUs.

The reference runs its sweep as a serial Python loop (sr-simulation-data-creation.ipynb cell-2 lines
744-794).  Cases are fully independent, so the multi-GPU path has NO data-path collective: the case list
is dealt round-robin to the ranks (one process per GPU), each rank runs its share -- several cases
concurrently on its GPU, each on its own stream with a capped persistent grid -- and rank 0 gathers the
per-case results at the end.
"""
from __future__ import annotations

import threading
import time
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np


@dataclass
class CaseSpec:
    """One member of the sweep.  kind: 'ldc' (single lid), 'ldc2' (double lid, u_bottom = 1) or 'bfs'."""
    kind: str
    Re: float
    nx: int = 400
    ny: int = 400
    max_iterations: int = 2000
    warm_start: bool = True
    name: str = ""

    def label(self):
        return self.name or f"{self.kind}_Re{self.Re:g}_{self.nx}x{self.ny}"


@dataclass
class CaseResult:
    label: str
    rank: int
    iterations: int
    converged: bool
    seconds: float
    total_sweeps: List[int]
    rms: List[float]
    fields: Optional[np.ndarray] = None       # (3, ny, nx) u, v, p in the reference's output orientation
    phases: Optional[Dict[str, float]] = None  # seconds spent constructing the solver, injecting the warm start, solving


def multibc_sweep(res_ldc=tuple(range(50, 701, 50)), res_bfs=(100, 200, 400), nx=400, ny=400, max_iterations=2000):
    """The multiBC ensemble of SURVEY.md section 8d config 3."""
    cases = [CaseSpec(k, float(Re), nx, ny, max_iterations) for Re in res_ldc for k in ("ldc", "ldc2")]
    cases += [CaseSpec("bfs", float(Re), nx, ny, max_iterations) for Re in res_bfs]
    return cases


def shard_cases(cases: Sequence, rank: int, world: int) -> List:
    """Static round-robin: case i belongs to rank i % world (expected cost is similar across the sweep)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return [c for i, c in enumerate(cases) if i % world == rank]


def _case_bc(spec: CaseSpec):
    from . import ldc
    if spec.kind != "ldc2":
        return None
    bc = ldc.BoundaryConditions()
    bc.u_boundaries['bottom'] = ldc.BoundaryCondition('dirichlet', 1.0)     # PyCFD_ML_accelerated.py:1387-1392
    return bc


def coarse_stage(cases: Sequence[CaseSpec], device: int = 0, lr_dim: int = 10, max_iterations: int = 2000) -> List[dict]:
    """run_coarse_simulation (LDC.py:696-761 / BFS.py:893-977) of every case in ONE launch: one CTA per case
    (srcfd_coarse_solve_batch), each bit-identical to its own CFDSolver(lr_dim x lr_dim).solve().
    Returns the reference's coarse field dicts {'u','v','p'} of shape (lr_dim, lr_dim), in case order."""
    from . import _capi as capi, bfs, ldc, solver as S
    if not cases:
        return []
    params, crit = [], []
    for spec in cases:
        wf = (bfs if spec.kind == "bfs" else ldc)._wf
        dt, scheme, lx, ly = wf._defaults(None, None, None, None)
        s_mesh = S.MeshParameters(nx=lr_dim, ny=lr_dim, lx=lx, ly=ly)
        cc = {'u': 1e-6, 'v': 1e-6, 'p': 1e-6, 'continuity': 1e-6}
        if wf.bfs:
            st = S.BFSSolverSettings(dt=dt, scheme=scheme, max_iterations=max_iterations, convergence_criteria=cc)
            params.append(S.make_params(s_mesh, S.FluidProperties(Re=spec.Re, rho=1.0), st, wf._default_bc(), 'BFS',
                                        1.0, 2.0, 1.0, True, device))
        else:
            st = S.SolverSettings(dt=dt, scheme=scheme, max_iterations=max_iterations, convergence_criteria=cc)
            params.append(S.make_params(s_mesh, S.FluidProperties(Re=spec.Re, rho=1.0), st,
                                        _case_bc(spec) or wf._default_bc(), None, device=device))
        crit.append((cc['u'], cc['v'], cc['p']))
    r = capi.coarse_solve_batch(params, max_iterations, np.array(crit))
    if r['nan'].any():
        raise ValueError("Solver failed: NaN/Inf in residuals")
    return [{n: r['Var'][i, k, 1:-1, 1:-1].T.copy() for k, n in enumerate('uvp')} for i in range(len(cases))]


def warm_start_fields(spec: CaseSpec, sr_files: dict, coarse: Optional[dict] = None) -> np.ndarray:
    """Steps 1-2 of the per-case workflow (coarse solve unless given, then ml_super_resolution): the (3, ny, nx)
    float32 initial guess of one case."""
    from . import bfs, ldc
    wf = (bfs if spec.kind == "bfs" else ldc)._wf
    if coarse is None:
        coarse = wf.run_coarse_simulation(Re=spec.Re, lr_dim=10, max_iterations=sr_files.get("coarse_iterations", 2000),
                                          bc=_case_bc(spec), save=False)
    kw = dict(use_aspect_ratio_correction=True, lx=10.0, ly=3.0) if spec.kind == "bfs" else {}
    hr = wf.ml_super_resolution(coarse, 10, spec.nx, sr_files["stats"], sr_files["encoder"], sr_files["decoder"], **kw)
    return np.stack([np.asarray(hr[c], dtype=np.float32) for c in "uvp"])


def run_case(spec: CaseSpec, device: int = 0, max_ctas: int = 0, sr_files: Optional[dict] = None, keep_fields=True,
             coarse: Optional[dict] = None, warm: Optional[np.ndarray] = None, pool: Optional[dict] = None) -> CaseResult:
    """Coarse solve -> SR warm start -> fine solve for one case on one GPU (the reference's per-case workflow).
    `coarse`: this case's coarse fields when coarse_stage already produced them; `warm`: its finished initial guess;
    `pool`: a worker's solvers by (family, grid) -- a sweep re-parameterises one solver per family (Re, BCs, budget:
    srcfd_set_params) instead of allocating page-locked arrays and device state for every case, which cost as much as
    30 outer iterations of a 400x400 case; the start state is re-made by _initialize_fields exactly as a new solver's."""
    from . import bfs, ldc, solver as S
    mod = bfs if spec.kind == "bfs" else ldc
    wf = mod._wf
    t0 = time.time()
    bc = _case_bc(spec)
    key = (spec.kind == "bfs", spec.nx, spec.ny, device, max_ctas)
    solver = pool.get(key) if pool is not None else None
    if solver is None:
        solver = wf._make_solver(spec.Re, spec.nx, spec.ny, *wf._defaults(None, None, None, None)[:2], None,
                                 spec.max_iterations, bc, 1.0, 2.0, 1.0, *wf._defaults(None, None, None, None)[2:], None,
                                 device=device, max_ctas=max_ctas)
        if pool is not None:
            pool[key] = solver
    else:
        solver.fluid = S.FluidProperties(Re=spec.Re, rho=1.0)
        solver.bc = bc or wf._default_bc()
        solver.settings.max_iterations = spec.max_iterations
        solver.residual_history = {'u': [], 'v': [], 'p': []}
        solver._initialize_fields()
    t1 = time.time()
    if warm is None and spec.warm_start and sr_files is not None:
        warm = warm_start_fields(spec, sr_files, coarse)
    if warm is not None:
        solver._sync_params()
        solver._handle.set_fields(warm)
        solver._handle.download(solver.Var, solver.VarOld, solver.Ff)
    t2 = time.time()
    n, _ = solver.solve("ensemble", verbose=False, save=False)
    t3 = time.time()
    fields = np.stack([solver.Var[k, 1:-1, 1:-1].T for k in range(3)]) if keep_fields else None
    return CaseResult(spec.label(), -1, int(n), bool(solver.converged), time.time() - t0,
                      [int(x) for x in solver.total_sweeps], [float(x) for x in solver.last_rms], fields,
                      dict(construct=t1 - t0, warm=t2 - t1, solve=t3 - t2))


def warm_stage(cases: Sequence[CaseSpec], sr_files: dict, device: int = 0) -> List[Optional[np.ndarray]]:
    """Warm start of a whole share of cases: ONE launch for every coarse solve (coarse_stage), then ONE batched SR call
    per case family (3 fields per case; statistics, both networks and the clean-up on the device)."""
    out: List[Optional[np.ndarray]] = [None] * len(cases)
    warm = [i for i, c in enumerate(cases) if c.warm_start]
    coarse = coarse_stage([cases[i] for i in warm], device, 10, sr_files.get("coarse_iterations", 2000))
    # one batched SR call per case family (the BFS cases share the aspect-ratio and adaptive-statistics settings)
    from . import bfs, ldc
    for kind_is_bfs, wf in ((False, ldc._wf), (True, bfs._wf)):
        idx = [(i, f) for i, f in zip(warm, coarse) if (cases[i].kind == "bfs") == kind_is_bfs]
        if not idx:
            continue
        by_dim = {}
        for i, f in idx:
            by_dim.setdefault(cases[i].nx, []).append((i, f))
        for nx, grp in by_dim.items():
            kw = dict(use_aspect_ratio_correction=True, lx=10.0, ly=3.0) if kind_is_bfs else {}
            hr = wf.ml_super_resolution_batch([f for _, f in grp], 10, nx, sr_files["stats"], sr_files["encoder"],
                                              sr_files["decoder"], **kw)
            for (i, _), h in zip(grp, hr):
                out[i] = np.stack([np.asarray(h[c], dtype=np.float32) for c in "uvp"])
    return out


def run_local(cases: Sequence[CaseSpec], device: int = 0, concurrency: int = 4, num_sms: int = 148,
              runner: Callable = run_case, warm_fields: Optional[List] = None, reuse_solvers: bool = True,
              **kw) -> List[CaseResult]:
    """Run this rank's cases, `concurrency` at a time, each with 1/concurrency of the SMs.  `warm_fields`: the cases'
    initial guesses when warm_stage ran already."""
    results: List[Optional[CaseResult]] = [None] * len(cases)
    lock, nxt = threading.Lock(), [0]
    max_ctas = max(1, num_sms // max(1, concurrency)) if concurrency > 1 else 0
    sr_files = kw.get("sr_files")
    if warm_fields is None:
        warm_fields = [None] * len(cases)
        if runner is run_case and sr_files is not None and any(c.warm_start for c in cases):
            warm_fields = warm_stage(cases, sr_files, device)      # the workers below only run fine solves

    def worker():
        pool: Dict = {}                                  # this worker's solvers, one per case family (run_case)
        while True:
            with lock:
                i = nxt[0]
                nxt[0] += 1
            if i >= len(cases):
                return
            extra = dict(warm=warm_fields[i]) if warm_fields[i] is not None else {}
            if runner is run_case and reuse_solvers:
                extra["pool"] = pool
            results[i] = runner(cases[i], device=device, max_ctas=max_ctas, **extra, **kw)

    threads = [threading.Thread(target=worker) for _ in range(max(1, min(concurrency, len(cases))))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    return [r for r in results if r is not None]


def run_ensemble(cases: Sequence[CaseSpec], concurrency: int = 4, runner: Callable = run_case, dist=None,
                 device: Optional[int] = None, **kw) -> Optional[List[CaseResult]]:
    """Shard `cases` over the ranks of an initialised torch.distributed group (or run them all when there is
    none) and gather every CaseResult on rank 0 (other ranks return None).  No collective on the data path."""
    rank, world = (dist.get_rank(), dist.get_world_size()) if dist is not None and dist.is_initialized() else (0, 1)
    mine = shard_cases(list(cases), rank, world)
    res = run_local(mine, device=rank if device is None else device, concurrency=concurrency, runner=runner, **kw)
    for r in res:
        r.rank = rank
    if world == 1:
        return res
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(res, gathered, dst=0)
    if rank != 0:
        return None
    order = {c.label(): i for i, c in enumerate(cases)}
    return sorted([r for part in gathered for r in part], key=lambda r: order[r.label])
