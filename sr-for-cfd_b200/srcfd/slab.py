"""Slab domain decomposition of the pressure relaxation across GPUs (BASELINE configs[3]: a large grid split along i).

Scope: the JACOBI-order pressure solve (the order BASELINE's north star names for the decomposed grid), i.e.
solve_pressure (PyCFD_ML_accelerated.py:292-314) with every cell of a sweep computed from the previous iterate.  One
process per GPU; rank r owns a contiguous block of interior rows.  The temporally blocked kernel advances H sweeps per
pass; a slab carries M*H halo rows on each side that has a neighbour and exchanges them once per M passes (M*H sweeps),
not per sweep: stale information from beyond the halo travels one row per sweep, so after M*H sweeps it has not reached
an owned row.  Per pass: one kernel pass and one all-reduce of the H per-sweep residual sums (in place on the device);
per M passes: one halo exchange (NCCL send/recv of M*H contiguous rows each way).  The break rule ("stop after the first sweep with rms < tol", LDC.py:310-313) is applied to the
globally reduced sums BEFORE a pass is committed; if an earlier sweep of the pass met the tolerance the pass is
repeated with fewer sweeps, so the result is the single-domain Jacobi result bit for bit.

The driver below is back-end agnostic (pass / commit / exchange / all-reduce are callables): tests run it on the CPU
with the oracle's arithmetic under gloo, the product binds it to libsrcfd (GpuSlab) and torch.distributed/NCCL.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Optional, Tuple

import numpy as np


@dataclass
class SlabPartition:
    """Rows 1..nx dealt to `world` ranks in contiguous blocks; halo rows only towards existing neighbours."""
    nx: int
    world: int
    rank: int
    halo: int

    def __post_init__(self):
        if self.world < 1 or not (0 <= self.rank < self.world):
            raise ValueError("bad rank/world")
        base, rem = divmod(self.nx, self.world)
        if base < max(1, self.halo) and self.world > 1:
            raise ValueError(f"{self.nx} rows over {self.world} ranks leaves slabs thinner than the halo ({self.halo})")
        counts = [base + (1 if r < rem else 0) for r in range(self.world)]
        self.own0 = 1 + sum(counts[: self.rank])           # first owned global row (1-based)
        self.n_own = counts[self.rank]
        self.own1 = self.own0 + self.n_own - 1
        self.lo = self.halo if self.rank > 0 else 0         # halo rows above (towards rank-1)
        self.hi = self.halo if self.rank < self.world - 1 else 0
        self.nx_local = self.lo + self.n_own + self.hi      # interior rows of the local grid

    # local plane rows: 0 = boundary row, 1..nx_local interior, nx_local+1 = boundary row
    @property
    def local_own0(self) -> int:
        return self.lo + 1

    @property
    def local_own1(self) -> int:
        return self.lo + self.n_own

    def global_rows(self) -> Tuple[int, int]:
        """Global row indices of local rows 0 and nx_local+1 (inclusive range of the local plane)."""
        g0 = self.own0 - self.lo - 1
        return g0, g0 + self.nx_local + 1

    def take(self, plane: np.ndarray) -> np.ndarray:
        """Local copy (nx_local+2, ny+2) of a global (nx+2, ny+2) plane."""
        g0, g1 = self.global_rows()
        return np.ascontiguousarray(plane[g0:g1 + 1])


def slab_jacobi_solve(part: SlabPartition, ncells_global: int, tol: float, max_iter: int,
                      run_pass: Callable[[int, bool], np.ndarray], commit: Callable[[], None],
                      exchange: Callable[[], None], allreduce_sum: Callable[[np.ndarray], np.ndarray],
                      sweeps_per_pass: Optional[int] = None) -> Tuple[int, float]:
    """The loop every rank runs.  run_pass(nsw, commit_now) -> per-sweep sums of R^2 over the OWNED rows (length nsw);
    exchange() refreshes the halo rows from the neighbours' owned rows; returns (sweeps, last rms).
    sweeps_per_pass (H) may be smaller than the halo depth: the halos are then exchanged every halo // H passes (stale
    data advances one row per sweep, so halo rows of depth M*H protect the owned rows for M passes)."""
    H = sweeps_per_pass or (part.halo if part.world > 1 else max(1, part.halo))
    budget = 0                                               # sweeps the current halo contents are still good for
    n, rms_last = 0, 0.0
    while n < max_iter:
        nsw = min(H, max_iter - n)
        if part.world > 1 and budget < nsw:
            exchange()
            budget = part.halo
        budget -= nsw
        sums = allreduce_sum(np.asarray(run_pass(nsw, False), dtype=np.float64))
        rms = np.sqrt(sums / float(ncells_global))
        hit = np.nonzero(rms < tol)[0]
        if hit.size and int(hit[0]) != nsw - 1:
            first = int(hit[0])
            sums = allreduce_sum(np.asarray(run_pass(first + 1, True), dtype=np.float64))      # plane was untouched: redo shorter
            rms = np.sqrt(sums / float(ncells_global))
            return n + first + 1, float(rms[first])
        commit()
        n += nsw
        rms_last = float(rms[nsw - 1])
        if hit.size:
            break
    return n, rms_last


def slab_jacobi_solve_blocks(part: SlabPartition, ncells_global: int, tol: float, max_iter: int, be, H: int, M: int
                             ) -> Tuple[int, float]:
    """Same result, fewer synchronisations: blocks of up to M passes (H sweeps each) run back to back and commit as
    they go; their M*H residual sums are reduced over the ranks ONCE per block.  A block is speculative: the plane
    (halos included) is saved first, and if some sweep inside the block met the tolerance the block is rolled back and
    replayed up to exactly that sweep.  Needs halos of depth >= M*H.  Back-end `be`: exchange(), snapshot(), restore(),
    run_pass(nsw, slot) (commits), reduce(nslots) -> (nslots, >=H) globally summed per-sweep sums of R^2."""
    if part.world > 1 and part.halo < M * H:
        raise ValueError(f"halo depth {part.halo} < {M} passes x {H} sweeps")
    n, rms_last = 0, 0.0
    while n < max_iter:
        plan, rem = [], max_iter - n
        while rem > 0 and len(plan) < M:
            plan.append(min(H, rem)); rem -= plan[-1]
        if part.world > 1:
            be.exchange()
        be.snapshot()
        for m, nsw in enumerate(plan):
            be.run_pass(nsw, m)
        S = be.reduce(len(plan))
        hit = None
        for m, nsw in enumerate(plan):
            r = np.sqrt(S[m, :nsw] / float(ncells_global))
            idx = np.nonzero(r < tol)[0]
            if idx.size:
                hit = (m, int(idx[0]), float(r[idx[0]]))
                break
        if hit is None:
            n += sum(plan)
            rms_last = float(np.sqrt(S[len(plan) - 1, plan[-1] - 1] / float(ncells_global)))
            continue
        m_, t_, r_ = hit
        total = n + sum(plan[:m_]) + t_ + 1
        if not (m_ == len(plan) - 1 and t_ == plan[-1] - 1):       # overshoot: roll back, replay up to that sweep
            be.restore()
            for m in range(m_):
                be.run_pass(plan[m], m)
            be.run_pass(t_ + 1, m_)
        return total, r_
    return n, rms_last


# ---------------------------------------------------------------------------------------------------------------------
# GPU back-end: one libsrcfd handle per rank on the local grid, torch.distributed for halos and the all-reduce
# ---------------------------------------------------------------------------------------------------------------------
class _DevRows:
    """H contiguous rows of the library's pressure plane as a CUDA array (no copy): torch.as_tensor(obj) aliases them."""

    def __init__(self, ptr: int, rows: int, pitch: int):
        self.__cuda_array_interface__ = {"shape": (rows, pitch), "typestr": "<f8", "data": (int(ptr), False), "version": 3,
                                         "strides": None}


class GpuSlab:
    """Pressure plane of one slab on one GPU.  p_global / Ff_global are the full-domain arrays (every rank builds or
    loads the same ones; only the local rows are uploaded)."""

    def __init__(self, nx: int, ny: int, dx: float, dy: float, dt: float, rho: float, Var_global: np.ndarray,
                 Ff_global: np.ndarray, world: int, rank: int, device: int = 0, halo: Optional[int] = None,
                 passes_per_exchange: int = 8):
        from . import _capi as capi
        self.capi = capi
        # the halo depth is the number of sweeps per pass of the kernel for a grid of the LOCAL size
        H_guess = 4 if (nx // world) * ny >= (1 << 20) else 8
        probe = halo if halo is not None else H_guess * max(1, passes_per_exchange)
        probe = max(1, min(probe, nx // world))             # a slab cannot be thinner than its halo
        self.part = SlabPartition(nx, world, rank, probe if world > 1 else 0)
        p = capi.Params()
        p.nx, p.ny = self.part.nx_local, ny
        p.dx, p.dy, p.volp, p.dt, p.nu, p.rho = dx, dy, dx * dy, dt, 1.0, rho
        p.scheme, p.inner_tol, p.inner_max, p.sweep_order, p.device = 0, 0.0, 1000, capi.ORDER_JACOBI, device
        for k in range(3):
            for s in range(4):
                p.bc_types[k][s] = 1 if k == 2 else 0
        self.h = capi.Handle(p)
        self.H = self.h.jacobi_pass_max()
        self.nsw_max = min(self.H, self.part.halo) if world > 1 else self.H      # sweeps per kernel pass
        self.ny, self.pitch = ny, ny + 2
        self.ncells_global = nx * ny
        g0, g1 = self.part.global_rows()
        self.h.upload(Var=np.ascontiguousarray(Var_global[:, g0:g1 + 1]), Ff=np.ascontiguousarray(Ff_global[:, g0:g1 + 1]))
        self._rhs_done = False
        ptrs = self.h.device_ptrs()
        self.p_ptr = ptrs[0] + 2 * (self.part.nx_local + 2) * self.pitch * 8       # plane k = 2
        self._sums = None
        self._ext = None

    # rows as torch tensors aliasing the library's memory
    def _rows(self, first_row: int, nrows: int):
        import torch
        return torch.as_tensor(_DevRows(self.p_ptr + first_row * self.pitch * 8, nrows, self.pitch), device=f"cuda:{self.h.params.device}")

    def run_pass(self, nsw: int, commit_now, slot: Optional[int] = None):
        """Enqueue one pass; the per-sweep sums stay on the device (slot `slot` of jacobi_sums_ptr).  Two calling
        conventions: run_pass(nsw, commit_now) for slab_jacobi_solve, run_pass(nsw, slot) (always commits) for the
        block driver."""
        if slot is None and not isinstance(commit_now, bool):
            slot, commit_now = int(commit_now), True
        self.h.k_jacobi_pass_device(nsw, self.part.local_own0, self.part.local_own1, recompute_rhs=not self._rhs_done,
                                    commit=bool(commit_now), slot=slot or 0)
        self._rhs_done = True
        return np.zeros(nsw)                                 # placeholder: the values live at jacobi_sums_ptr

    def snapshot(self):
        self.h.k_jacobi_snapshot(False)

    def restore(self):
        self.h.k_jacobi_snapshot(True)

    def _sums_tensor(self):
        import torch
        if self._sums is None:
            self._sums = torch.as_tensor(_DevRows(self.h.jacobi_sums_ptr(), 16, 8), device=f"cuda:{self.h.params.device}")
        return self._sums

    def reduce(self, nslots: int) -> np.ndarray:
        """(nslots, 8) per-sweep sums of the last block, summed over the ranks (one collective, one synchronisation)."""
        import torch.distributed as dist
        t = self._sums_tensor()
        with self._stream():
            if self.part.world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.SUM)
            return t[:nslots].cpu().numpy()                  # the one synchronisation of a block

    def read_sums(self, n: int) -> np.ndarray:
        """Per-sweep sums of the last pass (this slab's owned rows), copied to the host."""
        self.h.synchronize()
        return self._sums_tensor()[0, :n].cpu().numpy()

    def commit(self):
        self.h.k_jacobi_commit()

    def _stream(self):
        """The library's own stream as a torch stream: collectives issued under it are ordered with the kernels, so the
        block loop needs no host synchronisation besides reading the reduced sums."""
        import torch
        if self._ext is None:
            self._ext = torch.cuda.ExternalStream(self.h.stream(), device=f"cuda:{self.h.params.device}")
        return torch.cuda.stream(self._ext)

    def exchange(self):
        """Owned edge rows -> neighbours' halo rows (NCCL point-to-point, both directions in one batch)."""
        import torch.distributed as dist
        P = self.part
        if P.world == 1:
            return
        ops = []
        if P.lo:        # neighbour above: send my first H owned rows, receive its last H owned rows into my upper halo
            ops += [dist.P2POp(dist.isend, self._rows(P.local_own0, P.halo), P.rank - 1),
                    dist.P2POp(dist.irecv, self._rows(1, P.halo), P.rank - 1)]
        if P.hi:
            ops += [dist.P2POp(dist.isend, self._rows(P.local_own1 - P.halo + 1, P.halo), P.rank + 1),
                    dist.P2POp(dist.irecv, self._rows(P.local_own1 + 1, P.halo), P.rank + 1)]
        with self._stream():
            for w in dist.batch_isend_irecv(ops):
                w.wait()                                     # stream-level wait: the next pass is ordered after the receive

    def allreduce_sum(self, v: np.ndarray) -> np.ndarray:
        import torch, torch.distributed as dist
        t = self._sums_tensor()[0]
        with self._stream():
            if self.part.world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.SUM)
            return t[: len(v)].cpu().numpy()

    def solve(self, tol: float = 1e-6, max_iter: int = 1000) -> Tuple[int, float]:
        part = self.part if self.part.world > 1 else SlabPartition(self.part.nx, 1, 0, self.H)
        M = max(1, min(16, part.halo // self.nsw_max)) if self.part.world > 1 else 8
        return slab_jacobi_solve_blocks(part, self.ncells_global, tol, max_iter, self, self.nsw_max, M)

    def owned_rows(self) -> np.ndarray:
        """(n_own, ny+2) owned rows of the pressure plane."""
        nxl = self.part.nx_local
        Var = np.zeros((3, nxl + 2, self.ny + 2))
        self.h.download(Var=Var)
        return Var[2, self.part.local_own0:self.part.local_own1 + 1].copy()
