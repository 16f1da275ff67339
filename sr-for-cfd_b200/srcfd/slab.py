"""Slab domain decomposition of the fine-grid iteration across GPUs (BASELINE configs[3]: a large grid split along i).

Product path: `GpuSlab` + the group functions at the bottom (solve_pressure / solve_momentum / step) bind libsrcfd's
srcfd_slab_* entry points: the whole JACOBI-order outer iteration (PyCFD_ML_accelerated.py:432-501) decomposed, halo
rows and residual sums moved by the kernels themselves through peer-mapped memory (csrc/slab.cuh), the block loop and
the break rule inside the library.  torch.distributed is used ONCE, at set-up, to hand the 64-byte cudaIpc blobs round.

The two `slab_jacobi_solve*` functions below are the executable model of that block protocol (back-end agnostic: pass /
commit / exchange / all-reduce are callables); the CPU tests run them under gloo with the oracle's arithmetic.

Model scope: the JACOBI-order pressure solve (the order BASELINE's north star names for the decomposed grid), i.e.
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
# GPU path: libsrcfd's srcfd_slab_* entry points (csrc/slab.cuh, csrc/slab_api.inl)
# ---------------------------------------------------------------------------------------------------------------------
DEFAULT_HALO = 16


class GpuSlab:
    """One slab (rank) of a decomposed flow case on one GPU.

    params: srcfd Params of the WHOLE case (nx = global rows; dx, dy, BCs, ... as for an undivided handle;
    params.device = the CUDA device of THIS rank).  The local handle owns global rows part.global_rows()."""

    def __init__(self, params, world: int, rank: int, halo: int = DEFAULT_HALO):
        import copy
        import ctypes as C
        from . import _capi as capi
        self.capi = capi
        self.nx_global, self.ny = params.nx, params.ny
        self.part = SlabPartition(params.nx, world, rank, halo if world > 1 else 0)
        local = type(params).from_buffer_copy(bytes(params))
        local.nx = self.part.nx_local
        local.sweep_order = capi.ORDER_JACOBI
        self.h = capi.Handle(local)
        capi.check(capi.lib().srcfd_slab_configure(self.h._h, C.c_int(world), C.c_int(rank), C.c_int(params.nx), C.c_int(halo)))

    # ---- state
    def _rows(self, a):
        g0, g1 = self.part.global_rows()
        return np.ascontiguousarray(a[:, g0:g1 + 1])

    def upload_global(self, Var=None, VarOld=None, Ff=None):
        """Upload this slab's rows of full-domain (3|4, nx+2, ny+2) arrays."""
        self.h.upload(Var=None if Var is None else self._rows(Var), VarOld=None if VarOld is None else self._rows(VarOld),
                      Ff=None if Ff is None else self._rows(Ff))

    def download_local(self):
        nxl = self.part.nx_local
        Var = np.zeros((3, nxl + 2, self.ny + 2)); VarOld = np.zeros_like(Var); Ff = np.zeros((4, nxl + 2, self.ny + 2))
        self.h.download(Var=Var, VarOld=VarOld, Ff=Ff)
        return Var, VarOld, Ff

    def owned(self):
        """(Var, VarOld, Ff) restricted to the owned rows: shapes (3|3|4, n_own, ny+2)."""
        a, b = self.part.local_own0, self.part.local_own1 + 1
        return tuple(x[:, a:b].copy() for x in self.download_local())

    def owned_rows(self) -> np.ndarray:
        """(n_own, ny+2) owned rows of the pressure plane."""
        Var = np.zeros((3, self.part.nx_local + 2, self.ny + 2))
        self.h.download(Var=Var)
        return Var[2, self.part.local_own0:self.part.local_own1 + 1].copy()

    def info(self) -> dict:
        import ctypes as C
        r0, r1 = C.c_int32(0), C.c_int32(0)
        ex, hb, rp = C.c_int64(0), C.c_int64(0), C.c_int64(0)
        self.capi.check(self.capi.lib().srcfd_slab_info(self.h._h, C.byref(r0), C.byref(r1), C.byref(ex), C.byref(hb), C.byref(rp)))
        return dict(own_row0=r0.value, own_row1=r1.value, exchanges=ex.value, halo_bytes=hb.value, replays=rp.value)

    def kernel_stats(self) -> dict:
        import ctypes as C
        a, b, c, d = C.c_int64(0), C.c_int64(0), C.c_int64(0), C.c_int64(0)
        self.capi.check(self.capi.lib().srcfd_slab_kernel_stats(self.h._h, C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
        return dict(stream_solves=a.value, tile_solves=b.value, last_retries=c.value, last_warp_steps=d.value)

    def export_blob(self) -> bytes:
        import ctypes as C
        buf = (C.c_ubyte * 64)()
        self.capi.check(self.capi.lib().srcfd_slab_export(self.h._h, buf, C.c_int(64)))
        return bytes(buf)

    def attach_blob(self, peer_rank: int, blob: bytes):
        import ctypes as C
        buf = (C.c_ubyte * 64).from_buffer_copy(blob)
        self.capi.check(self.capi.lib().srcfd_slab_attach_ipc(self.h._h, C.c_int(peer_rank), buf))

    def close(self):
        self.h.close()


def attach_local(slabs):
    """Several slabs driven by ONE process: map each one's mailbox into the others (same device: plain pointers)."""
    import ctypes as C
    from . import _capi as capi
    for a in slabs:
        for b in slabs:
            if a is not b:
                capi.check(capi.lib().srcfd_slab_attach_local(a.h._h, C.c_int(b.part.rank), b.h._h))


def attach_distributed(slab: GpuSlab, group=None):
    """One process per GPU: hand the cudaIpc blobs round with torch.distributed (set-up only; any backend) and map the
    peers' mailboxes.  After this no torch / NCCL call is involved in the data path."""
    import torch.distributed as dist
    world = slab.part.world
    if world == 1:
        return
    blobs = [None] * world
    dist.all_gather_object(blobs, slab.export_blob(), group=group)
    for q in range(world):
        if q != slab.part.rank:
            slab.attach_blob(q, blobs[q])
    dist.barrier(group=group)


def _harr(slabs):
    import ctypes as C
    return (C.c_void_p * len(slabs))(*[s.h._h for s in slabs]), C.c_int(len(slabs))


def exchange(slabs, k: int):
    import ctypes as C
    from . import _capi as capi
    arr, n = _harr(slabs)
    capi.check(capi.lib().srcfd_slab_exchange(arr, n, C.c_int(k)))


def solve_pressure(slabs) -> Tuple[int, float]:
    """solve_pressure (PyCFD_ML_accelerated.py:292-314), JACOBI order, over the slabs this process drives."""
    import ctypes as C
    from . import _capi as capi
    arr, n = _harr(slabs)
    sw, rms = C.c_int32(0), C.c_double(0.0)
    capi.check(capi.lib().srcfd_slab_solve_pressure(arr, n, C.byref(sw), C.byref(rms)))
    return sw.value, rms.value


def solve_momentum(slabs, k: int, scheme: int) -> Tuple[int, float]:
    import ctypes as C
    from . import _capi as capi
    arr, n = _harr(slabs)
    sw, rms = C.c_int32(0), C.c_double(0.0)
    capi.check(capi.lib().srcfd_slab_solve_momentum(arr, n, C.c_int(k), C.c_int(scheme), C.byref(sw), C.byref(rms)))
    return sw.value, rms.value


def step(slabs, n_outer: int, crit=(1e-6, 1e-6, 1e-6)):
    """n_outer x (_implicit_solve + _convergence_check), PyCFD_ML_accelerated.py:408-419."""
    import ctypes as C
    from . import _capi as capi
    arr, n = _harr(slabs)
    c = (C.c_double * 3)(*[float(x) for x in crit])
    capi.check(capi.lib().srcfd_slab_step(arr, n, C.c_int64(int(n_outer)), c))
