"""Slab decomposition of the Jacobi pressure relaxation: the partition/halo/break logic on CPU with a world_size-2
gloo group.  The per-pass arithmetic is the oracle's (one Jacobi sweep = oracle.solve_pressure(order=JACOBI, max_iter=1)),
so the two-rank result must equal the single-domain oracle bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from oracle import oracle as O
from srcfd.slab import SlabPartition, slab_jacobi_solve, slab_jacobi_solve_blocks


def test_partition_covers_rows_once():
    for nx, world, halo in ((400, 1, 8), (400, 2, 8), (4096, 8, 4), (37, 3, 4), (64, 8, 8)):
        parts = [SlabPartition(nx, world, r, halo if world > 1 else 0) for r in range(world)]
        rows = [g for p in parts for g in range(p.own0, p.own1 + 1)]
        assert rows == list(range(1, nx + 1))
        for p in parts:
            assert p.lo == (halo if p.rank > 0 and world > 1 else 0) and p.hi == (halo if p.rank < world - 1 and world > 1 else 0)
            g0, g1 = p.global_rows()
            assert g0 == p.own0 - p.lo - 1 and g1 - g0 + 1 == p.nx_local + 2 and g0 >= 0 and g1 <= nx + 1
    with pytest.raises(ValueError):
        SlabPartition(20, 8, 0, 8)          # slabs thinner than the halo


class _NumpySlab:
    """Back-end with the oracle's arithmetic: local Var (3, nxl+2, ny+2) and Ff (4, ...)."""

    def __init__(self, part, Var, Ff, ny, dx, dy, dt, rho, dist):
        self.part, self.ny, self.dx, self.dy, self.dt, self.rho, self.dist = part, ny, dx, dy, dt, rho, dist
        g0, g1 = part.global_rows()
        self.Var = np.ascontiguousarray(Var[:, g0:g1 + 1]); self.Ff = np.ascontiguousarray(Ff[:, g0:g1 + 1])
        self.pending = None
        self.slots = np.zeros((16, 8))

    def _sweeps(self, nsw):
        V = self.Var.copy()
        sums = np.zeros(nsw)
        P = self.part
        rhs = self.rho / self.dt * (self.Ff[0] + self.Ff[1] + self.Ff[2] + self.Ff[3])
        for t in range(nsw):
            p = V[2]
            Fd = (self.dx * self.dy) * ((p[2:, 1:-1] - 2.0 * p[1:-1, 1:-1] + p[:-2, 1:-1]) / (self.dx * self.dx)
                                        + (p[1:-1, 2:] - 2.0 * p[1:-1, 1:-1] + p[1:-1, :-2]) / (self.dy * self.dy))
            R = rhs[1:-1, 1:-1] - Fd
            sums[t] = float(np.sum(R[P.local_own0 - 1:P.local_own1] ** 2))
            O.solve_pressure(V, self.Ff, P.nx_local, self.ny, self.dx, self.dy, self.dt, self.rho, self.dx * self.dy,
                             order=O.ORDER_JACOBI, tolerance=0.0, max_iter=1)
        return V, sums

    def run_pass(self, nsw, commit_now):
        V, sums = self._sweeps(nsw)
        if commit_now: self.Var = V
        else: self.pending = V
        return sums

    def commit(self):
        self.Var = self.pending

    # back-end interface of the block driver
    def snapshot(self): self.saved = self.Var.copy()
    def restore(self): self.Var = self.saved.copy()
    def run_pass_slot(self, nsw, slot):
        self.Var, sums = self._sweeps(nsw)
        self.slots[slot, :nsw] = sums
    def reduce(self, nslots):
        import torch
        t = torch.from_numpy(self.slots.copy()); self.dist.all_reduce(t); return t.numpy()[:nslots]

    def exchange(self):
        import torch
        P, d = self.part, self.dist
        if P.world == 1: return
        p = self.Var[2]
        reqs, recv = [], []
        if P.lo:
            reqs.append(d.isend(torch.from_numpy(np.ascontiguousarray(p[P.local_own0:P.local_own0 + P.halo])), P.rank - 1))
            t = torch.empty((P.halo, self.ny + 2), dtype=torch.float64); recv.append((t, 1)); reqs.append(d.irecv(t, P.rank - 1))
        if P.hi:
            reqs.append(d.isend(torch.from_numpy(np.ascontiguousarray(p[P.local_own1 - P.halo + 1:P.local_own1 + 1])), P.rank + 1))
            t = torch.empty((P.halo, self.ny + 2), dtype=torch.float64); recv.append((t, P.local_own1 + 1)); reqs.append(d.irecv(t, P.rank + 1))
        for r in reqs: r.wait()
        for t, row in recv: p[row:row + P.halo] = t.numpy()

    def allreduce(self, v):
        import torch
        t = torch.from_numpy(np.array(v, dtype=np.float64)); self.dist.all_reduce(t); return t.numpy()


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    nx, ny, H = 40, 28, 4
    rng = np.random.default_rng(0)
    Var = rng.uniform(-1, 1, (3, nx + 2, ny + 2)); Ff = 0.05 * rng.uniform(-1, 1, (4, nx + 2, ny + 2))
    dx, dy, dt, rho = 1.0 / nx, 1.0 / ny, 1e-3, 1.0
    out = []
    for tol, cap in ((0.0, 10), (0.0, 3), (30.0, 40), (24.0, 60), (1e-30, 9)):      # caps off the pass size, stops inside a pass
        halo = H if cap % 2 else 2 * H                   # also: halos twice as deep as a pass, exchanged every second pass
        part = SlabPartition(nx, world, rank, halo)
        be = _NumpySlab(part, Var, Ff, ny, dx, dy, dt, rho, dist)
        n, rms = slab_jacobi_solve(part, nx * ny, tol, cap, be.run_pass, be.commit, be.exchange, be.allreduce, sweeps_per_pass=H)
        # the same problem through the speculative block driver (M = 2 passes per exchange / reduction)
        part2 = SlabPartition(nx, world, rank, 2 * H)
        be2 = _NumpySlab(part2, Var, Ff, ny, dx, dy, dt, rho, dist)
        be2.run_pass = be2.run_pass_slot
        n2, rms2 = slab_jacobi_solve_blocks(part2, nx * ny, tol, cap, be2, H, 2)
        assert n2 == n and np.array_equal(be2.Var[2, part2.local_own0:part2.local_own1 + 1], be.Var[2, part.local_own0:part.local_own1 + 1])
        own = be.Var[2, part.local_own0:part.local_own1 + 1]
        gathered = [None] * world
        dist.all_gather_object(gathered, own)
        out.append((n, rms, np.concatenate(gathered, axis=0)))
    if rank == 0:
        q.put(out)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_slab_jacobi_matches_single_domain_oracle():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs: p.start()
    got = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60); assert p.exitcode == 0
    nx, ny = 40, 28
    rng = np.random.default_rng(0)
    Var = rng.uniform(-1, 1, (3, nx + 2, ny + 2)); Ff = 0.05 * rng.uniform(-1, 1, (4, nx + 2, ny + 2))
    counts = []
    for (tol, cap), (n, rms, rows) in zip(((0.0, 10), (0.0, 3), (30.0, 40), (24.0, 60), (1e-30, 9)), got):
        B = Var.copy()
        m = O.solve_pressure(B, Ff, nx, ny, 1.0 / nx, 1.0 / ny, 1e-3, 1.0, 1.0 / (nx * ny), order=O.ORDER_JACOBI, tolerance=tol, max_iter=cap)
        assert n == m, (tol, cap, n, m)
        assert np.array_equal(rows, B[2, 1:-1]), (tol, cap, np.max(np.abs(rows - B[2, 1:-1])))
        counts.append(n)
    assert counts[0] == 10 and counts[1] == 3 and 1 < counts[2] < 40 and counts[2] % 4 != 0 or counts[3] % 4 != 0


# ---- the set-up plumbing of the product path: cudaIpc blobs handed round once over torch.distributed (gloo here) -------------
class _FakeSlab:
    """Stands in for GpuSlab where there is no GPU: same attributes attach_distributed touches."""

    def __init__(self, world, rank):
        self.part = SlabPartition(64, world, rank, 8)
        self.attached = {}

    def export_blob(self):
        return bytes([self.part.rank + 1]) * 64

    def attach_blob(self, peer_rank, blob):
        self.attached[peer_rank] = blob


def _attach_worker(rank, world, port, q):
    import torch.distributed as dist
    from srcfd import slab
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    s = _FakeSlab(world, rank)
    slab.attach_distributed(s)
    q.put((rank, {k: v[:2] for k, v in s.attached.items()}))
    dist.barrier()
    dist.destroy_process_group()


def test_attach_distributed_hands_every_peer_blob_to_every_rank():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_attach_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs: p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60); assert p.exitcode == 0
    assert got[0] == {1: bytes([2, 2])} and got[1] == {0: bytes([1, 1])}      # everybody else's blob, never its own


def test_slab_entry_points_reject_bad_arguments_without_a_gpu():
    """The srcfd_slab_* entry points are exported and fail with an argument error (no crash, no CPU fallback) on null input."""
    import ctypes as C
    from srcfd import _capi as capi
    L = capi.lib()
    arr = (C.c_void_p * 1)(None)
    crit = (C.c_double * 3)(1e-6, 1e-6, 1e-6)
    assert L.srcfd_slab_step(arr, C.c_int(1), C.c_int64(1), crit) == capi.ERR_ARG
    assert L.srcfd_slab_step(None, C.c_int(0), C.c_int64(1), crit) == capi.ERR_ARG
    assert L.srcfd_slab_solve_pressure(arr, C.c_int(1), None, None) == capi.ERR_ARG
    assert L.srcfd_slab_exchange(arr, C.c_int(1), C.c_int(2)) == capi.ERR_ARG
    assert b"slab" in L.srcfd_last_error()
    assert L.srcfd_slab_configure(None, C.c_int(2), C.c_int(0), C.c_int(64), C.c_int(8)) == capi.ERR_ARG
    p = SlabPartition(100, 4, 3, 8)
    assert (p.own0, p.own1, p.lo, p.hi, p.nx_local) == (76, 100, 8, 0, 33) and p.global_rows() == (67, 101)
