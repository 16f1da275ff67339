"""Slab decomposition on the GPU back-end: the single-pass C-ABI entry (srcfd_k_jacobi_pass / _commit) driven by
srcfd.slab on two slabs of one device (threads stand in for ranks, host copies for NVLink), against the single-domain
oracle in Jacobi order.  The NCCL path itself is exercised by tools/slab_bench.py on a multi-GPU box."""
import threading

import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu


class _Pair:
    """Two GpuSlab objects in one process: exchange and all-reduce through host memory, in lock step."""

    def __init__(self, slabs):
        self.slabs = slabs
        self.bar = threading.Barrier(len(slabs))
        self.box = [None] * len(slabs)

    def exchange(self, r):
        s = self.slabs[r]; P = s.part
        nxl = P.nx_local
        V = np.zeros((3, nxl + 2, s.ny + 2)); s.h.download(Var=V)
        self.box[r] = V[2].copy()
        self.bar.wait()
        if P.lo:
            up = self.slabs[r - 1]; U = self.box[r - 1]
            V[2, 1:1 + P.halo] = U[up.part.local_own1 - P.halo + 1:up.part.local_own1 + 1]
        if P.hi:
            dn = self.slabs[r + 1]; D = self.box[r + 1]
            V[2, P.local_own1 + 1:P.local_own1 + 1 + P.halo] = D[dn.part.local_own0:dn.part.local_own0 + P.halo]
        self.bar.wait()
        s.h.upload(Var=V)

    def allreduce(self, r, v):
        self.box[r] = self.slabs[r].read_sums(len(v)).astype(np.float64)      # the pass left its sums on the device
        self.bar.wait()
        tot = sum(self.box[i] for i in range(len(self.slabs)))
        self.bar.wait()
        return tot


@pytest.mark.parametrize("world", [1, 2, 3])
def test_slab_jacobi_on_gpu_matches_oracle(world):
    from srcfd.slab import GpuSlab, slab_jacobi_solve
    nx, ny = 96, 70
    rng = np.random.default_rng(1)
    Var = rng.uniform(-1, 1, (3, nx + 2, ny + 2)); Ff = 0.05 * rng.uniform(-1, 1, (4, nx + 2, ny + 2))
    dx, dy, dt, rho = 1.0 / nx, 1.0 / ny, 1e-3, 1.0
    for tol, cap in ((0.0, 19), (30.0, 200), (1e-30, 8)):
        slabs = [GpuSlab(nx, ny, dx, dy, dt, rho, Var, Ff, world, r, device=0) for r in range(world)]
        pair = _Pair(slabs)
        res = [None] * world

        def run(r):
            s = slabs[r]
            res[r] = slab_jacobi_solve(s.part if world > 1 else type(s.part)(nx, 1, 0, s.H), nx * ny, tol, cap, s.run_pass, s.commit,
                                       lambda: pair.exchange(r), lambda v: pair.allreduce(r, v), sweeps_per_pass=s.nsw_max)

        th = [threading.Thread(target=run, args=(r,)) for r in range(world)]
        for t in th: t.start()
        for t in th: t.join(timeout=120)
        assert all(x is not None for x in res)
        rows = np.concatenate([s.owned_rows() for s in slabs], axis=0)
        if world == 1:                                      # the block driver GpuSlab.solve() uses (no communication needed here)
            s1 = GpuSlab(nx, ny, dx, dy, dt, rho, Var, Ff, 1, 0, device=0)
            n1, _ = s1.solve(tol, cap)
            assert n1 == res[0][0] and np.array_equal(s1.owned_rows(), rows)
            s1.h.close()
        B = Var.copy()
        m = O.solve_pressure(B, Ff, nx, ny, dx, dy, dt, rho, dx * dy, order=O.ORDER_JACOBI, tolerance=tol, max_iter=cap)
        assert all(n == m for n, _ in res), (world, tol, cap, res, m)
        assert np.array_equal(rows, B[2, 1:-1]), (world, tol, cap, np.max(np.abs(rows - B[2, 1:-1])))
        for s in slabs: s.h.close()


def _nccl_worker(rank, world, port, q):
    import os
    import torch, torch.distributed as dist
    from srcfd.slab import GpuSlab
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), NCCL_DEBUG="WARN")
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    nx, ny = 256, 200
    rng = np.random.default_rng(4)
    Var = rng.uniform(-1, 1, (3, nx + 2, ny + 2)); Ff = 0.05 * rng.uniform(-1, 1, (4, nx + 2, ny + 2))
    out = []
    for tol, cap in ((0.0, 70), (60.0, 300)):
        s = GpuSlab(nx, ny, 1.0 / nx, 1.0 / ny, 1e-3, 1.0, Var, Ff, world, rank, device=rank, passes_per_exchange=2)
        n, rms = s.solve(tol, cap)
        rows = [None] * world
        dist.all_gather_object(rows, s.owned_rows())
        out.append((n, np.concatenate(rows, axis=0)))
        s.h.close()
    if rank == 0:
        q.put(out)
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpu_nccl_slab_matches_oracle():
    """The real thing: two processes, two GPUs, NCCL halo exchange and all-reduce (skipped on a single-GPU box)."""
    import socket
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs: p.start()
    got = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120); assert p.exitcode == 0
    nx, ny = 256, 200
    rng = np.random.default_rng(4)
    Var = rng.uniform(-1, 1, (3, nx + 2, ny + 2)); Ff = 0.05 * rng.uniform(-1, 1, (4, nx + 2, ny + 2))
    for (tol, cap), (n, rows) in zip(((0.0, 70), (60.0, 300)), got):
        B = Var.copy()
        m = O.solve_pressure(B, Ff, nx, ny, 1.0 / nx, 1.0 / ny, 1e-3, 1.0, 1.0 / (nx * ny), order=O.ORDER_JACOBI, tolerance=tol, max_iter=cap)
        assert n == m, (tol, cap, n, m)
        assert np.array_equal(rows, B[2, 1:-1]), (tol, cap, np.max(np.abs(rows - B[2, 1:-1])))
