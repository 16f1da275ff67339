"""Ensemble sharding and gathering (the N>1 host logic) on CPU with a world_size-2 gloo group."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from srcfd import ensemble as E


def test_shard_cases_partition():
    cases = E.multibc_sweep()
    assert len(cases) == 14 * 2 + 3
    for world in (1, 2, 4, 8):
        parts = [E.shard_cases(cases, r, world) for r in range(world)]
        assert sum(len(p) for p in parts) == len(cases)
        assert sorted(c.label() for p in parts for c in p) == sorted(c.label() for c in cases)
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    with pytest.raises(ValueError):
        E.shard_cases(cases, 2, 2)


def _fake_runner(spec, device=0, max_ctas=0, **kw):
    f = np.full((3, 2, 2), spec.Re, dtype=np.float64)
    return E.CaseResult(spec.label(), -1, int(spec.Re), True, 0.0, [1, 2, 3], [0.0, 0.0, 0.0], f)


def test_run_local_concurrency_keeps_order():
    cases = E.multibc_sweep(res_ldc=(50, 100, 150), res_bfs=(100,))
    res = E.run_local(cases, concurrency=3, runner=_fake_runner)
    assert [r.label for r in res] == [c.label() for c in cases]


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cases = E.multibc_sweep(res_ldc=(50, 100, 150, 200), res_bfs=(100, 400))
    out = E.run_ensemble(cases, concurrency=2, runner=_fake_runner, dist=dist, device=0)
    if rank == 0:
        q.put([(r.label, r.rank, r.iterations, float(r.fields[0, 0, 0])) for r in out])
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_gather():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    cases = E.multibc_sweep(res_ldc=(50, 100, 150, 200), res_bfs=(100, 400))
    assert [g[0] for g in got] == [c.label() for c in cases]                 # gathered in sweep order
    assert [g[1] for g in got] == [i % 2 for i in range(len(cases))]         # round-robin ownership
    assert all(g[2] == int(c.Re) and g[3] == c.Re for g, c in zip(got, cases))


def test_run_local_hands_each_case_its_warm_field():
    """run_local(warm_fields=...) passes case i its own initial guess (the warm stage ran up front) and nothing for
    cases without one."""
    seen = {}

    def runner(spec, device=0, max_ctas=0, warm=None, **kw):
        seen[spec.label()] = None if warm is None else float(warm[0, 0, 0])
        return _fake_runner(spec, device, max_ctas)

    cases = E.multibc_sweep(res_ldc=(50, 100), res_bfs=(400,))
    warm = [np.full((3, 2, 2), float(i)) if i % 2 == 0 else None for i in range(len(cases))]
    res = E.run_local(cases, concurrency=4, runner=runner, warm_fields=warm)
    assert [r.label for r in res] == [c.label() for c in cases]
    assert seen == {c.label(): (float(i) if i % 2 == 0 else None) for i, c in enumerate(cases)}
