"""Slab decomposition on the GPU (csrc/slab.cuh + slab_api.inl through the C ABI) against the single-domain oracle in
JACOBI order: inner solves and whole outer iterations, 1/2/3 slabs.

On one GPU the slabs of a case are driven by ONE process and share one stream (srcfd_slab_attach_local): pushes, gates
and sweeps execute in enqueue order, so the peer-mailbox protocol (stores + sequence flags, double buffering, rank-order
sums, speculative blocks with replay) is exercised without kernels that wait for one another.  The two-process cudaIpc
path needs two GPUs (test at the bottom; it also runs inside `bench.py --gpus N`, which reports `slab_parity`)."""
import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _params(case: O.Case, device=0):
    from srcfd import _capi as capi
    p = capi.Params()
    p.nx, p.ny = case.nx, case.ny
    p.dx, p.dy = case.lx / case.nx, case.ly / case.ny
    p.volp = p.dx * p.dy
    p.dt, p.nu, p.rho = case.dt, 1.0 / case.Re, case.rho
    p.scheme = capi.SCHEME_QUICK if case.scheme == "QUICK" else capi.SCHEME_UPWIND
    for k in range(3):
        for s in range(4):
            p.bc_types[k][s] = int(case.bc_types[k][s]); p.bc_values[k][s] = float(case.bc_values[k][s])
    p.bfs_enabled = int(case.bfs)
    p.bfs_step_h, p.bfs_h, p.bfs_Ub = case.step_h, case.h, case.Ub
    p.relax_enabled = int(case.relax is not None)
    a = case.relax or (1.0, 1.0, 1.0)
    p.relax[0], p.relax[1], p.relax[2] = a
    p.inner_tol, p.inner_max = case.inner_tol, case.inner_max
    p.sweep_order = capi.ORDER_JACOBI
    p.device = device
    return p


def _make(case, world, halo, Var=None, VarOld=None, Ff=None):
    from srcfd import slab
    slabs = [slab.GpuSlab(_params(case), world, r, halo=halo) for r in range(world)]
    slab.attach_local(slabs)
    for s in slabs:
        s.upload_global(Var=Var, VarOld=VarOld, Ff=Ff)
    return slabs


def _gather(slabs):
    parts = [s.owned() for s in slabs]
    return tuple(np.concatenate([p[i] for p in parts], axis=1) for i in range(3))


def _close(slabs):
    for s in slabs:
        s.close()


@pytest.fixture(params=["tiles", "stream", "stream1"])
def pressure_kernel(request, monkeypatch):
    """The pressure kernels of the slab path: the shared-memory tile kernel (what thin planes like these tests' get by
    default) and the warp-streaming kernel with two columns per lane and with one (forced: no probe solve, no fallback)."""
    monkeypatch.delenv("SRCFD_JTB2_FORCE", raising=False)
    monkeypatch.delenv("SRCFD_JTB2_COLS", raising=False)
    if request.param != "tiles":
        monkeypatch.setenv("SRCFD_JTB2_FORCE", "1")
        monkeypatch.setenv("SRCFD_JTB2_COLS", "1" if request.param == "stream1" else "2")
    return request.param


@pytest.mark.parametrize("world", [1, 2, 3])
def test_slab_pressure_matches_oracle(world, pressure_kernel):
    from srcfd import slab
    nx, ny = 120, 70
    rng = np.random.default_rng(1)
    Var = rng.uniform(-1, 1, (3, nx + 2, ny + 2)); Ff = 0.05 * rng.uniform(-1, 1, (4, nx + 2, ny + 2))
    dx, dy, dt, rho = 1.0 / nx, 1.0 / ny, 1e-3, 1.0
    # never met / met in the middle of a block (replay) / cap in the middle of a pass / met at once / long run
    for tol, cap in ((0.0, 19), (30.0, 200), (1e-30, 8), (1e9, 50), (20.0, 300)):
        case = O.Case(nx=nx, ny=ny, dt=dt, inner_tol=tol, inner_max=cap, order=O.ORDER_JACOBI)
        slabs = _make(case, world, 8, Var=Var, Ff=Ff)
        n, rms = slab.solve_pressure(slabs)
        B = Var.copy()
        m, hist = O.solve_pressure(B, Ff, nx, ny, dx, dy, dt, rho, dx * dy, order=O.ORDER_JACOBI, tolerance=tol, max_iter=cap, rms_hist=True)
        got = _gather(slabs)[0]
        assert n == m, (world, tol, cap, n, m)
        assert np.array_equal(got[2], B[2, 1:-1]), (world, tol, cap, np.max(np.abs(got[2] - B[2, 1:-1])))
        assert abs(rms - hist[-1]) <= 1e-12 * abs(hist[-1]), (rms, hist[-1])
        if world > 1 and cap > 8:
            assert slabs[0].info()["exchanges"] >= 2 and slabs[0].info()["halo_bytes"] > 0
        ks = slabs[0].kernel_stats()
        assert (ks["stream_solves"], ks["tile_solves"]) == ((0, 1) if pressure_kernel == "tiles" else (1, 0))
        _close(slabs)


@pytest.mark.parametrize("world", [1, 2])
def test_slab_pressure_rest_and_denormal_fields(world, pressure_kernel):
    """A field at rest with a small source (zeros beyond a front that decays into the denormals -- the start of every
    cavity run) and a field of tiny values: the streaming kernel's rows-at-rest shortcut, its redo path and the scaled
    division (div_mid) against the oracle, signed zeros included.  The front lies across the lanes of a strip in the
    first grid and along the rows in the second."""
    from srcfd import slab
    rng = np.random.default_rng(5)
    for nx, ny, lx, ly, cap in ((150, 230, 1.0, 6 * 230 / 150, 220), (230, 150, 6 * 230 / 150, 1.0, 220), (96, 140, 1.0, 1.0, 37)):
        Var = np.zeros((3, nx + 2, ny + 2)); Ff = np.zeros((4, nx + 2, ny + 2))
        if cap == 37:
            Var[2] = rng.uniform(-1, 1, (nx + 2, ny + 2)) * 2.0 ** rng.integers(-1074, -940, (nx + 2, ny + 2))
            Ff[:] = rng.uniform(-1, 1, Ff.shape) * 2.0 ** rng.integers(-1074, -960, Ff.shape)
            Var[2, 30:60] = 0.0; Ff[:, 29:61] = 0.0            # rows at rest inside a tiny-valued field
        else:
            Ff[:, 3:9, 4:10] = 1e-3 * rng.uniform(-1, 1, (4, 6, 6))
        dx, dy = lx / nx, ly / ny
        case = O.Case(nx=nx, ny=ny, lx=lx, ly=ly, dt=1e-3, inner_tol=0.0, inner_max=cap, order=O.ORDER_JACOBI)
        slabs = _make(case, world, 16, Var=Var, Ff=Ff)
        n, rms = slab.solve_pressure(slabs)
        B = Var.copy()
        m, hist = O.solve_pressure(B, Ff, nx, ny, dx, dy, 1e-3, 1.0, dx * dy, order=O.ORDER_JACOBI, tolerance=0.0, max_iter=cap, rms_hist=True)
        got = _gather(slabs)[0]
        assert n == m, (world, nx, ny, n, m)
        assert np.array_equal(got[2], B[2, 1:-1]), (world, nx, ny, np.max(np.abs(got[2] - B[2, 1:-1])))
        assert np.array_equal(np.signbit(got[2]), np.signbit(B[2, 1:-1]))
        assert abs(rms - hist[-1]) <= 1e-12 * abs(hist[-1]) + 1e-300, (rms, hist[-1])
        if cap > 37:
            a = np.abs(B[2, 1:-1, 1:-1])
            assert np.count_nonzero((a > 0) & (a < 2.0 ** -1022)) > 50 and np.count_nonzero(a == 0) > 1000
        _close(slabs)


@pytest.fixture(params=["two_per_pass", "one_per_launch"])
def momentum_kernel(request, monkeypatch):
    """The momentum sweeps of the slab path: two sweeps per pass over HBM (k_slab_sweep2, the default; an odd sweep left
    over runs alone) and one sweep per launch (k_slab_sweep)."""
    monkeypatch.delenv("SRCFD_SWEEP2_CHUNKS", raising=False)
    monkeypatch.setenv("SRCFD_SLAB_SWEEP2", "1" if request.param == "two_per_pass" else "0")
    return request.param


@pytest.mark.parametrize("world", [1, 2, 3])
@pytest.mark.parametrize("scheme", ["UPWIND", "QUICK"])
def test_slab_momentum_matches_oracle(world, scheme, momentum_kernel):
    from srcfd import slab, _capi as capi
    nx, ny = 96, 50
    rng = np.random.default_rng(2)
    Var = rng.uniform(-1, 1, (3, nx + 2, ny + 2)); VarOld = Var + 0.01 * rng.uniform(-1, 1, Var.shape)
    Ff = 0.002 * rng.uniform(-1, 1, (4, nx + 2, ny + 2))
    # wall boundaries: the boundary-face fluxes along i are exactly zero, as linear_interpolation produces them, so QUICK's
    # out-of-plane second neighbours at i = 1 / i = nx are loaded but never used (hazard H4; the one thing not decomposed)
    Ff[2, 1, :] = 0.0; Ff[0, nx, :] = 0.0
    fn = O.solve_momentum_quick if scheme == "QUICK" else O.solve_momentum_upwind
    for tol, cap in ((0.0, 7), (1e-30, 23), (1e9, 40)):
        for k in (0, 1):
            case = O.Case(nx=nx, ny=ny, Re=100.0, dt=1e-3, scheme=scheme, inner_tol=tol, inner_max=cap, order=O.ORDER_JACOBI)
            slabs = _make(case, world, 12, Var=Var, VarOld=VarOld, Ff=Ff)
            n, rms = slab.solve_momentum(slabs, k, capi.SCHEME_QUICK if scheme == "QUICK" else capi.SCHEME_UPWIND)
            B = Var.copy()
            m = fn(B, VarOld, Ff, k, nx, ny, 1.0 / nx, 1.0 / ny, 1e-3, 1.0 / 100.0, (1.0 / nx) * (1.0 / ny), order=O.ORDER_JACOBI, tolerance=tol, max_iter=cap)
            got = _gather(slabs)[0]
            assert n == m, (world, scheme, tol, cap, k, n, m)
            assert np.array_equal(got[k], B[k, 1:-1]), (world, scheme, tol, cap, k, np.max(np.abs(got[k] - B[k, 1:-1])))
            _close(slabs)


@pytest.mark.parametrize("scheme", ["UPWIND", "QUICK"])
def test_slab_momentum_two_per_pass_many_strips_and_chunks(scheme, monkeypatch):
    """k_slab_sweep2 on a plane of several column strips and row chunks (ragged last strip and last chunk, chunk counts
    forced through SRCFD_SWEEP2_CHUNKS as well), even and odd sweep counts, a tolerance met inside a block (replay), on
    uploaded fluxes with NON-zero boundary-face fluxes, so that QUICK's out-of-plane second neighbours (hazard H4) are
    used at both levels of a pass: fields and sweep counts equal to the oracle's, sums equal to the one-sweep kernel's."""
    from srcfd import slab, _capi as capi
    nx, ny = 203, 131
    rng = np.random.default_rng(12)
    Var = rng.uniform(-1, 1, (3, nx + 2, ny + 2)); VarOld = Var + 0.01 * rng.uniform(-1, 1, Var.shape)
    Ff = 0.002 * rng.uniform(-1, 1, (4, nx + 2, ny + 2))
    fn = O.solve_momentum_quick if scheme == "QUICK" else O.solve_momentum_upwind
    sc = capi.SCHEME_QUICK if scheme == "QUICK" else capi.SCHEME_UPWIND
    for chunks in ("", "1", "5"):
        monkeypatch.setenv("SRCFD_SLAB_SWEEP2", "1")
        if chunks:
            monkeypatch.setenv("SRCFD_SWEEP2_CHUNKS", chunks)
        else:
            monkeypatch.delenv("SRCFD_SWEEP2_CHUNKS", raising=False)
        for tol, cap, k in ((0.0, 6, 0), (0.0, 7, 1), (1e-30, 1, 0), (None, 30, 1)):
            if tol is None:                                   # a tolerance the 4th sweep meets: overshoot and replay
                B = Var.copy()
                _, hist = fn(B, VarOld, Ff, k, nx, ny, 1.0 / nx, 1.0 / ny, 1e-3, 1.0 / 100.0, (1.0 / nx) * (1.0 / ny), order=O.ORDER_JACOBI,
                             tolerance=0.0, max_iter=6, rms_hist=True)
                tol = 0.5 * (hist[2] + hist[3])
                assert hist[3] < tol < hist[2]
            case = O.Case(nx=nx, ny=ny, Re=100.0, dt=1e-3, scheme=scheme, inner_tol=tol, inner_max=cap, order=O.ORDER_JACOBI)
            slabs = _make(case, 1, 12, Var=Var, VarOld=VarOld, Ff=Ff)
            n, rms = slab.solve_momentum(slabs, k, sc)
            B = Var.copy()
            m, hist = fn(B, VarOld, Ff, k, nx, ny, 1.0 / nx, 1.0 / ny, 1e-3, 1.0 / 100.0, (1.0 / nx) * (1.0 / ny), order=O.ORDER_JACOBI,
                         tolerance=tol, max_iter=cap, rms_hist=True)
            got = _gather(slabs)[0]
            assert n == m, (scheme, chunks, tol, cap, k, n, m)
            assert np.array_equal(got[k], B[k, 1:-1]), (scheme, chunks, tol, cap, k, np.max(np.abs(got[k] - B[k, 1:-1])))
            assert abs(rms - hist[-1]) <= 1e-12 * abs(hist[-1]), (rms, hist[-1])
            _close(slabs)


@pytest.mark.parametrize("world", [1, 2, 3])
@pytest.mark.parametrize("scheme", ["UPWIND", "QUICK"])
def test_slab_momentum_with_library_fluxes(world, scheme, monkeypatch, momentum_kernel):
    """Momentum solves on face fluxes as the library's own kernels leave them (BCs, linear_interpolation, update_flux on
    every slab's local rows, halo rows included) instead of uploaded ones: the slabs' fluxes equal the undivided ones,
    the W/S planes are the negated E/N planes of the neighbouring cell -- which lets the sweep take the
    west flux from the row above instead of loading it (k_slab_sweep<OP, true>) -- and the solves match the oracle with
    and without that shortcut."""
    from srcfd import slab, _capi as capi
    nx, ny = 96, 50
    rng = np.random.default_rng(4)
    Var = rng.uniform(-1, 1, (3, nx + 2, ny + 2)); VarOld = Var + 0.01 * rng.uniform(-1, 1, Var.shape)
    case = O.Case(nx=nx, ny=ny, Re=100.0, dt=1e-3, scheme=scheme, inner_tol=0.0, inner_max=9, order=O.ORDER_JACOBI)

    def prepare(slabs):
        for s in slabs:
            for k in range(3):
                s.h.k_apply_bc(k)
            s.h.k_linear_interpolation(); s.h.k_update_flux()

    one = _make(case, 1, 12, Var=Var, VarOld=VarOld)
    prepare(one)
    V1 = np.zeros_like(Var); F1 = np.zeros((4, nx + 2, ny + 2))
    one[0].h.download(Var=V1, Ff=F1)
    _close(one)
    assert np.array_equal(F1[2, 2:nx + 1, 1:-1], -F1[0, 1:nx, 1:-1]) and np.array_equal(F1[3, 1:-1, 2:ny + 1], -F1[1, 1:-1, 1:ny])
    assert np.count_nonzero(F1[0, 1:nx, 1:-1]) > 0.9 * (nx - 1) * ny
    fn = O.solve_momentum_quick if scheme == "QUICK" else O.solve_momentum_upwind
    sc = capi.SCHEME_QUICK if scheme == "QUICK" else capi.SCHEME_UPWIND
    for k in (0, 1):
        B = V1.copy()
        m = fn(B, VarOld, F1, k, nx, ny, 1.0 / nx, 1.0 / ny, 1e-3, 1.0 / 100.0, (1.0 / nx) * (1.0 / ny), order=O.ORDER_JACOBI, tolerance=0.0, max_iter=9)
        for four in ("0", "1", "w"):                      # both shortcuts / all four stored planes / the west shortcut only
            monkeypatch.setenv("SRCFD_SLAB_FOUR_FACES", "1" if four == "1" else "0")
            monkeypatch.setenv("SRCFD_SLAB_SOUTH_FROM_NORTH", "0" if four == "w" else "1")
            slabs = _make(case, world, 12, Var=Var, VarOld=VarOld)
            prepare(slabs)
            if world > 1:
                assert np.array_equal(_gather(slabs)[2], F1[:, 1:-1])
            n, _ = slab.solve_momentum(slabs, k, sc)
            got = _gather(slabs)[0]
            assert n == m == 9
            assert np.array_equal(got[k], B[k, 1:-1]), (world, scheme, k, four, np.max(np.abs(got[k] - B[k, 1:-1])))
            _close(slabs)


def _cases():
    ldc = O.Case(nx=72, ny=40, Re=100.0, dt=1e-3, scheme="QUICK", order=O.ORDER_JACOBI, inner_max=60)
    bfs = O.bfs_case(64, 36, order=O.ORDER_JACOBI, inner_max=45)
    return [("ldc_quick", ldc, 4), ("bfs_upwind", bfs, 5)]


@pytest.mark.parametrize("world", [1, 2, 3])
@pytest.mark.parametrize("which", [0, 1])
def test_slab_outer_iterations_match_oracle(world, which, pressure_kernel):
    """srcfd_slab_step: whole outer iterations (momentum x2, interpolation, pressure, relaxation, BCs, correction, flux
    update, residual norms) decomposed -- fields bit-equal to the single-domain JACOBI-order oracle."""
    from srcfd import slab
    name, case, its = _cases()[which]
    o = O.OracleSolver(case)
    slabs = _make(case, world, 10, Var=o.Var, VarOld=o.VarOld, Ff=o.Ff)
    for s in slabs:                                        # same start as OracleSolver: _initialize_fields on zero fields
        s.h.initialize_fields(True)
    sweeps = np.zeros(3, dtype=np.int64)
    for _ in range(its):
        sweeps += o.implicit_solve()
        conv, rms_o = o.convergence_check((0.0, 0.0, 0.0))
    slab.step(slabs, its, (0.0, 0.0, 0.0))
    Var, VarOld, Ff = _gather(slabs)
    st = slabs[0].h.status()
    assert st["iterations"] == its
    assert np.array_equal(st["total_sweeps"], sweeps), (name, world, st["total_sweeps"], sweeps)
    assert np.array_equal(Var, o.Var[:, 1:-1]), (name, world, np.max(np.abs(Var - o.Var[:, 1:-1])))
    assert np.array_equal(Ff, o.Ff[:, 1:-1]), (name, world)
    assert np.array_equal(VarOld, o.VarOld[:, 1:-1]), (name, world)
    assert np.allclose(st["rms"], rms_o, rtol=1e-11, atol=0), (st["rms"], rms_o)
    # the boundary rows of the domain live on the first and the last slab
    lo = slabs[0].download_local()[0]; hi = slabs[-1].download_local()[0]
    assert np.array_equal(lo[:, 0], o.Var[:, 0]) and np.array_equal(hi[:, -1], o.Var[:, -1])
    _close(slabs)


def test_slab_step_stops_on_convergence_like_the_single_domain_path():
    from srcfd import slab
    case = O.Case(nx=48, ny=32, Re=100.0, dt=1e-3, scheme="UPWIND", order=O.ORDER_JACOBI, inner_max=40)
    o = O.OracleSolver(case)
    n_o, last, _ = o.solve(200, (1.6, 1.0, 26.0))
    assert 1 < n_o < 200
    slabs = _make(case, 2, 8)
    for s in slabs:
        s.h.initialize_fields(True)
    slab.step(slabs, 200, (1.6, 1.0, 26.0))
    st = slabs[0].h.status()
    assert st["iterations"] == n_o and st["converged"]
    assert np.array_equal(_gather(slabs)[0], o.Var[:, 1:-1])
    _close(slabs)


# ---- two processes, two GPUs: mailboxes mapped with cudaIpc, stores over NVLink ----------------------------------------
def _ipc_worker(rank, world, port, q):
    import os
    import torch, torch.distributed as dist
    from srcfd import slab
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    name, case, its = _cases()[0]
    o = O.OracleSolver(case)
    s = slab.GpuSlab(_params(case, device=rank), world, rank, halo=10)
    slab.attach_distributed(s)
    s.upload_global(Var=o.Var, VarOld=o.VarOld, Ff=o.Ff)
    s.h.initialize_fields(True)
    slab.step([s], its, (0.0, 0.0, 0.0))
    rows = [None] * world
    dist.all_gather_object(rows, s.owned()[0])
    st = s.h.status()
    if rank == 0:
        q.put((np.concatenate(rows, axis=1), st["total_sweeps"]))
    dist.barrier()
    s.close()
    dist.destroy_process_group()


def test_two_gpu_ipc_slab_matches_oracle():
    import socket
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (the same path runs inside bench.py --gpus N: slab_parity)")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_ipc_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs: p.start()
    Var, sweeps = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120); assert p.exitcode == 0
    name, case, its = _cases()[0]
    o = O.OracleSolver(case)
    tot = np.zeros(3, dtype=np.int64)
    for _ in range(its):
        tot += o.implicit_solve(); o.convergence_check((0.0, 0.0, 0.0))
    assert np.array_equal(sweeps, tot)
    assert np.array_equal(Var, o.Var[:, 1:-1])
