"""srcfd_coarse_solve_batch (one CTA per case, whole solve() in one launch) against the CPU oracle.  `-m gpu`.

Bar: Var, VarOld, Ff, iteration count, per-field sweep totals and the every-100-iterations history BIT-EXACT /
equal to OracleSolver.solve (the reference's single-thread sweep order); rms values to 1e-12 relative (tree sums
on the device against the reference's sequential sum).
"""
import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _params(case: O.Case):
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
    p.sweep_order = capi.ORDERS["GS_LEX"]
    return p


def _ldc2(nx, ny, Re, **kw):
    c = O.Case(nx=nx, ny=ny, Re=Re, **kw)
    c.bc_values[0][3] = 1.0                       # double lid (PyCFD_ML_accelerated.py:1387-1392)
    return c


CASES = [
    O.Case(nx=10, ny=10, Re=100.0, dt=1e-3, scheme="QUICK"),
    O.Case(nx=10, ny=10, Re=1000.0, dt=1e-3, scheme="QUICK"),
    O.Case(nx=10, ny=10, Re=400.0, dt=1e-3, scheme="UPWIND"),
    _ldc2(10, 10, 300.0, dt=1e-3, scheme="QUICK"),
    O.bfs_case(10, 10),
    O.bfs_case(10, 10, Re=100.0, scheme="QUICK"),
]


def _check(r, i, o, n, rms, hist):
    assert int(r["iterations"][i]) == n
    assert np.array_equal(r["Var"][i], o.Var), np.max(np.abs(r["Var"][i] - o.Var))
    assert np.array_equal(r["VarOld"][i], o.VarOld) and np.array_equal(r["Ff"][i], o.Ff)
    assert r["total_sweeps"][i].tolist() == o.total_sweeps.tolist()
    np.testing.assert_allclose(r["rms"][i], rms, rtol=1e-12)
    assert len(r["hist"][i]) == len(hist)
    np.testing.assert_allclose(r["hist"][i], hist, rtol=1e-12)


def test_batch_of_coarse_cases_bit_exact():
    """The 10x10 coarse stage of six different cases (cavity, double lid, BFS; QUICK and UPWIND) in one launch."""
    from srcfd import _capi as capi
    its = 700
    r = capi.coarse_solve_batch([_params(c) for c in CASES], its, (1e-6, 1e-6, 1e-6), hist_cap=its // 100 + 1)
    assert not r["nan"].any()
    for i, c in enumerate(CASES):
        o = O.OracleSolver(c)
        n, rms, hist = o.solve(its)
        _check(r, i, o, n, rms, hist)


@pytest.mark.parametrize("nx,ny,scheme", [(1, 1, "UPWIND"), (2, 3, "QUICK"), (7, 12, "QUICK"), (13, 9, "UPWIND"),
                                          (24, 20, "QUICK"), (30, 30, "UPWIND")])
def test_ragged_and_largest_grids(nx, ny, scheme):
    from srcfd import _capi as capi
    from srcfd.solver import fits_one_cta
    assert fits_one_cta(nx, ny)
    its = 30
    cs = [O.Case(nx=nx, ny=ny, Re=100.0, dt=1e-3, scheme=scheme), O.bfs_case(nx, ny, scheme=scheme)]
    r = capi.coarse_solve_batch([_params(c) for c in cs], its, (1e-6,) * 3, hist_cap=2)
    for i, c in enumerate(cs):
        o = O.OracleSolver(c)
        n, rms, hist = o.solve(its)
        _check(r, i, o, n, rms, hist)


def test_converges_and_stops_like_the_reference():
    """Loose outer criteria: every case stops at its own iteration, VarOld is NOT refreshed on the converged one
    (LDC.py:498-499), and the other CTAs keep going."""
    from srcfd import _capi as capi
    crit = (5e-2, 5e-2, 5e+1)
    r = capi.coarse_solve_batch([_params(c) for c in CASES], 5000, crit, hist_cap=51)
    assert r["converged"].any()
    for i, c in enumerate(CASES):
        o = O.OracleSolver(c)
        n, rms, hist = o.solve(5000, crit)
        _check(r, i, o, n, rms, hist)
    assert len(set(r["iterations"].tolist())) > 1


def test_resume_equals_one_run():
    from srcfd import _capi as capi
    ps = [_params(c) for c in CASES[:3]]
    a = capi.coarse_solve_batch(ps, 120, (1e-6,) * 3)
    b = capi.coarse_solve_batch(ps, 50, (1e-6,) * 3)
    b = capi.coarse_solve_batch(ps, 70, (1e-6,) * 3, state=(b["Var"], b["VarOld"], b["Ff"]))
    assert np.array_equal(a["Var"], b["Var"]) and np.array_equal(a["Ff"], b["Ff"]) and np.array_equal(a["VarOld"], b["VarOld"])


def test_nan_flag_and_too_large_grid():
    from srcfd import _capi as capi
    p = _params(CASES[0])
    V = np.zeros((1, 3, 12, 12)); V[0, 0, 5, 5] = np.nan
    r = capi.coarse_solve_batch([p], 5, (1e-6,) * 3, state=(V, V.copy(), np.zeros((1, 4, 12, 12))))
    assert r["nan"][0] and r["iterations"][0] == 1
    big = _params(O.Case(nx=64, ny=64))
    with pytest.raises(Exception, match="too large"):
        capi.coarse_solve_batch([big], 1, (1e-6,) * 3)


def test_cfdsolver_small_grid_uses_the_resident_path_and_matches_the_kernel_path():
    """CFDSolver.solve on a 10x10 mesh goes through the one-CTA launch; the whole-GPU kernels give the same bits."""
    from srcfd import bfs, ldc
    its = 250
    mk = lambda: ldc.CFDSolver(ldc.MeshParameters(nx=10, ny=10), ldc.FluidProperties(Re=100.0),
                               ldc.SolverSettings(dt=1e-3, scheme='QUICK', max_iterations=its), ldc.BoundaryConditions())
    a, b = mk(), mk()
    b.resident_solve = False
    na, _ = a.solve("x", verbose=False, save=False)
    nb, _ = b.solve("x", verbose=False, save=False)
    o = O.OracleSolver(O.Case(nx=10, ny=10, Re=100.0, dt=1e-3, scheme="QUICK")); m, _, hist = o.solve(its)
    assert na == nb == m
    for s in (a, b):
        assert np.array_equal(s.Var, o.Var) and np.array_equal(s.VarOld, o.VarOld) and np.array_equal(s.Ff, o.Ff)
        assert list(s.total_sweeps) == o.total_sweeps.tolist()
        np.testing.assert_allclose(np.array([s.residual_history[n] for n in 'uvp']).T, hist, rtol=1e-12)
    np.testing.assert_allclose(a.residual, b.residual, rtol=1e-12)
    assert list(a.last_sweeps) == list(b.last_sweeps)


def test_ensemble_coarse_stage_matches_per_case_workflow():
    from srcfd import bfs, ensemble as E, ldc
    cases = [E.CaseSpec("ldc", 100.0), E.CaseSpec("ldc2", 300.0), E.CaseSpec("bfs", 400.0), E.CaseSpec("ldc", 700.0)]
    got = E.coarse_stage(cases, max_iterations=400)
    for spec, f in zip(cases, got):
        wf = (bfs if spec.kind == "bfs" else ldc)._wf
        wf.verbose = False
        ref = wf.run_coarse_simulation(Re=spec.Re, lr_dim=10, max_iterations=400, bc=E._case_bc(spec), save=False)
        for n in 'uvp':
            assert f[n].shape == (10, 10) and np.array_equal(f[n], ref[n])


def _rel(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b)


def test_reference_golden_outputs_at_the_reference_budget(golden_dir):
    """BASELINE config 2 stage A on the GPU: the reference's own committed outputs -- the table in
    outputs/bfs_Re400_centerline.dat and the u/v/p vectors of the 18 coarse result .h5 files
    (tests/golden/reference_outputs.npz) -- reproduced by 100 000-iteration 10x10 solves, all cases in one launch.
    Tolerances as in tests/test_oracle_golden.py (the file's 6-decimal rounding; the reference's run-to-run scatter)."""
    import os
    from srcfd import _capi as capi
    gr = np.load(os.path.join(golden_dir, "reference_outputs.npz"))
    # "bfs code given by sir.py":856-861 sets case_type after the constructor: first BC pass without the inlet
    init = O.OracleSolver(O.bfs_case(10, 10), bfs_at_init=False)
    sir = capi.coarse_solve_batch([_params(O.bfs_case(10, 10))], 100000, (1e-6,) * 3,
                                  state=(init.Var[None].copy(), init.VarOld[None].copy(), init.Ff[None].copy()))
    dat = gr["centerline_dat"]
    assert np.max(np.abs(sir["Var"][0, 0, 10 // 2, 1:-1] - dat[:, 1])) <= 5.01e-7
    assert np.max(np.abs(sir["Var"][0, 1, 1:-1, 10 // 2] - dat[:, 3])) <= 5.01e-7

    names = list(gr["coarse_names"])
    cases, idx = [O.bfs_case(10, 10)], [None]
    for i, nm in enumerate(names):
        if "bfs" in nm:
            continue
        c = O.Case(nx=10, ny=10, Re=1000.0 if "Re1000" in nm else 800.0, dt=1e-3, scheme="QUICK")
        if nm.startswith("07-11"):
            c.bc_values[0][3] = 1.0
        cases.append(c); idx.append(i)
    r = capi.coarse_solve_batch([_params(c) for c in cases], 100000, (1e-6,) * 3)
    assert not r["nan"].any()
    flat = lambda i: [r["Var"][i, k, 1:-1, 1:-1].T.flatten() for k in range(3)]
    mine, hits = flat(0), 0
    for i, nm in enumerate(names):
        if "bfs_coarse_Re400" in nm:
            g = gr[f"coarse_{i}"]; hits += 1
            assert _rel(mine[0], g[0]) <= 1e-8 and _rel(mine[1], g[1]) <= 1e-8 and _rel(mine[2], g[2]) <= 5e-6, nm
    assert hits == 14
    for j in range(1, len(cases)):
        g, m = gr[f"coarse_{idx[j]}"], flat(j)
        pd, pg = m[2] - m[2].mean(), g[2] - g[2].mean()
        tol = 1e-7 if r["iterations"][j] == 100000 else 1e-3
        assert _rel(m[0], g[0]) <= tol and _rel(m[1], g[1]) <= tol and _rel(pd, pg) <= tol, names[idx[j]]
    assert len(cases) == 5


def test_ensemble_warm_stage_then_concurrent_fine_solves(golden_dir):
    """run_local with SR files: one batched coarse launch + SR passes up front, fine solves concurrently -- the same
    fields as the per-case workflow (coarse solve, SR and fine solve one case after the other)."""
    import os
    from srcfd import bfs, ensemble as E, ldc, sr
    bfs._wf.verbose = ldc._wf.verbose = False
    files = dict(stats=os.path.join(golden_dir, "stats_10to400_multiBC.txt"),
                 encoder=sr.load_model(os.path.join(golden_dir, "encoder10_multiBC.h5")),
                 decoder=sr.synthetic_decoder(seed=0), coarse_iterations=300)
    cases = [E.CaseSpec("ldc", 100.0, 400, 400, 3), E.CaseSpec("bfs", 400.0, 400, 400, 3), E.CaseSpec("ldc2", 500.0, 400, 400, 3)]
    seq = [E.run_case(c, sr_files=files) for c in cases]
    par = E.run_local(cases, concurrency=3, sr_files=files)
    for a, b in zip(seq, par):
        assert a.label == b.label and a.iterations == b.iterations == 3 and a.total_sweeps == b.total_sweeps
        assert np.array_equal(a.fields, b.fields)
        assert np.isfinite(b.fields).all() and np.abs(b.fields).max() > 0
