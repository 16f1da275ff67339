"""Parity of the CUDA path (through the C ABI / ctypes) against the CPU oracle.  `-m gpu`.

Bars: fields BIT-EXACT for every kernel and every sweep order (the kernels evaluate the reference's
expressions in the reference's order, fp64, no FMA contraction); residual sums to 1e-13 relative
(fixed-order tree sums on the device vs the reference's sequential sum).
"""
import os

import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def K():
    from srcfd import kernels
    return kernels


def rnd_state(seed, Nx, Ny, ff_scale=0.05):
    rng = np.random.default_rng(seed)
    Var = rng.uniform(-1, 1, (3, Nx + 2, Ny + 2))
    VarOld = Var + 0.01 * rng.uniform(-1, 1, Var.shape)
    Ff = ff_scale * rng.uniform(-1, 1, (4, Nx + 2, Ny + 2))
    return Var, VarOld, Ff


SIZES = [(1, 1), (2, 3), (13, 9), (64, 48), (33, 130)]


@pytest.mark.parametrize("Nx,Ny", SIZES)
def test_tail_kernels_bit_exact(K, Nx, Ny):
    Var, VarOld, Ff = rnd_state(Nx * 100 + Ny, Nx, Ny)
    dx, dy, dt, rho = 1.3 / Nx, 0.9 / Ny, 2e-3, 1.0
    A, B = Ff.copy(), Ff.copy()
    K.linear_interpolation(Var, A, Nx, Ny, dx, dy); O.linear_interpolation(Var, B, Nx, Ny, dx, dy)
    assert np.array_equal(A, B)
    A, B = Ff.copy(), Ff.copy()
    K.update_flux(Var, A, dt, rho, Nx, Ny, dx, dy); O.update_flux(Var, B, dt, rho, Nx, Ny, dx, dy)
    assert np.array_equal(A, B)
    for k in range(3):
        t = np.array([k % 2, 1, 0, (k + 1) % 2], dtype=np.int32); v = np.array([0.3, -0.2, 1.0, 0.5])
        A, B = Var.copy(), Var.copy()
        K.apply_bc_configured(A, k, Nx, Ny, t, v); O.apply_bc_configured(B, k, Nx, Ny, t, v)
        assert np.array_equal(A, B)
    for k in (0, 1, 2):
        A, B = Var.copy(), Var.copy()
        K.apply_bfs_inlet(A, k, Nx, Ny, 3.0 / Ny, 1.0, 2.0, 1.0); O.apply_bfs_inlet(B, k, Nx, Ny, 3.0 / Ny, 1.0, 2.0, 1.0)
        assert np.array_equal(A, B)
    A, B = Var.copy(), Var.copy()
    K.under_relax_field(A, VarOld, 1, Nx, Ny, 0.5); O.under_relax_field(B, VarOld, 1, Nx, Ny, 0.5)
    assert np.array_equal(A, B)
    A, B = VarOld.copy(), VarOld.copy()
    K.copy_new_to_old(Var, A, 3, Nx, Ny); O.copy_new_to_old(Var, B, 3, Nx, Ny)
    assert np.array_equal(A, B)
    A, B = Var.copy(), Var.copy()
    ra, rb = np.array([0.5, 0.25, 0.125]), np.array([0.5, 0.25, 0.125])
    K.correct_velocity(A, VarOld, dt, rho, Nx, Ny, dx, dy, ra); O.correct_velocity(B, VarOld, dt, rho, Nx, Ny, dx, dy, rb)
    assert np.array_equal(A, B)
    np.testing.assert_allclose(ra, rb, rtol=1e-13)


ORDERS = [("GS_LEX", O.ORDER_GS_LEX), ("JACOBI", O.ORDER_JACOBI), ("RED_BLACK", O.ORDER_RB)]


@pytest.mark.parametrize("Nx,Ny", SIZES + [(600, 20), (1100, 7)])     # the last two span 2 and 3 row bands
@pytest.mark.parametrize("oname,ocode", ORDERS)
def test_inner_solves_bit_exact(K, Nx, Ny, oname, ocode):
    Var, VarOld, Ff = rnd_state(7 + Nx + Ny, Nx, Ny)
    dx, dy = 1.3 / Nx, 0.9 / Ny
    volp, dt, nu, rho = dx * dy, 2e-3, 1 / 250.0, 1.0
    cap = 60
    A, B = Var.copy(), Var.copy()
    n = K.solve_pressure(A, Ff, Nx, Ny, dx, dy, dt, rho, volp, sweep_order=oname, max_iter=cap)
    m = O.solve_pressure(B, Ff, Nx, Ny, dx, dy, dt, rho, volp, order=ocode, max_iter=cap)
    assert n == m and np.array_equal(A, B), (n, m, np.max(np.abs(A - B)))
    for k in (0, 1):
        A, B = Var.copy(), Var.copy()
        n = K.solve_momentum_upwind(A, VarOld, Ff, k, Nx, Ny, dx, dy, dt, nu, volp, sweep_order=oname, max_iter=cap)
        m = O.solve_momentum_upwind(B, VarOld, Ff, k, Nx, Ny, dx, dy, dt, nu, volp, order=ocode, max_iter=cap)
        assert n == m and np.array_equal(A, B), ("upwind", k, n, m, np.max(np.abs(A - B)))
        if oname == "RED_BLACK":
            continue            # undefined for the 9-point QUICK stencil; the library refuses it
        A, B = Var.copy(), Var.copy()
        n = K.solve_momentum_quick(A, VarOld, Ff, k, Nx, Ny, dx, dy, dt, nu, volp, sweep_order=oname, max_iter=cap)
        m = O.solve_momentum_quick(B, VarOld, Ff, k, Nx, Ny, dx, dy, dt, nu, volp, order=ocode, max_iter=cap)
        assert n == m and np.array_equal(A, B), ("quick", k, n, m, np.max(np.abs(A - B)))


def _tiny_state(seed, Nx, Ny):
    """Pressure and fluxes spread over 2^-1074 .. 2^-940: every quotient of the update is denormal or next to it."""
    rng = np.random.default_rng(seed)
    Var = np.zeros((3, Nx + 2, Ny + 2)); Ff = np.zeros((4, Nx + 2, Ny + 2))
    Var[2] = rng.uniform(-1, 1, (Nx + 2, Ny + 2)) * 2.0 ** rng.integers(-1074, -940, (Nx + 2, Ny + 2))
    Ff[:] = rng.uniform(-1, 1, Ff.shape) * 2.0 ** rng.integers(-1074, -960, Ff.shape)
    Var[2][rng.uniform(size=Var[2].shape) < 0.1] = 0.0
    return Var, Ff


def _front_state(Nx, Ny):
    """A field at rest with a small source: the pressure front decays into the denormals, zeros beyond it."""
    Var = np.zeros((3, Nx + 2, Ny + 2)); Ff = np.zeros((4, Nx + 2, Ny + 2))
    Ff[:, 3:9, 4:10] = 1e-3 * np.random.default_rng(3).uniform(-1, 1, (4, 6, 6))
    return Var, Ff


@pytest.mark.parametrize("oname,ocode", ORDERS)
@pytest.mark.parametrize("env", [{}, {"SRCFD_GS3": "0"}, {"SRCFD_K3": "1"}, {"SRCFD_K3": "4"}])
def test_pressure_tiny_denormal_and_rest_fields_bit_exact(K, oname, ocode, env, monkeypatch):
    """The divisions that leave the fast reciprocal sequence (zero, tiny and denormal numerators; denormal quotients with
    their double-rounding ties) take the scaled sequence div_mid: same bits as the host's IEEE division."""
    from srcfd import kernels
    if env and oname != "GS_LEX":
        pytest.skip("the knobs select reference-order kernels")
    for k_, v_ in env.items():
        monkeypatch.setenv(k_, v_)
    for h in kernels._cache.values():
        h.close()
    kernels._cache.clear()
    try:
        for Nx, Ny, lx, ly, cap in ((70, 50, 1.3, 0.9, 25), (150, 230, 1.0, 6 * 230 / 150, 220), (230, 150, 6 * 230 / 150, 1.0, 220)):
            dx, dy = lx / Nx, ly / Ny
            for Var, Ff in ((_tiny_state(Nx, Nx, Ny), _front_state(Nx, Ny)) if cap == 25 else (_front_state(Nx, Ny),)):
                A, B = Var.copy(), Var.copy()
                n = K.solve_pressure(A, Ff, Nx, Ny, dx, dy, 1e-3, 1.0, dx * dy, sweep_order=oname, tolerance=0.0, max_iter=cap)
                m = O.solve_pressure(B, Ff, Nx, Ny, dx, dy, 1e-3, 1.0, dx * dy, order=ocode, tolerance=0.0, max_iter=cap)
                assert n == m and np.array_equal(A, B), (oname, env, Nx, Ny, n, m, np.max(np.abs(A - B)))
                assert np.array_equal(np.signbit(A[2]), np.signbit(B[2]))            # signed zeros too
                if cap > 25 and oname == "JACOBI":          # (in-place orders carry the source further per sweep)
                    a = np.abs(B[2, 1:-1, 1:-1])
                    assert np.count_nonzero((a > 0) & (a < 2.0 ** -1022)) > 50 and np.count_nonzero(a == 0) > 1000
    finally:
        for h in kernels._cache.values():
            h.close()
        kernels._cache.clear()


@pytest.mark.parametrize("env", [{"SRCFD_GS3": "0"}, {"SRCFD_K3": "1"}, {"SRCFD_K3": "2"}, {"SRCFD_K3": "3"}, {"SRCFD_K3": "4"},
                                 {"SRCFD_K3": "4", "SRCFD_NBUF": "2"}, {"SRCFD_SKIP_IDLE": "0"}])
def test_pressure_kernel_generations_bit_exact(K, env, monkeypatch):
    """Both reference-order pressure kernels (banded K-sweep groups; full-height groups with diagonal streams), every
    group size, the smallest boundary ring, sweep counts that are and are not multiples of the group size, and the
    speculative-run rerun when the tolerance is met early: all bit-exact against the oracle."""
    from srcfd import kernels
    for k_, v_ in env.items():
        monkeypatch.setenv(k_, v_)
    for h in kernels._cache.values():
        h.close()
    kernels._cache.clear()                                  # the knobs are read when a handle is created
    try:
        for (Nx, Ny), caps in (((64, 48), (1, 7, 60)), ((33, 130), (5, 61)), ((400, 37), (13,)), ((480, 9), (10,)), ((1, 1), (3,)),
                               ((2, 3), (9,))):
            Var, VarOld, Ff = rnd_state(11 + Nx + Ny, Nx, Ny)
            dx, dy = 1.3 / Nx, 0.9 / Ny
            volp, dt, rho = dx * dy, 2e-3, 1.0
            for cap in caps:
                A, B = Var.copy(), Var.copy()
                n = K.solve_pressure(A, Ff, Nx, Ny, dx, dy, dt, rho, volp, max_iter=cap)
                m = O.solve_pressure(B, Ff, Nx, Ny, dx, dy, dt, rho, volp, max_iter=cap)
                assert n == m and np.array_equal(A, B), (env, Nx, Ny, cap, n, m, np.max(np.abs(A - B)))
        # early convergence: guesses alternately too large / too small exercise the rerun path
        Nx, Ny = 40, 30
        Var, VarOld, Ff = rnd_state(5, Nx, Ny, ff_scale=1e-4)
        Var[2] *= 1e-3
        dx, dy = 1.0 / Nx, 1.0 / Ny
        counts = []
        for tol in (1e-2, 1e-6, 1e-3, 1e-7, 1e-4):
            A, B = Var.copy(), Var.copy()
            n = K.solve_pressure(A, Ff, Nx, Ny, dx, dy, 1e-3, 1.0, dx * dy, tolerance=tol, max_iter=500)
            m = O.solve_pressure(B, Ff, Nx, Ny, dx, dy, 1e-3, 1.0, dx * dy, tolerance=tol, max_iter=500)
            assert n == m and np.array_equal(A, B), (env, tol, n, m)
            counts.append(n)
        assert len(set(counts)) >= 3, counts
        # limits of the full-height kernel: 481 rows and sweep caps >= 4094 take the banded kernel; results must not change
        for (nx2, ny2, cap, tol) in ((481, 6, 9, 1e-6), (20, 20, 4500, 1e-6)):
            V2, _, F2 = rnd_state(3 + nx2, nx2, ny2, ff_scale=1e-3)
            A, B = V2.copy(), V2.copy()
            n = K.solve_pressure(A, F2, nx2, ny2, 1.0 / nx2, 1.0 / ny2, 1e-3, 1.0, 1.0 / (nx2 * ny2), tolerance=tol, max_iter=cap)
            m = O.solve_pressure(B, F2, nx2, ny2, 1.0 / nx2, 1.0 / ny2, 1e-3, 1.0, 1.0 / (nx2 * ny2), tolerance=tol, max_iter=cap)
            assert n == m and np.array_equal(A, B), (env, nx2, ny2, cap, n, m)
        # zero field: the reciprocal fast path is out of range for 0/b and the IEEE path must give the same bits
        Z = np.zeros_like(Var); Fz = Ff.copy()
        A, B = Z.copy(), Z.copy()
        n = K.solve_pressure(A, Fz, Nx, Ny, dx, dy, 1e-3, 1.0, dx * dy, max_iter=6)
        m = O.solve_pressure(B, Fz, Nx, Ny, dx, dy, 1e-3, 1.0, dx * dy, max_iter=6)
        assert n == m and np.array_equal(A, B)
    finally:
        for h in kernels._cache.values():
            h.close()
        kernels._cache.clear()


@pytest.mark.parametrize("env", [{"SRCFD_JTB_H": "8"}, {"SRCFD_JTB_H": "4"}, {"SRCFD_JTB": "0"}])
def test_jacobi_temporally_blocked_pressure_bit_exact(K, env, monkeypatch):
    """JACOBI order (opt-in): the temporally blocked kernel (H sweeps per pass through shared-memory tiles with halos) and
    the one-sweep-per-barrier kernel against the oracle's Jacobi restatement -- partial tiles, planes smaller than a
    tile, sweep caps that are not multiples of H, and early stops inside a pass (the pass is repeated with fewer sweeps)."""
    from srcfd import kernels
    for k_, v_ in env.items():
        monkeypatch.setenv(k_, v_)
    for h in kernels._cache.values():
        h.close()
    kernels._cache.clear()
    try:
        for (Nx, Ny), caps in (((64, 48), (1, 7, 16, 61)), ((33, 130), (5, 24)), ((100, 129), (19,)), ((1, 1), (3,)), ((2, 3), (9,)),
                               ((300, 70), (12,))):
            Var, VarOld, Ff = rnd_state(21 + Nx + Ny, Nx, Ny)
            dx, dy = 1.3 / Nx, 0.9 / Ny
            for cap in caps:
                A, B = Var.copy(), Var.copy()
                n = K.solve_pressure(A, Ff, Nx, Ny, dx, dy, 2e-3, 1.0, dx * dy, sweep_order="JACOBI", max_iter=cap)
                m = O.solve_pressure(B, Ff, Nx, Ny, dx, dy, 2e-3, 1.0, dx * dy, order=O.ORDER_JACOBI, max_iter=cap)
                assert n == m and np.array_equal(A, B), (env, Nx, Ny, cap, n, m, np.max(np.abs(A - B)))
        Nx, Ny = 40, 30
        Var, VarOld, Ff = rnd_state(5, Nx, Ny, ff_scale=1e-4)
        Var[2] *= 1e-3
        dx, dy = 1.0 / Nx, 1.0 / Ny
        counts = []
        for tol in (1e-2, 1e-6, 1e-3, 1e-7, 1e-4, 3e-5):
            A, B = Var.copy(), Var.copy()
            n = K.solve_pressure(A, Ff, Nx, Ny, dx, dy, 1e-3, 1.0, dx * dy, sweep_order="JACOBI", tolerance=tol, max_iter=900)
            m = O.solve_pressure(B, Ff, Nx, Ny, dx, dy, 1e-3, 1.0, dx * dy, order=O.ORDER_JACOBI, tolerance=tol, max_iter=900)
            assert n == m and np.array_equal(A, B), (env, tol, n, m)
            counts.append(n)
        assert len(set(counts)) >= 4, counts
    finally:
        for h in kernels._cache.values():
            h.close()
        kernels._cache.clear()


def test_random_shapes_and_caps_bit_exact(K):
    """Seeded sweep over grid shapes (1..70 rows/columns, the ragged cases of warps, bands and tiles), sweep caps and
    tolerances: pressure in reference and Jacobi order, upwind and QUICK momentum in reference order, against the oracle."""
    rng = np.random.default_rng(2024)
    for trial in range(20):
        Nx, Ny = int(rng.integers(1, 71)), int(rng.integers(1, 71))
        cap = int(rng.integers(1, 40))
        tol = float(10.0 ** rng.uniform(-8, 1))
        Var, VarOld, Ff = rnd_state(1000 + trial, Nx, Ny, ff_scale=float(10.0 ** rng.uniform(-4, -1)))
        dx, dy = float(rng.uniform(0.5, 2.0)) / Nx, float(rng.uniform(0.5, 2.0)) / Ny
        volp, dt, nu, rho = dx * dy, float(10.0 ** rng.uniform(-4, -2)), float(10.0 ** rng.uniform(-3, -1)), 1.0
        for oname, ocode in (("GS_LEX", O.ORDER_GS_LEX), ("JACOBI", O.ORDER_JACOBI)):
            A, B = Var.copy(), Var.copy()
            n = K.solve_pressure(A, Ff, Nx, Ny, dx, dy, dt, rho, volp, sweep_order=oname, tolerance=tol, max_iter=cap)
            m = O.solve_pressure(B, Ff, Nx, Ny, dx, dy, dt, rho, volp, order=ocode, tolerance=tol, max_iter=cap)
            assert n == m and np.array_equal(A, B), ("pressure", oname, trial, Nx, Ny, cap, tol, n, m)
        k = trial & 1
        A, B = Var.copy(), Var.copy()
        n = K.solve_momentum_upwind(A, VarOld, Ff, k, Nx, Ny, dx, dy, dt, nu, volp, tolerance=tol, max_iter=cap)
        m = O.solve_momentum_upwind(B, VarOld, Ff, k, Nx, Ny, dx, dy, dt, nu, volp, tolerance=tol, max_iter=cap)
        assert n == m and np.array_equal(A, B), ("upwind", trial, Nx, Ny, cap, tol, n, m)
        A, B = Var.copy(), Var.copy()
        n = K.solve_momentum_quick(A, VarOld, Ff, k, Nx, Ny, dx, dy, dt, nu, volp, tolerance=tol, max_iter=cap)
        m = O.solve_momentum_quick(B, VarOld, Ff, k, Nx, Ny, dx, dy, dt, nu, volp, tolerance=tol, max_iter=cap)
        assert n == m and np.array_equal(A, B), ("quick", trial, Nx, Ny, cap, tol, n, m)


def test_break_semantics_and_rollback(K):
    """Exact 'stop after the first sweep with rms < tol' behaviour, including the speculative-group
    rollback of the wavefront order (the cached handle carries the previous call's sweep count as its guess)."""
    Nx, Ny = 40, 30
    Var, VarOld, Ff = rnd_state(99, Nx, Ny, ff_scale=0.01)
    dx, dy = 1.0 / Nx, 1.0 / Ny
    volp, dt, nu = dx * dy, 1e-3, 0.01
    counts = []
    for tol in (1e-3, 1e-9, 1e-5, 1e-12, 1e-4, 1e-7):          # guesses alternately too large / too small
        A, B = Var.copy(), Var.copy()
        n = K.solve_momentum_upwind(A, VarOld, Ff, 0, Nx, Ny, dx, dy, dt, nu, volp, tolerance=tol, max_iter=400)
        m = O.solve_momentum_upwind(B, VarOld, Ff, 0, Nx, Ny, dx, dy, dt, nu, volp, tolerance=tol, max_iter=400)
        assert n == m and np.array_equal(A, B), (tol, n, m)
        counts.append(n)
    assert len(set(counts)) > 3


def test_golden_numba_kernels(K, golden_dir):
    """CUDA path against outputs of the reference's own numba kernels (tests/golden/numba_kernels.npz)."""
    g = np.load(os.path.join(golden_dir, "numba_kernels.npz"))
    Nx, Ny = int(g["Nx"]), int(g["Ny"])
    dx, dy, volp, dt, nu, rho = (float(g[k]) for k in ("dx", "dy", "volp", "dt", "nu", "rho"))
    Var, VarOld, Ff = g["Var"], g["VarOld"], g["Ff"]
    A = Ff.copy(); K.linear_interpolation(Var, A, Nx, Ny, dx, dy); assert np.array_equal(A, g["linear_interpolation"])
    A = Ff.copy(); K.update_flux(Var, A, dt, rho, Nx, Ny, dx, dy); assert np.array_equal(A, g["update_flux"])
    A = Var.copy(); K.solve_pressure(A, Ff, Nx, Ny, dx, dy, dt, rho, volp); assert np.array_equal(A, g["solve_pressure"])
    for k in (0, 1):
        A = Var.copy(); K.solve_momentum_upwind(A, VarOld, Ff, k, Nx, Ny, dx, dy, dt, nu, volp)
        assert np.array_equal(A, g[f"solve_momentum_upwind_k{k}"])
        A = Var.copy(); K.solve_momentum_quick(A, VarOld, Ff, k, Nx, Ny, dx, dy, dt, nu, volp)
        assert np.array_equal(A, g[f"solve_momentum_quick_k{k}"])
    A = Var.copy(); r = np.zeros(3); K.correct_velocity(A, VarOld, dt, rho, Nx, Ny, dx, dy, r)
    assert np.array_equal(A, g["correct_velocity"])
    np.testing.assert_allclose(r, g["correct_velocity_residual"], rtol=1e-13)


def _bfs_bc(mod):
    bc = mod.BoundaryConditions()
    bc.u_boundaries['left'] = mod.BoundaryCondition('dirichlet', 0.0)
    bc.u_boundaries['right'] = mod.BoundaryCondition('neumann', 0.0)
    bc.v_boundaries['right'] = mod.BoundaryCondition('neumann', 0.0)
    bc.p_boundaries['right'] = mod.BoundaryCondition('dirichlet', 0.0)
    return bc


def test_golden_numba_solves(golden_dir):
    """Drop-in CFDSolver against the reference's own CFDSolver.solve() outputs (numba_solves.npz)."""
    from srcfd import bfs, ldc
    g = np.load(os.path.join(golden_dir, "numba_solves.npz"))
    s = ldc.CFDSolver(ldc.MeshParameters(nx=24, ny=20), ldc.FluidProperties(Re=100.0),
                      ldc.SolverSettings(dt=1e-3, scheme='QUICK', max_iterations=30), ldc.BoundaryConditions())
    n, _ = s.solve("x", verbose=False, save=False)
    assert n == 30 and np.array_equal(s.Var, g["ldc_24x20_quick_30_Var"])
    assert np.array_equal(s.VarOld, g["ldc_24x20_quick_30_VarOld"]) and np.array_equal(s.Ff, g["ldc_24x20_quick_30_Ff"])
    s = bfs.CFDSolver(bfs.MeshParameters(nx=20, ny=16), bfs.FluidProperties(Re=400.0),
                      bfs.SolverSettings(dt=2e-3, scheme='UPWIND', max_iterations=250), _bfs_bc(bfs))
    s.solve("x", verbose=False, save=False)
    assert np.array_equal(s.Var, g["bfs_20x16_upwind_250_Var"]) and np.array_equal(s.Ff, g["bfs_20x16_upwind_250_Ff"])
    hist = np.array([s.residual_history[k] for k in "uvp"]).T
    np.testing.assert_allclose(hist, g["bfs_20x16_upwind_250_hist"], rtol=1e-12)
    s = bfs.CFDSolver(bfs.MeshParameters(nx=20, ny=16), bfs.FluidProperties(Re=400.0),
                      bfs.SolverSettings(dt=2e-3, scheme='QUICK', max_iterations=40), _bfs_bc(bfs))
    s.solve("x", verbose=False, save=False)
    assert np.array_equal(s.Var, g["bfs_20x16_quick_40_Var"])          # QUICK out-of-plane reads (H4) on the inlet


def test_warm_start_golden(golden_dir):
    from srcfd import ldc
    g = np.load(os.path.join(golden_dir, "numba_solves.npz"))
    w = g["warm_fields"]
    ldc._wf.verbose = False
    s, n, _ = ldc.run_fine_simulation_with_ml_init(100.0, 24, 20, {"u": w[0], "v": w[1], "p": w[2]}, dt=1e-3,
                                                   scheme='QUICK', max_iterations=5, save=False)
    assert n == 5 and np.array_equal(s.Var, g["warm_ldc_24x20_quick_5_Var"])


@pytest.mark.parametrize("order,ocode", ORDERS[:2])
def test_composed_solver_vs_oracle(order, ocode):
    """Whole outer iterations (momentum x2, interpolation, pressure, correction, BCs, flux update,
    convergence bookkeeping) at a size with real pipelining."""
    from srcfd import ldc
    nx, ny, its = 96, 72, 12
    st = ldc.SolverSettings(dt=1e-3, scheme='QUICK', max_iterations=its, sweep_order=order)
    s = ldc.CFDSolver(ldc.MeshParameters(nx=nx, ny=ny), ldc.FluidProperties(Re=100.0), st, ldc.BoundaryConditions())
    n, _ = s.solve("x", verbose=False, save=False)
    o = O.OracleSolver(O.Case(nx=nx, ny=ny, Re=100.0, dt=1e-3, scheme="QUICK", order=ocode))
    m, rms, _ = o.solve(its)
    assert n == m == its
    assert np.array_equal(s.Var, o.Var) and np.array_equal(s.VarOld, o.VarOld) and np.array_equal(s.Ff, o.Ff)
    assert s.total_sweeps.tolist() == o.total_sweeps.tolist()
    np.testing.assert_allclose(np.sqrt(s.residual / (nx * ny)) / 1e-3, rms, rtol=1e-12)


@pytest.mark.slow
def test_full_size_400x400_outer_iterations_bit_exact():
    """BASELINE's own grid: two whole outer iterations of the BFS Re=400 UPWIND case and of the LDC Re=100 QUICK case on
    400 x 400 (1000-sweep pressure solves through the full-height kernel, paired momentum launch), bit for bit against
    the oracle, plus a linearity-free size-independent check: the sweep counters and the residual norms."""
    from srcfd import bfs, ldc
    nx = ny = 400
    s = bfs.CFDSolver(bfs.MeshParameters(nx=nx, ny=ny), bfs.FluidProperties(Re=400.0),
                      bfs.SolverSettings(dt=2e-3, max_iterations=2), _bfs_bc(bfs))
    n, _ = s.solve("x", verbose=False, save=False)
    o = O.OracleSolver(O.bfs_case(nx, ny))
    m, rms, _ = o.solve(2)
    assert n == m == 2 and s.total_sweeps.tolist() == o.total_sweeps.tolist()
    assert np.array_equal(s.Var, o.Var) and np.array_equal(s.Ff, o.Ff) and np.array_equal(s.VarOld, o.VarOld)
    np.testing.assert_allclose(np.sqrt(s.residual / (nx * ny)) / 2e-3, rms, rtol=1e-12)
    s = ldc.CFDSolver(ldc.MeshParameters(nx=nx, ny=ny), ldc.FluidProperties(Re=100.0),
                      ldc.SolverSettings(dt=1e-3, scheme='QUICK', max_iterations=2), ldc.BoundaryConditions())
    n, _ = s.solve("x", verbose=False, save=False)
    o = O.OracleSolver(O.Case(nx=nx, ny=ny, Re=100.0, dt=1e-3, scheme="QUICK"))
    m, rms, _ = o.solve(2)
    assert n == m == 2 and s.total_sweeps.tolist() == o.total_sweeps.tolist()
    assert np.array_equal(s.Var, o.Var) and np.array_equal(s.Ff, o.Ff)


def test_stepwise_api_matches_solve():
    """_implicit_solve / _convergence_check (host-array API) == solve()."""
    from srcfd import bfs
    mk = lambda: bfs.CFDSolver(bfs.MeshParameters(nx=30, ny=24), bfs.FluidProperties(Re=400.0),
                               bfs.SolverSettings(dt=2e-3, max_iterations=4), _bfs_bc(bfs))
    a, b = mk(), mk()
    a.solve("x", verbose=False, save=False)
    for _ in range(4):
        b._implicit_solve()
        conv, rms = b._convergence_check()
    assert np.array_equal(a.Var, b.Var) and np.array_equal(a.VarOld, b.VarOld) and np.array_equal(a.Ff, b.Ff)


@pytest.mark.parametrize("scheme", ["UPWIND", "QUICK"])
def test_bfs_implicit_solve_with_arbitrary_ghosts(scheme):
    """The host-array path uploads whatever the caller holds, ghost cells included.  The paired u/v momentum launch must
    then still give the reference's result: the k=0 inlet pass rewrites part of the v ghost column BEFORE the v solve
    (BFS.py:562), and QUICK's u solve over-reads that column (hazard H4)."""
    from srcfd import bfs
    nx, ny = 40, 30
    rng = np.random.default_rng(3)
    s = bfs.CFDSolver(bfs.MeshParameters(nx=nx, ny=ny), bfs.FluidProperties(Re=400.0),
                      bfs.SolverSettings(dt=2e-3, scheme=scheme, max_iterations=4), _bfs_bc(bfs))
    o = O.OracleSolver(O.bfs_case(nx, ny, scheme=scheme))
    for it in range(3):
        V = 0.1 * rng.uniform(-1, 1, s.Var.shape); Vo = V + 0.01 * rng.uniform(-1, 1, V.shape)
        F = 0.01 * rng.uniform(-1, 1, s.Ff.shape)
        s.Var[...] = V; s.VarOld[...] = Vo; s.Ff[...] = F
        o.Var[...] = V; o.VarOld[...] = Vo; o.Ff[...] = F
        s._implicit_solve(); sw = o.implicit_solve()
        assert s.last_sweeps.tolist() == sw.tolist()
        assert np.array_equal(s.Var, o.Var) and np.array_equal(s.Ff, o.Ff), (scheme, it, np.max(np.abs(s.Var - o.Var)))
        s._implicit_solve(); sw = o.implicit_solve()          # second call: same arrays, now with consistent ghosts
        assert np.array_equal(s.Var, o.Var) and np.array_equal(s.Ff, o.Ff)


@pytest.mark.parametrize("scheme", ["UPWIND", "QUICK"])
def test_ldc_implicit_solve_host_arrays_with_arbitrary_ghosts(scheme):
    """Same for the cavity (no inlet override, no relaxation calls): _implicit_solve on whatever the host arrays hold."""
    from srcfd import ldc
    nx, ny = 36, 44
    rng = np.random.default_rng(8)
    s = ldc.CFDSolver(ldc.MeshParameters(nx=nx, ny=ny), ldc.FluidProperties(Re=100.0),
                      ldc.SolverSettings(dt=1e-3, scheme=scheme, max_iterations=4), ldc.BoundaryConditions())
    o = O.OracleSolver(O.Case(nx=nx, ny=ny, Re=100.0, dt=1e-3, scheme=scheme))
    for it in range(2):
        V = 0.1 * rng.uniform(-1, 1, s.Var.shape); Vo = V + 0.01 * rng.uniform(-1, 1, V.shape)
        F = 0.01 * rng.uniform(-1, 1, s.Ff.shape)
        s.Var[...] = V; s.VarOld[...] = Vo; s.Ff[...] = F
        o.Var[...] = V; o.VarOld[...] = Vo; o.Ff[...] = F
        s._implicit_solve(); sw = o.implicit_solve()
        assert s.last_sweeps.tolist() == sw.tolist()
        assert np.array_equal(s.Var, o.Var) and np.array_equal(s.Ff, o.Ff), (scheme, it, np.max(np.abs(s.Var - o.Var)))


def test_nan_raises_value_error():
    from srcfd import ldc
    s = ldc.CFDSolver(ldc.MeshParameters(nx=16, ny=16), ldc.FluidProperties(Re=100.0),
                      ldc.SolverSettings(dt=1e-3, max_iterations=3), ldc.BoundaryConditions())
    s.Var[0, 5, 5] = np.nan
    with pytest.raises(ValueError, match="NaN/Inf in residuals"):
        s.solve("x", verbose=False, save=False)


def test_bfs_sir_late_case_type():
    """"bfs code given by sir.py":856-861 sets case_type after construction: first BC pass has no inlet."""
    from srcfd import ldc, solver as S
    bc = _bfs_bc(ldc)
    bc.u_boundaries['top'] = ldc.BoundaryCondition('dirichlet', 0.0)
    st = S.BFSSolverSettings(dt=2e-3, scheme='UPWIND', max_iterations=300, relaxation_factors={'u': .5, 'v': .5, 'p': .2})
    s = S.CFDSolver(S.MeshParameters(nx=10, ny=10, lx=10.0, ly=3.0), S.FluidProperties(Re=400), st, bc, relaxed=True)
    s.case_type, s.h, s.step_height, s.Ub = 'BFS', 2.0, 1, 1.0
    s.solve("x", verbose=False, save=False)
    o = O.OracleSolver(O.bfs_case(10, 10), bfs_at_init=False); o.solve(300)
    assert np.array_equal(s.Var, o.Var)


def test_ensemble_concurrent_cases_match_sequential():
    """Several cases on one GPU at once (own streams, capped persistent grids) give the bits they give alone."""
    from srcfd import ensemble as E
    cases = [E.CaseSpec("ldc", 100.0, 48, 40, 6, warm_start=False), E.CaseSpec("ldc2", 300.0, 48, 40, 6, warm_start=False),
             E.CaseSpec("bfs", 400.0, 48, 40, 6, warm_start=False), E.CaseSpec("ldc", 700.0, 48, 40, 6, warm_start=False)]
    seq = [E.run_case(c) for c in cases]
    par = E.run_local(cases, concurrency=3)
    one = E.run_local(cases, concurrency=1)               # one worker: cases 2 and 4 re-parameterise the first case's solver
    for a, b, c in zip(seq, par, one):
        assert a.label == b.label == c.label and a.iterations == b.iterations == c.iterations
        assert a.total_sweeps == b.total_sweeps == c.total_sweeps
        assert np.array_equal(a.fields, b.fields) and np.array_equal(a.fields, c.fields)
    o = O.OracleSolver(O.bfs_case(48, 40)); o.solve(6)
    assert np.array_equal(par[2].fields, np.stack([o.Var[k, 1:-1, 1:-1].T for k in range(3)]))


@pytest.mark.parametrize("nx,ny,Re,its", [(24, 20, 100.0, 1500), (40, 36, 400.0, 600)])
def test_long_run_with_drifting_sweep_counts_kernel_path(nx, ny, Re, its):
    """The converging regime on the whole-GPU kernels (resident one-CTA path switched off): pressure sweep counts fall
    from the cap through the one-sweep-group range of k_solve_gs3 (runs aimed 2 sweeps past the guess and finished from
    the boundary ring, follow-up runs, reruns) -- every outer iteration's break must land on the reference's sweep."""
    from srcfd import ldc
    s = ldc.CFDSolver(ldc.MeshParameters(nx=nx, ny=ny), ldc.FluidProperties(Re=Re),
                      ldc.SolverSettings(dt=1e-3, scheme='QUICK', max_iterations=its), ldc.BoundaryConditions())
    s.resident_solve = False
    n, _ = s.solve("x", verbose=False, save=False)
    o = O.OracleSolver(O.Case(nx=nx, ny=ny, Re=Re, dt=1e-3, scheme="QUICK"))
    m, rms, hist = o.solve(its)
    assert n == m
    assert list(s.total_sweeps) == o.total_sweeps.tolist()
    assert np.array_equal(s.Var, o.Var) and np.array_equal(s.VarOld, o.VarOld) and np.array_equal(s.Ff, o.Ff)
    assert 3 * its < o.total_sweeps[2] < 1000 * its          # neither at the cap throughout nor trivially short


def test_methods_still_run_after_a_converged_solve():
    """A converged (or NaN) solve leaves the device-side stop flag set so that the iterations queued behind it are no-ops;
    the reference's methods always run, so the next _apply_bc_wrapper / _implicit_solve on the same solver must too."""
    from srcfd import ldc
    st = ldc.SolverSettings(dt=1e-3, scheme='QUICK', max_iterations=400,
                            convergence_criteria={'u': 1.6, 'v': 1.0, 'p': 26.0, 'continuity': 26.0})
    s = ldc.CFDSolver(ldc.MeshParameters(nx=48, ny=40), ldc.FluidProperties(Re=100.0), st, ldc.BoundaryConditions())
    s.resident_solve = False                               # the whole-GPU kernels (the handle with the stop flag)
    n, _ = s.solve("x", verbose=False, save=False)
    o = O.OracleSolver(O.Case(nx=48, ny=40, Re=100.0, dt=1e-3, scheme="QUICK"))
    m, _, _ = o.solve(400, (1.6, 1.0, 26.0))
    assert n == m and 1 < n < 400 and np.array_equal(s.Var, o.Var)
    s.bc.u_boundaries['top'] = ldc.BoundaryCondition('dirichlet', 2.0)       # change a BC, as the reference's callers may
    for k in range(3):
        o.p.bc_values[0][2] = 2.0
        s._apply_bc_wrapper(k); o.apply_bc(k)
    assert np.array_equal(s.Var, o.Var) and s.Var[0, 5, -1] != 0.0
    before = s.Var.copy()
    s._implicit_solve(); o.implicit_solve()
    assert not np.array_equal(s.Var, before)
    assert np.array_equal(s.Var, o.Var) and np.array_equal(s.Ff, o.Ff)


def test_red_black_sor_pressure_bit_exact_and_faster_to_tolerance():
    """SURVEY 8f-4 (a better Poisson solver behind the same criterion): red-black SOR, p += omega * R/ap.  Bit-equal to
    the oracle's restatement for every omega, and the same tolerance is reached in a fraction of the sweeps."""
    from srcfd import kernels as K
    Nx, Ny = 48, 40
    rng = np.random.default_rng(7)
    Var = rng.uniform(-1, 1, (3, Nx + 2, Ny + 2)); Ff = 0.01 * rng.uniform(-1, 1, (4, Nx + 2, Ny + 2))
    dx, dy = 1.0 / Nx, 1.0 / Ny
    counts = {}
    try:
        for w in (1.0, 1.5, 1.85):
            O.set_sor_omega(w)
            A, B = Var.copy(), Var.copy()
            n = K.solve_pressure(A, Ff, Nx, Ny, dx, dy, 1e-3, 1.0, dx * dy, sweep_order="RED_BLACK", tolerance=1e-4, max_iter=4000, sor_omega=w)
            m = O.solve_pressure(B, Ff, Nx, Ny, dx, dy, 1e-3, 1.0, dx * dy, order=O.ORDER_RB, tolerance=1e-4, max_iter=4000)
            assert n == m and np.array_equal(A, B), (w, n, m, np.max(np.abs(A - B)))
            counts[w] = n
    finally:
        O.set_sor_omega(1.0)
    assert counts[1.85] * 5 < counts[1.0] and counts[1.5] < counts[1.0]


@pytest.mark.parametrize("omega", [1.0, 1.5])
def test_rb_jacobi_order_composed_solver_vs_oracle(omega):
    """sweep_order = RB_JACOBI (north_star's "Jacobi / red-black-SOR sweeps"): Jacobi momentum (QUICK allowed), red-black
    pressure with the SOR factor -- whole outer iterations bit-equal to the oracle's restatement of that combination."""
    from srcfd import ldc
    nx, ny, its = 40, 36, 12
    st = ldc.SolverSettings(dt=1e-3, scheme='QUICK', max_iterations=its, sweep_order="RB_JACOBI", sor_omega=omega)
    s = ldc.CFDSolver(ldc.MeshParameters(nx=nx, ny=ny), ldc.FluidProperties(Re=100.0), st, ldc.BoundaryConditions())
    n, _ = s.solve("x", verbose=False, save=False)
    try:
        O.set_sor_omega(omega)
        o = O.OracleSolver(O.Case(nx=nx, ny=ny, Re=100.0, dt=1e-3, scheme="QUICK", order=O.ORDER_RB_JACOBI))
        m, _, _ = o.solve(its)
    finally:
        O.set_sor_omega(1.0)
    assert n == m == its
    assert list(s.total_sweeps) == o.total_sweeps.tolist()
    assert np.array_equal(s.Var, o.Var) and np.array_equal(s.Ff, o.Ff)
