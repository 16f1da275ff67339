"""CPU model of k_slab_sweep2's schedule (csrc/slab.cuh): the strip / chunk tiling, the register windows and the input
ring indexed exactly as the kernel's unrolled step loop indexes them, the out-of-plane reads of QUICK at both levels and
the per-sweep sums -- with numpy vectors standing in for the 32 lanes of a warp.  It checks the INDEXING of the kernel
against two JACOBI sweeps of the CPU oracle (the arithmetic itself is the GPU parity tests' job).  Development tool and
CPU test helper (tests/test_sweep2_schedule_cpu.py); never part of the product path.

    python tools/sim_sweep2.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


class K:
    def __init__(s, nx, ny, dx, dy, dt, nu):
        s.nx, s.ny, s.pitch = nx, ny, ny + 2
        s.plane = (nx + 2) * (ny + 2)
        s.volp = dx * dy
        s.dx2, s.dy2 = dx * dx, dy * dy
        s.ap_d = -s.volp * (2.0 / (dx * dx) + 2.0 / (dy * dy))
        s.volp_dt = s.volp / dt
        s.neg_nu = -nu
        s.neg_nu_ap_d = (-nu) * s.ap_d


def _finish(c, vold, Fc, ap_c, Fd, k):
    R = -(k.volp_dt * (c - vold) + Fc + k.neg_nu * Fd)
    ap = k.volp_dt + ap_c + k.neg_nu_ap_d
    return c + R / ap, R


def _fd(c, ip, im, jp, jm, k):
    return k.volp * ((ip - 2.0 * c + im) / k.dx2 + (jp - 2.0 * c + jm) / k.dy2)


def upwind(c, ip, im, jp, jm, vold, fE, fN, fW, fS, k):
    sf = np.zeros_like(c)
    ue = np.where(fE >= 0, c, ip); sf = np.where(fE >= 0, sf + fE, sf)
    uw = np.where(fW >= 0, c, im); sf = np.where(fW >= 0, sf + fW, sf)
    un = np.where(fN >= 0, c, jp); sf = np.where(fN >= 0, sf + fN, sf)
    us = np.where(fS >= 0, c, jm); sf = np.where(fS >= 0, sf + fS, sf)
    Fc = ue * fE + uw * fW + un * fN + us * fS
    return _finish(c, vold, Fc, sf * k.volp, _fd(c, ip, im, jp, jm, k), k)


def quick(c, ip, im, jp, jm, ip2, im2, jp2, jm2, vold, fE, fN, fW, fS, k):
    sf = np.zeros_like(c)

    def face(f, d, u, dd, uu):
        pos = 0.75 * c + 0.375 * d - 0.125 * u
        neg = 0.75 * d + 0.375 * c - 0.125 * dd
        return np.where(f >= 0, pos, neg), np.where(f >= 0, 0.75 * f, 0.375 * f)
    ue, a = face(fE, ip, im, ip2, None); sf = sf + a
    uw, a = face(fW, im, ip, im2, None); sf = sf + a
    un, a = face(fN, jp, jm, jp2, None); sf = sf + a
    us, a = face(fS, jm, jp, jm2, None); sf = sf + a
    Fc = ue * fE + uw * fW + un * fN + us * fS
    return _finish(c, vold, Fc, sf * k.volp, _fd(c, ip, im, jp, jm, k), k)


def shfl_down(v, d):
    out = v.copy(); out[:32 - d] = v[d:]; return out


def shfl_up(v, d):
    out = v.copy(); out[d:] = v[:32 - d]; return out


def sweep2_pass(Var, VarOld, Ff, kpl, src, dst, k, quick_scheme, paired, r0, r1, RB, flat_tail):
    """One launch of k_slab_sweep2.  Var: the flat (3, nx+2, ny+2) state (ghost source G), src/dst: (nx+2, ny+2) planes.
    flat_tail: what a read past the end of plane kpl returns (row 0 of the next plane, or padding)."""
    NB = 2 if quick_scheme else 1
    W = 2 * NB + 1
    UNR = 3 if quick_scheme else 6
    RING = UNR % W == 0
    nx, ny = k.nx, k.ny
    strips = (ny + (32 - 2 * NB) - 1) // (32 - 2 * NB)
    nchunks = (nx + RB - 1) // RB
    G = Var[kpl]
    S1 = 0.0; S2 = 0.0
    lanes = np.arange(32)
    for chunk in range(nchunks):
        for strip in range(strips):
            i_lo = 1 + chunk * RB; i_hi = min(nx, i_lo + RB - 1)
            if i_lo > i_hi:
                continue
            j = 1 - NB + strip * (32 - 2 * NB) + lanes
            jin = (j >= 1) & (j <= ny)
            own = jin & (lanes >= NB) & (lanes <= 31 - NB)
            jc = np.clip(j, 0, ny + 1); ja = np.clip(j, 1, ny)
            lo_row, hi_row = 1 - NB, nx + NB
            c1a, c1b = max(i_lo - NB, 1), min(i_hi + NB, nx)

            def row0(r):
                r = min(max(r, lo_row), hi_row)
                if NB == 2 and r < 0:
                    return G[nx + 1]
                if NB == 2 and r > nx + 1:
                    return flat_tail
                return src[r]
            w0 = [np.zeros(32) for _ in range(W)]; w1 = [np.zeros(32) for _ in range(W)]
            q = [None] * (NB + 1)
            fE_up = np.zeros(32)
            s_first, s_last = i_lo - 2 * NB, i_hi + 2 * NB + 1
            sb = s_first
            while sb <= s_last:
                for p in range(UNR):
                    s = sb + p
                    ra, rb = s - NB, s - 1 - 2 * NB
                    l1 = c1a <= ra <= c1b
                    W0 = (lambda kk: w0[(p + kk + 4 * W) % W]) if RING else (lambda kk: w0[kk + 2 * NB])
                    W1 = (lambda kk: w1[(p + kk + 4 * W) % W]) if RING else (lambda kk: w1[kk + 1 + 2 * NB])
                    x0 = row0(s)[jc].copy()
                    if l1:
                        vjp = src[ra][ja + 1]; vjm = src[ra][ja - 1]
                        if NB == 2:
                            flat = np.concatenate([G.reshape(-1), flat_tail])
                            vjp2 = np.where(ja + 2 <= ny + 1, src[ra][np.minimum(ja + 2, ny + 1)], flat[(ra + 1) * k.pitch])
                            vjm2 = np.where(ja - 2 >= 0, src[ra][np.maximum(ja - 2, 0)], G[ra][ny + 1])
                        inn = dict(vold=VarOld[kpl][ra][ja], fE=Ff[0][ra][ja], fN=Ff[1][ra][ja], fS=Ff[3][ra][ja])
                        fWs = Ff[2][ra][ja]
                    if i_lo <= rb <= i_hi:
                        c = W1(-1 - NB)
                        wjp, wjm = shfl_down(c, 1), shfl_up(c, 1)
                        x = q[p % (NB + 1)]
                        if not quick_scheme:
                            nv, R = upwind(c, W1(-NB), W1(-2 - NB), wjp, wjm, x["vold"], x["fE"], x["fN"], x["fW"], x["fS"], k)
                        else:
                            flat = np.concatenate([G.reshape(-1), flat_tail])
                            wjp2, wjm2 = shfl_down(c, 2), shfl_up(c, 2)
                            wjp2 = np.where(own & (j + 2 > ny + 1), flat[(rb + 1) * k.pitch], wjp2)
                            wjm2 = np.where(own & (j - 2 < 0), G[rb][ny + 1], wjm2)
                            nv, R = quick(c, W1(-NB), W1(-2 - NB), wjp, wjm, W1(-1), W1(-1 - 2 * NB), wjp2, wjm2,
                                          x["vold"], x["fE"], x["fN"], x["fW"], x["fS"], k)
                        dst[rb][j[own]] = nv[own]
                        if r0 <= rb <= r1:
                            S2 += float(np.sum((R * R)[own]))
                    if not RING:
                        for t in range(W - 1):
                            w0[t] = w0[t + 1]
                        w0[2 * NB] = x0
                    else:
                        w0[(p + 4 * W) % W] = x0
                    y = W0(-NB).copy()
                    if l1:
                        inn["fW"] = -fE_up if (paired and ra > c1a) else fWs
                        fE_up = inn["fE"]
                        if not quick_scheme:
                            nv, R = upwind(W0(-NB), W0(1 - NB), W0(-1 - NB), vjp, vjm, inn["vold"], inn["fE"], inn["fN"], inn["fW"], inn["fS"], k)
                        else:
                            nv, R = quick(W0(-NB), W0(1 - NB), W0(-1 - NB), vjp, vjm, W0(0), W0(-2 * NB), vjp2, vjm2,
                                          inn["vold"], inn["fE"], inn["fN"], inn["fW"], inn["fS"], k)
                        y = np.where(jin, nv, y)
                        if i_lo <= ra <= i_hi and r0 <= ra <= r1:
                            S1 += float(np.sum((R * R)[own]))
                        q[p % (NB + 1)] = inn
                    if RING:
                        w1[p % W] = y
                    else:
                        for t in range(W - 1):
                            w1[t] = w1[t + 1]
                        w1[W - 1] = y
                sb += UNR
    return S1, S2


def check(nx, ny, quick_scheme, RB, kpl=0, seed=0, paired=False):
    from oracle import oracle as O
    rng = np.random.default_rng(seed)
    Var = rng.uniform(-1, 1, (3, nx + 2, ny + 2)); VarOld = Var + 0.01 * rng.uniform(-1, 1, Var.shape)
    Ff = 0.002 * rng.uniform(-1, 1, (4, nx + 2, ny + 2))
    if paired:                                                # W/S planes = negated E/N planes of the neighbouring cell
        Ff[2, 1:, :] = -Ff[0, :-1, :]; Ff[3, :, 1:] = -Ff[1, :, :-1]
    dx, dy, dt, nu = 1.0 / nx, 1.0 / ny, 1e-3, 1.0 / 100.0
    k = K(nx, ny, dx, dy, dt, nu)
    fn = O.solve_momentum_quick if quick_scheme else O.solve_momentum_upwind
    B = Var.copy()
    m = fn(B, VarOld, Ff, kpl, nx, ny, dx, dy, dt, nu, dx * dy, order=O.ORDER_JACOBI, tolerance=0.0, max_iter=2)
    assert m == 2
    src = Var[kpl].copy(); dst = Var[kpl].copy()
    tail = Var[kpl + 1].reshape(-1) if kpl < 2 else np.zeros((ny + 2) * 4)
    sweep2_pass(Var, VarOld, Ff, kpl, src, dst, k, quick_scheme, paired, 1, nx, RB, tail)
    err = np.max(np.abs(dst - B[kpl]))
    return err, np.array_equal(dst, B[kpl])


if __name__ == "__main__":
    for quick_scheme in (False, True):
        for (nx, ny, RB) in ((96, 50, 32), (40, 61, 40), (33, 28, 7), (5, 3, 5), (70, 30, 35), (64, 90, 13)):
            for paired in (False, True):
                err, same = check(nx, ny, quick_scheme, RB, kpl=0 if not quick_scheme else 1, paired=paired)
                print(f"{'QUICK ' if quick_scheme else 'UPWIND'} nx={nx:3d} ny={ny:3d} RB={RB:3d} paired={int(paired)}: max err {err:.3e} bit-equal {same}")
                assert err < 1e-13, err
