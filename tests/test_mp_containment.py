"""Soundness of the kernel's algorithm against an INDEPENDENT 60-digit computation (mpmath), SURVEY.md 4 test plan (c),(d).

GPU == oracle c2 is checked bit for bit elsewhere; that proves host and device agree, not that either is right.  Here the true flow
phi_T, its derivative D phi_T and the frame D phi_T (A_u + A_s E) are computed with a 60-digit Taylor integrator written from the
formulas of f_linearize (ODE_functions.cpp:206,239) -- no code shared with the oracle -- from SAMPLED initial data: corners and
interior points of the initial box p + A_u (x, y), (x, y) in Uxy, and corners of the eigenfunction-error box.  Every sampled true
value must lie in the enclosures the algorithm returns.  The boxes are chosen WIDE (1e-4 .. 1e-3 in x) so that the parts that are
invisible for thin boxes carry weight: the second variation Xi_j over the hull, its closed form to h^4 with the majorant remainder
g5 h^5, the second-order terms collected in R, the box for (Psi(phi) - Psi(x)) W inside a step, and (eight steps, L_- = 1/4 ... 3 1/8)
the propagation of Theta_j / R / C / r across steps including one basis renormalisation."""
import itertools

import numpy as np
import pytest

mp = pytest.importorskip("mpmath")
mp.mp.dps = 60
MPF = mp.mpf


def _taylor_step(n, b, c, z0, P0, h, order=36):
    """One Taylor step of (z, P)' = (f(z), Df(z) P) in 60-digit arithmetic.  z = (u, v), P = 2n x m matrix.  Returns callable eval(t)."""
    m = len(P0[0])
    u = [[z0[i]] for i in range(n)]; v = [[z0[n + i]] for i in range(n)]
    U = [[[P0[i][k]] for k in range(m)] for i in range(n)]; V = [[[P0[n + i][k]] for k in range(m)] for i in range(n)]
    half = MPF(1) / 2

    def cauchy(a, bb, k):
        return mp.fsum(a[j] * bb[k - j] for j in range(k + 1))

    for k in range(order):
        w = [[(u[i][j] - half) if j == 0 else u[i][j] for j in range(k + 1)] for i in range(n)]
        g = [[cauchy(u[i], [(1 - u[i][j]) if j == 0 else -u[i][j] for j in range(k + 1)], q) for q in range(k + 1)] for i in range(n)]
        F, M = [], [[None] * n for _ in range(n)]
        for i in range(n):
            acc = -b[i] * cauchy(g[i], w[i], k)
            if i > 0:
                acc -= c[i - 1] * cauchy(g[i - 1], w[i], k)
            if i < n - 1:
                acc -= c[i] * cauchy(g[i + 1], w[i], k)
            F.append(acc)
        Mk = [[[MPF(0)] * (k + 1) for _ in range(n)] for _ in range(n)]
        for q in range(k + 1):
            for i in range(n):
                d = -b[i] * (3 * g[i][q] - (half if q == 0 else 0))
                if i > 0:
                    d -= c[i - 1] * g[i - 1][q]
                if i < n - 1:
                    d -= c[i] * g[i + 1][q]
                Mk[i][i][q] = d
            for e in range(n - 1):
                o = 2 * c[e] * cauchy(w[e], w[e + 1], q)
                Mk[e][e + 1][q] = o; Mk[e + 1][e][q] = o
        newV = [[mp.fsum(cauchy(Mk[i][l], U[l][kk], k) for l in range(n)) for kk in range(m)] for i in range(n)]
        for i in range(n):
            u[i].append(v[i][k] / (k + 1)); v[i].append(F[i] / (k + 1))
            for kk in range(m):
                U[i][kk].append(V[i][kk][k] / (k + 1)); V[i][kk].append(newV[i][kk] / (k + 1))

    def ev(t):
        t = MPF(t)
        pw = [t ** j for j in range(order + 1)]
        s = lambda ser: mp.fsum(ser[j] * pw[j] for j in range(order + 1))
        z = [s(u[i]) for i in range(n)] + [s(v[i]) for i in range(n)]
        Pm = [[s(U[i][kk]) for kk in range(m)] for i in range(n)] + [[s(V[i][kk]) for kk in range(m)] for i in range(n)]
        return z, Pm
    return ev


def _det(Mx):
    return mp.det(mp.matrix(Mx))


def _inside(val, lo, hi):
    return MPF(lo) <= val <= MPF(hi)


@pytest.mark.parametrize("n,params,xbox,L_minus", [
    (3, [1.0, .98, .96, .04, .02], 1e-4, 2 ** -5),          # ONE full step
    (2, [1.0, .97, .05], 1e-3, 2 ** -5),
    (3, [1.0, .98, .96, .04, -.02], 3e-5, 1.28125),         # 41 steps: one basis renormalisation (after step 32), 9 more steps
    (2, [1.0, .97, -.05], 1e-4, 1.28125),
])
def test_sampled_true_values_lie_in_the_enclosures(pkg, orc, n, params, xbox, L_minus):
    abi = pkg.abi
    D, g = 2 * n, 14
    desc, _ = pkg.setup_front(n, params, abi.SetupOpts.default(L_minus_override=L_minus, xy_nbd=xbox, eig_err=1e-3))
    # put the front away from the equilibrium so that the nonlinearity matters: start where u ~ 0.3 (the descriptor's point is a 1e-6
    # neighbourhood of 0, where the field is almost linear)
    base = np.array([0.30, 0.22, 0.35, 0.27][:n] + [0.15, 0.11, 0.17, 0.13][:n])
    for i in range(D):
        desc.p_i[i].lo = desc.p_i[i].hi = float(base[i])
        desc.Uxy[i].lo, desc.Uxy[i].hi = -xbox * (1 + 0.3 * i), xbox * (1 + 0.2 * i)            # asymmetric, all 2n directions wide
    res, ser = orc.run_front(abi, desc, orc.MODE_C2, series=True)
    assert res.status == 0 and res.steps == int(np.ceil(L_minus * 32))
    if L_minus > 1:
        assert res.steps == 41
    A = [[MPF(desc.A_u[i * abi.CCP_MAX_DIM + j]) for j in range(D)] for i in range(D)]
    b = [MPF(desc.b[i].lo) for i in range(n)]; c = [MPF(desc.c[i].lo) for i in range(n - 1)]
    h = MPF(1) / 32
    rng = np.random.default_rng(5)
    corners = list(itertools.product((0, 1), repeat=D))
    picks = [corners[i] for i in rng.choice(len(corners), size=3 if L_minus > 1 else 8, replace=False)] + [None, "rand"]
    ld = abi.CCP_MAX_N + 2
    for pick in picks:
        if pick is None:
            rho = [MPF(0)] * D
        elif pick == "rand":
            rho = [MPF(desc.Uxy[i].lo) + (MPF(desc.Uxy[i].hi) - MPF(desc.Uxy[i].lo)) * MPF(float(rng.uniform())) for i in range(D)]
        else:
            rho = [MPF(desc.Uxy[i].hi if pick[i] else desc.Uxy[i].lo) for i in range(D)]
        E = [[MPF(desc.Eu_err[j * abi.CCP_MAX_N + k].hi if (j + k + (0 if pick is None else sum(pick if pick != "rand" else (1,)))) % 2 else desc.Eu_err[j * abi.CCP_MAX_N + k].lo)
              for k in range(n)] for j in range(n)]
        z = [MPF(base[i]) + mp.fsum(A[i][j] * rho[j] for j in range(D)) for i in range(D)]
        P = [[A[i][j] for j in range(D)] for i in range(D)]                                   # D phi_t A, all 2n columns
        t_done = MPF(0)
        for s in range(res.steps):
            hs = h if s < res.steps - 1 else MPF(L_minus) - t_done
            ev = _taylor_step(n, b, c, z, P, hs)
            # determinants of the top frame on the grid of this step (first and last step only: the series rows are known there)
            if s in (0, res.steps - 1):
                for gi in (0, 6, g - 1):
                    row = ser[s * g + gi]
                    tl = MPF(row[0]) - t_done if s == 0 else None
                    for frac in (0, MPF(1) / 2, 1):
                        tt = (MPF(row[0]) + (MPF(row[1]) - MPF(row[0])) * frac) - t_done
                        tt = max(MPF(0), min(hs, tt))
                        _, Pt = ev(tt)
                        W = [[Pt[i][k] + mp.fsum(Pt[i][n + j] * E[j][k] for j in range(n)) for k in range(n)] for i in range(n)]
                        dv = _det(W)
                        assert _inside(dv, row[6], row[7]), ("det range", s, gi, float(frac))
                        if frac == 0:
                            assert _inside(dv, row[2], row[3]), ("det left", s, gi)
                        if frac == 1:
                            assert _inside(dv, row[4], row[5]), ("det right", s, gi)
            z, P = ev(hs)
            t_done += hs
        for i in range(D):
            assert _inside(z[i], res.last_frame[i * ld + n + 1].lo, res.last_frame[i * ld + n + 1].hi), ("phi(L_-)", i)
            for k in range(D):
                assert _inside(P[i][k], res.flow_P[i * abi.CCP_MAX_DIM + k].lo, res.flow_P[i * abi.CCP_MAX_DIM + k].hi), ("flow_P", i, k)
            for k in range(n):
                wv = P[i][k] + mp.fsum(P[i][n + j] * E[j][k] for j in range(n))
                assert _inside(wv, res.last_frame[i * ld + k].lo, res.last_frame[i * ld + k].hi), ("last frame", i, k)
    # the enclosures are informative: second-order terms of a box this wide are far above rounding, yet the widths stay of first order
    wid = max(res.last_frame[i * ld + n + 1].hi - res.last_frame[i * ld + n + 1].lo for i in range(D))
    assert wid < 60 * xbox * np.exp(1.2 * L_minus)
