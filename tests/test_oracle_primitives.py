"""Interval primitives of the oracle against arbitrary-precision interval arithmetic (mpmath.iv) and exact rationals
(test plan (a)): every result must contain the exact result and be at most 1 ulp wider per bound."""
import ctypes as C
from fractions import Fraction

import numpy as np
import pytest

mpmath = pytest.importorskip("mpmath")


def _ulp_below(x):
    return float(np.nextafter(x, -np.inf))


def _ulp_above(x):
    return float(np.nextafter(x, np.inf))


@pytest.mark.parametrize("op,name", [(0, "add"), (1, "sub"), (2, "mul"), (3, "div")])
def test_ival_ops_contain_exact_and_are_tight(pkg, orc, op, name):
    rng = np.random.default_rng(100 + op)
    for _ in range(400):
        a = np.sort(rng.standard_normal(2) * 10.0 ** rng.integers(-8, 8))
        b = np.sort(rng.standard_normal(2) * 10.0 ** rng.integers(-8, 8))
        if op == 3 and b[0] <= 0 <= b[1]:
            continue
        lo, hi = orc.ival_op(pkg.abi, op, (float(a[0]), float(a[1])), (float(b[0]), float(b[1])))
        A = [Fraction(float(a[0])), Fraction(float(a[1]))]
        B = [Fraction(float(b[0])), Fraction(float(b[1]))]
        if op == 0:
            cand = [A[0] + B[0], A[1] + B[1]]
        elif op == 1:
            cand = [A[0] - B[1], A[1] - B[0]]
        elif op == 2:
            cand = [x * y for x in A for y in B]
        else:
            cand = [x / y for x in A for y in B]
        elo, ehi = min(cand), max(cand)
        assert Fraction(lo) <= elo and ehi <= Fraction(hi), name                      # contains the exact result
        assert Fraction(_ulp_above(lo)) > elo and Fraction(_ulp_below(hi)) < ehi, name  # and is the tightest double enclosure
        # cross-check with mpmath's own interval type
        mpmath.iv.prec = 200
        ma, mb = mpmath.iv.mpf([float(a[0]), float(a[1])]), mpmath.iv.mpf([float(b[0]), float(b[1])])
        mr = [ma + mb, ma - mb, ma * mb, (ma / mb) if op == 3 else ma][op]
        assert lo <= float(mpmath.mpf(mr.a)) and float(mpmath.mpf(mr.b)) <= hi


def test_mt_dot_encloses_exact_sum(pkg, orc):
    """The 4-FMA (mid,t) Cauchy accumulation used by mode c1 and the kernel: the result contains every exact sum of
    products of members of the operand intervals."""
    L = orc.lib(pkg.abi)
    rng = np.random.default_rng(5)
    dp = C.POINTER(C.c_double)
    for _ in range(200):
        n = int(rng.integers(1, 22))
        xm = rng.standard_normal(n) * 10.0 ** rng.integers(-6, 6, n)
        ym = rng.standard_normal(n) * 10.0 ** rng.integers(-6, 6, n)
        xr = np.abs(xm) * 10.0 ** rng.integers(-17, -3, n)
        yr = np.abs(ym) * 10.0 ** rng.integers(-17, -3, n)
        xt = np.nextafter(np.abs(xm) + xr, np.inf)
        yt = np.nextafter(np.abs(ym) + yr, np.inf)
        out = (C.c_double * 2)()
        L.orc_mt_dot(xm.ctypes.data_as(dp), xt.ctypes.data_as(dp), ym.ctypes.data_as(dp), yt.ctypes.data_as(dp), n, out)
        lo = sum(min(Fraction(float(a)) * Fraction(float(b)) for a in (x - r, x + r) for b in (y - s, y + s))
                 for x, r, y, s in zip(xm, xr, ym, yr))
        hi = sum(max(Fraction(float(a)) * Fraction(float(b)) for a in (x - r, x + r) for b in (y - s, y + s))
                 for x, r, y, s in zip(xm, xr, ym, yr))
        assert Fraction(out[0]) <= lo and hi <= Fraction(out[1])
        # not absurdly wide: within 4x of the exact width plus rounding noise
        width = float(hi - lo)
        noise = 64 * np.finfo(float).eps * float(sum(abs(Fraction(float(x)) * Fraction(float(y))) for x, y in zip(xm, ym)))
        assert out[1] - out[0] <= 4 * width + noise


def test_det_encloses_float_det(pkg, orc):
    L = orc.lib(pkg.abi)
    rng = np.random.default_rng(9)
    dp = C.POINTER(C.c_double)
    for n in (2, 3, 4):
        for _ in range(100):
            M = rng.standard_normal((n, n))
            lo = np.ascontiguousarray(M - 1e-12 * np.abs(M))
            hi = np.ascontiguousarray(M + 1e-12 * np.abs(M))
            out = (C.c_double * 2)()
            L.orc_det(lo.ctypes.data_as(dp), hi.ctypes.data_as(dp), n, out)
            d = float(np.linalg.det(M))
            assert out[0] <= d + 1e-9 * abs(d) and d - 1e-9 * abs(d) <= out[1]
            assert out[1] - out[0] < 1e-8 * max(1.0, np.abs(M).max() ** n * n ** n)


def test_mt_scale_and_ipfma_enclose_exact_results(pkg, orc):
    """The two primitives introduced with the first-order tracking: a / k directly in (mid,t) form (scalek / mt_scale) and the fused
    interval * point accumulation (ipfma), against exact rational arithmetic."""
    import ctypes as C
    from fractions import Fraction
    L = orc.lib(pkg.abi)
    L.orc_debug_primitives.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_size_t]
    dp = C.POINTER(C.c_double)
    rng = np.random.default_rng(11)
    n = 400
    x = rng.standard_normal((n, 8)) * np.exp(rng.uniform(-30, 30, (n, 8)))
    x[:, 1] = np.abs(x[:, 0]) * (1 + np.abs(rng.standard_normal(n)) * np.exp(rng.uniform(-40, 2, n)))     # t >= |m|, radius from tiny to large
    x[::7, 1] = np.abs(x[::7, 0])                                                                          # points
    x[:, 2] = rng.integers(1, 22, n)
    out = np.zeros((n, 4))
    assert L.orc_debug_primitives(21, x.ctypes.data_as(dp), out.ctypes.data_as(dp), n) == 0
    for i in range(n):
        m, t, k = Fraction(float(x[i, 0])), Fraction(float(x[i, 1])), int(x[i, 2])
        r = t - abs(m)
        m2, t2 = Fraction(float(out[i, 0])), Fraction(float(out[i, 1]))
        r2 = t2 - abs(m2)
        assert r2 >= 0
        assert m2 - r2 <= (m - r) / k and (m + r) / k <= m2 + r2                  # encloses [m - r, m + r] / k
        assert r2 <= r / k + t / k * Fraction(1, 2 ** 48) + Fraction(1, 2 ** 990)        # and is tight (inflation 2^-50 of t)
    # ipfma: a + x * p
    y = rng.standard_normal((n, 8)) * np.exp(rng.uniform(-10, 10, (n, 8)))
    for a in (0, 2):
        lo, hi = np.minimum(y[:, a], y[:, a + 1]), np.maximum(y[:, a], y[:, a + 1]); y[:, a], y[:, a + 1] = lo, hi
    assert L.orc_debug_primitives(22, y.ctypes.data_as(dp), out.ctypes.data_as(dp), n) == 0
    for i in range(n):
        alo, ahi, xlo, xhi, p = (Fraction(float(v)) for v in y[i, :5])
        cands = [xlo * p, xhi * p]
        lo, hi = alo + min(cands), ahi + max(cands)
        assert Fraction(float(out[i, 0])) <= lo and hi <= Fraction(float(out[i, 1]))
        assert Fraction(float(np.nextafter(out[i, 0], np.inf))) > lo and Fraction(float(np.nextafter(out[i, 1], -np.inf))) < hi   # one rounding per bound
