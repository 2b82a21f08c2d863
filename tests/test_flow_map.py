"""C^1 flow maps of the boundary-value stage (boundaryValueProblem::Gxy / DGxy, boundaryValueProblem.cpp:9-115)
served by the same kernel: image = last_frame column n+1, [D Phi_T] A_i = flow_P.  The reference ships no vectors for them, so the
checks are properties of the exact flow that any valid enclosure must satisfy:
  * a high-accuracy float integration (scipy DOP853) of the front ODE and of its variational equation lies inside the enclosures;
  * D Phi_T is symplectic:  P^T Omega P = A^T Omega A  must be contained in the interval evaluation of the left-hand side;
  * reversibility: the backward map (STABLE = true, field f_minus) of the forward image contains the initial set;
  * the unstable columns agree with the reference-structured oracle (mode ref integrates exactly those n columns).
CPU tests run the kernel's serial twin (oracle c2); the GPU test runs the engine and compares bit for bit with c2."""
import ctypes as C

import numpy as np
import pytest

CONFIG1 = [1.0, 0.98, 0.96, 0.04, 0.02]
T_FLOW = 6.0


def _field(n, b, c):
    def f(t, z):
        u, v = z[:n], z[n:2 * n]
        g = u * (1 - u); w = u - 0.5
        F = -b * g * w
        M = np.zeros((n, n))
        for i in range(n):
            M[i, i] = -b[i] * (3 * g[i] - 0.5)
        for e in range(n - 1):
            F[e] -= c[e] * g[e + 1] * w[e]; F[e + 1] -= c[e] * g[e] * w[e + 1]
            M[e, e] -= c[e] * g[e + 1]; M[e + 1, e + 1] -= c[e] * g[e]
            M[e, e + 1] = M[e + 1, e] = 2 * c[e] * w[e] * w[e + 1]
        P = z[2 * n:].reshape(2 * n, 2 * n)
        dP = np.vstack([P[n:], M @ P[:n]])
        return np.concatenate([v, F, dP.ravel()])
    return f


def _float_flow(n, params, x0, A, T):
    from scipy.integrate import solve_ivp
    b, c = np.array(params[:n]), np.array(params[n:])
    z0 = np.concatenate([x0, A.ravel()])
    sol = solve_ivp(_field(n, b, c), (0.0, T), z0, method="DOP853", rtol=1e-13, atol=1e-18)
    z = sol.y[:, -1]
    return z[:2 * n], z[2 * n:].reshape(2 * n, 2 * n)


def _front_set(pkg, n, params):
    # thin box: these are properties of the flow map itself (the boundary-value stage brings its own boxes), with tolerances sized for it
    desc, _ = pkg.setup_front(n, params, pkg.abi.SetupOpts.default(xy_nbd=1e-19))
    D = 2 * n
    p_i = np.array([[desc.p_i[i].lo, desc.p_i[i].hi] for i in range(D)])
    Uxy = np.array([[desc.Uxy[i].lo, desc.Uxy[i].hi] for i in range(D)])
    A = np.array([[desc.A_u[i * pkg.abi.CCP_MAX_DIM + j] for j in range(D)] for i in range(D)])
    return p_i, A, Uxy


def _imul(a, b):       # interval product of (lo,hi) arrays with one ulp of outward slack
    c = np.stack([a[..., 0] * b[..., 0], a[..., 0] * b[..., 1], a[..., 1] * b[..., 0], a[..., 1] * b[..., 1]])
    return np.stack([np.nextafter(c.min(axis=0), -np.inf), np.nextafter(c.max(axis=0), np.inf)], axis=-1)


def _check_properties(pkg, n, params, run):
    """`run(descs) -> result records`; checks float containment, symplecticity and the reversibility round trip."""
    D = 2 * n
    p_i, A, Uxy = _front_set(pkg, n, params)
    fwd = run([pkg.flow_desc(n, params, p_i, A, Uxy, T_FLOW)])[0]
    assert fwd.status == 0 and fwd.steps == int(T_FLOW * 32)
    img, P = pkg.flow_result(fwd, n)
    assert np.all(img[:, 0] <= img[:, 1]) and np.all(P[..., 0] <= P[..., 1])
    # (1) float flow from the centre of the set
    xf, Pf = _float_flow(n, params, p_i.mean(axis=1), A, T_FLOW)
    tol_x = 1e-9 * np.abs(xf).max(); tol_P = 1e-9 * np.abs(Pf).max()
    assert np.all((xf >= img[:, 0] - tol_x) & (xf <= img[:, 1] + tol_x))
    assert np.all((Pf >= P[..., 0] - tol_P) & (Pf <= P[..., 1] + tol_P))
    assert (img[:, 1] - img[:, 0]).max() < 1e-6 * np.abs(xf).max()           # and the enclosures are not vacuous
    assert ((P[..., 1] - P[..., 0]) / np.abs(Pf).max()).max() < 1e-6
    # (2) symplecticity:  P^T Omega P contains A^T Omega A
    Om = np.block([[np.zeros((n, n)), np.eye(n)], [-np.eye(n), np.zeros((n, n))]])
    target = A.T @ Om @ A
    # Omega P: rows n.. of P on top, minus rows ..n below (exact sign flips)
    OmP = np.concatenate([P[n:], np.stack([-P[:n, :, 1], -P[:n, :, 0]], axis=-1)], axis=0)
    lhs = np.zeros((D, D, 2))
    for i in range(D):
        for j in range(D):
            t = _imul(P[:, i], OmP[:, j])
            lhs[i, j] = [np.nextafter(t[:, 0].sum() - 1e-300, -np.inf) - 8 * np.spacing(np.abs(t).sum()), np.nextafter(t[:, 1].sum(), np.inf) + 8 * np.spacing(np.abs(t).sum())]
    assert np.all((target >= lhs[..., 0]) & (target <= lhs[..., 1]))
    # (3) reversibility: backward map (field f_minus, STABLE = true) of the forward image contains the initial set
    back = run([pkg.flow_desc(n, params, img, np.eye(D), np.zeros((D, 2)), T_FLOW, stable=True)])[0]
    assert back.status == 0
    bimg, _ = pkg.flow_result(back, n, stable=True)
    mid0 = p_i.mean(axis=1)
    assert np.all((mid0 >= bimg[:, 0]) & (mid0 <= bimg[:, 1]))
    assert (bimg[:, 1] - bimg[:, 0]).max() < 1e-3                              # the round trip stays informative
    return fwd


def _orc_runner(pkg, orc):
    def run(descs):
        arr = (pkg.abi.FrontDesc * len(descs))(*descs)
        res, _ = orc.run_batch(pkg.abi, arr, orc.MODE_C2, threads=1)
        return list(res)
    return run


@pytest.mark.parametrize("n,params", [(3, CONFIG1), (2, [1.0, 0.97, 0.05])])
def test_flow_map_properties_oracle(pkg, orc, n, params):
    _check_properties(pkg, n, params, _orc_runner(pkg, orc))


def test_flow_map_unstable_columns_match_ref(pkg, orc):
    """Mode ref integrates the n unstable columns (one 4n-dimensional trajectory each); with Eu_err = 0 its last frame is flow_P[:, :n]."""
    n, D = 3, 6
    p_i, A, Uxy = _front_set(pkg, n, CONFIG1)
    d = pkg.flow_desc(n, CONFIG1, p_i, A, Uxy, T_FLOW)
    d.grid = 2
    rr, _ = orc.run_front(pkg.abi, d, orc.MODE_REF, threads=3)
    cr, _ = orc.run_front(pkg.abi, d, orc.MODE_C2)
    _, P = pkg.flow_result(cr, n)
    ld = pkg.abi.CCP_MAX_N + 2
    for i in range(D):
        for k in range(n):
            a = rr.last_frame[i * ld + k]
            assert max(a.lo, P[i, k, 0]) <= min(a.hi, P[i, k, 1])                           # both contain the truth
            assert P[i, k, 0] <= 0.5 * (a.lo + a.hi) <= P[i, k, 1]
            assert (P[i, k, 1] - P[i, k, 0]) <= 1.25 * (a.hi - a.lo)
        a = rr.last_frame[i * ld + n + 1]; b = cr.last_frame[i * ld + n + 1]
        assert b.lo <= 0.5 * (a.lo + a.hi) <= b.hi and (b.hi - b.lo) <= 1.25 * (a.hi - a.lo)


@pytest.mark.gpu
@pytest.mark.parametrize("n,params", [(3, CONFIG1), (2, [1.0, 0.97, 0.05])])
def test_flow_map_gpu(engine, pkg, orc, n, params):
    def run(descs):
        arr = (pkg.abi.FrontDesc * len(descs))(*descs)
        gpu = engine.frame_count_batch(arr)
        cpu, _ = orc.run_batch(pkg.abi, arr, orc.MODE_C2, threads=1)
        a = np.frombuffer(gpu, dtype=np.uint64); b = np.frombuffer(cpu, dtype=np.uint64)
        assert np.array_equal(a, b), "GPU flow map differs from oracle c2"
        return list(gpu)
    _check_properties(pkg, n, params, run)
