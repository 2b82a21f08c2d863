"""localManifold::boundDFU (localManifold.cpp:347-409): hull over N^n sub-boxes of Ainv Df(p + A w) A.
The reference ships no vectors for it, so the oracle (oracle/dfu.hpp) is checked through properties every valid enclosure has --
the float derivative at random points of the box lies inside, a finer subdivision gives a subset, the undivided box equals the
direct interval evaluation -- and the CUDA kernel is compared with the oracle bit for bit (the hull is order independent)."""
import numpy as np
import pytest


def _float_DF(n, params, A, p, w):
    b, c = np.array(params[:n]), np.array(params[n:])
    x = p + A @ w
    u = x[:n]; wv = u - 0.5
    M = np.zeros((n, n))
    for i in range(n):
        M[i, i] = 3 * b[i] * wv[i] ** 2 - b[i] / 4
    for e in range(n - 1):
        M[e, e] += c[e] * wv[e + 1] ** 2 - c[e] / 4; M[e + 1, e + 1] += c[e] * wv[e] ** 2 - c[e] / 4
        M[e, e + 1] = M[e + 1, e] = 2 * c[e] * wv[e] * wv[e + 1]
    Df = np.block([[np.zeros((n, n)), np.eye(n)], [M, np.zeros((n, n))]])
    return np.linalg.inv(A) @ Df @ A


def _case(pkg, n, params, stable, N, size=2e-2, cone=0.1):
    desc, _ = pkg.setup_front(n, params)
    D = 2 * n
    A = np.array([[desc.A_u[i * pkg.abi.CCP_MAX_DIM + j] for j in range(D)] for i in range(D)])
    p = np.zeros(D)
    U = np.zeros((D, 2))
    big, small = (slice(n, D), slice(0, n)) if stable else (slice(0, n), slice(n, D))
    U[big] = [-size, size]; U[small] = [-cone * size, cone * size]
    return pkg.dfu_desc(n, params, A, p, U, subdivision=N, stable=stable), A, p, U


def _mat(res, n):
    D = 2 * n
    return np.array([[[res.DF[i * 8 + j].lo, res.DF[i * 8 + j].hi] for j in range(D)] for i in range(D)])


CASES = [(3, [1.0, .98, .96, .04, .02]), (2, [1.0, .97, .05]), (4, [1.0, .98, .96, .94, .04, .02, .03])]


@pytest.mark.parametrize("n,params", CASES)
@pytest.mark.parametrize("stable", [False, True])
def test_bound_dfu_oracle_properties(pkg, orc, n, params, stable):
    abi = pkg.abi
    d1, A, p, U = _case(pkg, n, params, stable, 1)
    d5, _, _, _ = _case(pkg, n, params, stable, 5)
    d15, _, _, _ = _case(pkg, n, params, stable, 15 if n < 4 else 8)
    r = orc.bound_dfu(abi, (abi.DfuDesc * 3)(d1, d5, d15))
    assert [q.status for q in r] == [0, 0, 0] and [q.parts for q in r] == [1, 5 ** n, (15 if n < 4 else 8) ** n]
    M1, M5, M15 = (_mat(q, n) for q in r)
    # finer subdivision => subset (up to rounding of the part end points), and it does tighten
    tol = 1e-12 * np.abs(M1).max()
    assert np.all(M5[..., 0] >= M1[..., 0] - tol) and np.all(M5[..., 1] <= M1[..., 1] + tol)
    assert np.all(M15[..., 0] >= M5[..., 0] - tol) and np.all(M15[..., 1] <= M5[..., 1] + tol)
    assert (M15[..., 1] - M15[..., 0]).sum() < 0.9 * (M1[..., 1] - M1[..., 0]).sum()
    # float derivative at random points (and at the corners) of the box
    rng = np.random.default_rng(7)
    pts = [U[:, 0] + (U[:, 1] - U[:, 0]) * rng.random(2 * n) for _ in range(200)] + [U[:, 0], U[:, 1], U.mean(axis=1)]
    for w in pts:
        F = _float_DF(n, params, A, p, w)
        assert np.all((F >= M15[..., 0] - 1e-10) & (F <= M15[..., 1] + 1e-10))
    # bad descriptor
    bad = abi.DfuDesc.from_buffer_copy(bytes(d1)); bad.subdivision = 0
    assert orc.bound_dfu(abi, (abi.DfuDesc * 1)(bad))[0].status == abi.CCP_ST_BAD_DESC


@pytest.mark.gpu
def test_bound_dfu_gpu_bit_identical_to_oracle(engine, pkg, orc):
    abi = pkg.abi
    descs = []
    for n, params in CASES:
        for stable in (False, True):
            for N in (1, 4, 15 if n < 4 else 6):
                descs.append(_case(pkg, n, params, stable, N)[0])
    bad = abi.DfuDesc.from_buffer_copy(bytes(descs[0])); bad.subdivision = 99
    descs.append(bad)
    arr = (abi.DfuDesc * len(descs))(*descs)
    gpu = engine.bound_dfu_batch(arr)
    cpu = orc.bound_dfu(abi, arr)
    a = np.frombuffer(gpu, dtype=np.uint64); b = np.frombuffer(cpu, dtype=np.uint64)
    assert np.array_equal(a, b), "GPU boundDFU differs from the oracle"
    assert gpu[len(descs) - 1].status == abi.CCP_ST_BAD_DESC and all(q.status == 0 for q in gpu[: len(descs) - 1])
