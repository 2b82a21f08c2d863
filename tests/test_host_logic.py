"""Host-side pieces either side of the hot path: sweep generator, float shooting, sharding, return codes."""
import math

import numpy as np


def test_parameter_list_matches_reference_formula(pkg):
    # Construct_ParameterList(24,4,.02,.06), heteroclinic.cpp:386-427, restated independently here
    got = pkg.parameter_list(24, 4, 0.02, 0.06)
    exp = []
    step = (0.06 - 0.02) / 3
    for i in range(24):
        theta = (i + .5) * 2 * math.pi / 24
        for j in range(4):
            r = 0.02 + step * j
            exp.append([1, .975, .95, r * math.cos(theta), r * math.sin(theta)])
            theta += math.pi / (24 * 4)
    assert got.shape == (96, 5)
    assert np.allclose(got, np.array(exp), rtol=0, atol=1e-15)
    assert not np.any(got[:, 3:] == 0.0)


def test_setup_reproduces_reference_L_minus(pkg):
    """L_- = sup(2*T*0.84) of the n=3 default run is recorded in plot_det.txt (last row)."""
    desc, info = pkg.setup_front(3, [1.0, .98, .96, .04, .02])
    assert info.converged == 1
    assert abs(desc.L_minus - 33.84706312634513) < 5e-9
    assert desc.order == 20 and desc.step_log2 == 5 and desc.grid == 14
    A = np.array(desc.A_u).reshape(8, 8)[:6, :6]
    # symplectic normalisation (utils.cpp:456-495): omega(p_j, q_j) = 1, det of the top unstable block = (8 sqrt(det K))^-1/2
    J = np.block([[np.zeros((3, 3)), -np.eye(3)], [np.eye(3), np.zeros((3, 3))]])   # omega(V,W) = V.JW, utils.cpp:433-444
    for j in range(3):
        assert abs(A[:, j] @ J @ A[:, 5 - j] - 1.0) < 1e-12
    assert abs(np.linalg.det(A[:3, :3]) - 0.6040554841315) < 1e-9
    # p_i lies on the unstable eigenspace: p_i = A_u (x,0)
    p = np.array([0.5 * (desc.p_i[i].lo + desc.p_i[i].hi) for i in range(6)])
    x = np.linalg.solve(A, p)
    assert np.max(np.abs(x[3:])) < 1e-18 and abs(np.max(np.abs(x[:3])) - 1e-6) < 2e-7


def test_setup_n2_and_n4(pkg):
    d2, i2 = pkg.setup_front(2, [1.0, .97, .05])
    assert i2.converged == 1 and abs(i2.T - 16.0571) < 1e-3 and abs(d2.L_minus - 26.9760) < 1e-3
    d4, i4 = pkg.setup_front(4, [1.0, .98, .96, .94, .04, .02, .03])
    assert i4.converged == 1 and abs(d4.L_minus - 33.92) < 0.05


def test_reference_sweep_all_converge(pkg):
    descs, infos = pkg.setup_sweep(3, pkg.parameter_list(24, 4))
    assert all(i.converged == 1 for i in infos)
    L = np.array([d.L_minus for d in descs])
    assert 33.4 < L.min() and L.max() < 36.0          # 33.50 .. 35.94


def test_shard_partition_is_exact(pkg):
    for world in (1, 2, 4, 8):
        seen = sorted(i for r in range(world) for i in pkg.shard(1000, r, world))
        assert seen == list(range(1000))


def test_frame_det_return_codes(pkg):
    r = pkg.abi.FrontResult()
    r.status, r.failure_count, r.zero_count = 0, 0, 2
    assert pkg.frame_det_code(r) == 2
    r.failure_count = 3
    assert pkg.frame_det_code(r) == -3                  # propagateManifold.cpp:178-183
    r.status = pkg.abi.CCP_ST_ENCLOSURE
    assert pkg.frame_det_code(r) == -10                 # exception path, heteroclinic.cpp:460-467


def test_subdivide_front_covers_the_box(pkg):
    desc, _ = pkg.setup_front(3, [1.0, .98, .96, .04, .02])
    subs = pkg.subdivide_front(desc, 4)
    assert len(subs) == 64
    lo = min(s.Eu_err[0].lo for s in subs)
    hi = max(s.Eu_err[0].hi for s in subs)
    assert lo <= desc.Eu_err[0].lo and hi >= desc.Eu_err[0].hi
