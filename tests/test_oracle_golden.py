"""Pins the CPU oracle (both modes) to the only output the reference recorded: plot_det.txt.

The reference has no tests; plot_det.txt holds det_series[i][2] (range enclosure of det of the top frame) for all
15,176 sub-intervals of its n=3 default run.  Criteria: same time grid, every file enclosure meets
ours, file midpoints lie inside ours, our radius <= 2x the file's, and the same two certified sign-change rows.
"""
import os

import numpy as np
import pytest

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "plot_det_golden.npz"))


def _check_against_file(res, ser):
    assert res.status == 0
    assert res.steps == 1084 and res.series_length == 15176 == len(ser)
    t_mid = (ser[:, 0] + ser[:, 1]) / 2
    t_rad = (ser[:, 1] - ser[:, 0]) / 2
    assert np.max(np.abs(t_mid - G["t_mid"])) < 1e-11          # 1083 steps of 2^-5 + one partial step, grid 14
    assert np.max(np.abs(t_rad - G["t_rad"])) < 1e-11
    lo, hi = ser[:, 6], ser[:, 7]
    flo, fhi = G["det_mid"] - G["det_rad"], G["det_mid"] + G["det_rad"]
    assert np.all(np.maximum(lo, flo) <= np.minimum(hi, fhi)), "an enclosure misses the reference's enclosure"
    assert np.all((G["det_mid"] >= lo) & (G["det_mid"] <= hi)), "reference midpoint outside our enclosure"
    ratio = ((hi - lo) / 2) / G["det_rad"]
    assert ratio.max() <= 2.0 and ratio.min() >= 0.4, (ratio.min(), ratio.max())
    # the reference's run certifies 2 conjugate points: midpoint sign changes between rows 10993/10994 and 11463/11464
    assert (res.zero_count, res.failure_count) == (2, 0)
    assert [res.conj_index[i] for i in range(res.n_conj)] == [10993, 11463]
    sgn = np.sign(G["det_mid"])
    assert list(np.nonzero(sgn[1:] != sgn[:-1])[0]) == [10993, 11463]


def test_ref_mode_matches_plot_det(config1_ref):
    _check_against_file(*config1_ref)


def test_c2_mode_matches_plot_det(config1_c2):
    _check_against_file(*config1_c2)


def test_c2_and_ref_agree(config1_ref, config1_c2):
    """The structured algorithm (what the GPU runs) against the reference-structured restatement."""
    rr, rs = config1_ref
    cr, cs = config1_c2
    assert (rr.zero_count, rr.failure_count, rr.steps) == (cr.zero_count, cr.failure_count, cr.steps)
    assert [rr.conj_index[i] for i in range(rr.n_conj)] == [cr.conj_index[i] for i in range(cr.n_conj)]
    assert np.array_equal(rs[:, 0:2], cs[:, 0:2])                       # identical time labels
    for name, col in (("L", 2), ("R", 4), ("V", 6), ("D", 8)):
        rlo, rhi, clo, chi = rs[:, col], rs[:, col + 1], cs[:, col], cs[:, col + 1]
        assert np.all(np.maximum(rlo, clo) <= np.minimum(rhi, chi)), name       # both contain the truth
        mid = (rlo + rhi) / 2
        assert np.all((mid >= clo) & (mid <= chi)), name                        # c2 contains ref's point values
    # widths of what the reference pins (range enclosure, derivative): within 2x -- in fact within 0.1 %
    for col in (6, 8):
        ratio = (cs[:, col + 1] - cs[:, col]) / (rs[:, col + 1] - rs[:, col])
        assert 0.5 <= ratio.min() and ratio.max() <= 2.0
        assert abs(np.median(ratio) - 1.0) < 1e-2
    # point evaluations (det at the grid points): the point values (midpoints) agree to 1e-12 relative where the
    # enclosure is that tight; widths at most 1.05x the reference-structured run's on EVERY row (north_star: within 2x) -- first-order
    # tracking of the frame's dependence on the front's box + centres advanced by increments, oracle/c2.hpp; mode c1 is up to 173x wider
    for col in (2, 4):
        rm, cm = (rs[:, col] + rs[:, col + 1]) / 2, (cs[:, col] + cs[:, col + 1]) / 2
        rel = np.abs(rm - cm) / np.abs(rm)
        tight = (rs[:, col + 1] - rs[:, col]) / np.abs(rm) < 1e-13
        assert tight.sum() == 0 or rel[tight].max() < 1e-12
        ratio = (cs[:, col + 1] - cs[:, col]) / (rs[:, col + 1] - rs[:, col])
        assert ratio.max() <= 1.05 and np.median(ratio) < 1.01, (ratio.max(), np.median(ratio))


def test_c1_legacy_mode_still_consistent(pkg, orc, config1_desc, config1_c2):
    """Mode c1 (no first-order tracking; kept as a cross-check) gives the same counts and contains c2's point values."""
    cr, cs = config1_c2
    r1, s1 = orc.run_front(pkg.abi, config1_desc[0], orc.MODE_C1, series=True)
    assert (r1.zero_count, r1.failure_count, r1.steps) == (cr.zero_count, cr.failure_count, cr.steps)
    for col in (2, 4, 6, 8):
        mid = (cs[:, col] + cs[:, col + 1]) / 2
        assert np.all((mid >= s1[:, col]) & (mid <= s1[:, col + 1]))


def test_first_and_last_frame_consistent(config1_ref, config1_c2, pkg):
    rr, _ = config1_ref
    cr, _ = config1_c2
    for name, ld, ncol in (("last_frame", 6, 5), ("first_frame", 5, 4)):
        a, b = getattr(rr, name), getattr(cr, name)
        for row in range(6):
            for col in range(ncol):
                x, y = a[row * ld + col], b[row * ld + col]
                assert max(x.lo, y.lo) <= min(x.hi, y.hi), (name, row, col)     # enclosures of the same quantity meet
    # first frame = A_u [I; Err]: exact initial data in both modes
    for row in range(6):
        for col in range(3):
            x, y = rr.first_frame[row * 5 + col], cr.first_frame[row * 5 + col]
            assert abs(x.lo - y.lo) <= 1e-15 * (1 + abs(x.lo)) and abs(x.hi - y.hi) <= 1e-15 * (1 + abs(x.hi))


@pytest.mark.parametrize("n,params,expected", [
    (2, [1.0, .97, .05], 1),            # n=2 default, heteroclinic.cpp:518-520  [float-probe: 1 zero at t~20.36]
    (2, [1.0, .97, -.05], 0),
    (3, [1.0, .98, .96, .04, -.02], 1),  # sign variants of the n=3 default
    (3, [1.0, .98, .96, -.04, .02], 1),
    (3, [1.0, .98, .96, -.04, -.02], 0),
    (4, [1.0, .98, .96, .94, .04, .02, .03], 3),   # chain extension, n=4
])
def test_counts_other_configs_c2(pkg, orc, n, params, expected):
    desc, info = pkg.setup_front(n, params)
    assert info.converged == 1
    res, _ = orc.run_front(pkg.abi, desc, orc.MODE_C2)
    assert res.status == 0 and res.failure_count == 0
    assert res.zero_count == expected


def test_ref_and_c2_counts_n2(pkg, orc):
    desc, _ = pkg.setup_front(2, [1.0, .97, .05])
    r, _ = orc.run_front(pkg.abi, desc, orc.MODE_REF, threads=2)
    c, _ = orc.run_front(pkg.abi, desc, orc.MODE_C2)
    assert (r.status, r.zero_count, r.failure_count, r.steps) == (c.status, c.zero_count, c.failure_count, c.steps) == (0, 1, 0, 864)
    assert [r.conj_index[i] for i in range(r.n_conj)] == [c.conj_index[i] for i in range(c.n_conj)]


def test_edge_cases(pkg, orc):
    abi = pkg.abi
    desc, _ = pkg.setup_front(3, [1.0, .98, .96, .04, .02], abi.SetupOpts.default(L_minus_override=0.01))
    for mode in (orc.MODE_REF, orc.MODE_C2):
        r, s = orc.run_front(abi, desc, mode, series=True)          # a single partial step
        assert (r.status, r.steps, r.series_length, r.zero_count, r.failure_count) == (0, 1, 14, 0, 0)
        assert abs(s[-1, 1] - 0.01) < 1e-15
    bad = abi.FrontDesc.from_buffer_copy(bytes(desc)); bad.grid = 99
    assert orc.run_front(abi, bad, orc.MODE_C2)[0].status == abi.CCP_ST_BAD_DESC
    bad = abi.FrontDesc.from_buffer_copy(bytes(desc)); bad.L_minus = -1.0
    assert orc.run_front(abi, bad, orc.MODE_REF)[0].status == abi.CCP_ST_BAD_DESC
    # a point far from the front leaves the a-priori box: enclosure failure, not garbage
    far = abi.FrontDesc.from_buffer_copy(bytes(desc)); far.L_minus = 2.0
    for i in range(3):
        far.p_i[i].lo, far.p_i[i].hi = 3.0, 3.0
    assert orc.run_front(abi, far, orc.MODE_C2)[0].status == abi.CCP_ST_ENCLOSURE


@pytest.mark.parametrize("n,params,bound", [
    (2, [1.0, .97, .05], 1.25),
    (3, [1.0, .98, .96, .04, -.02], 1.25),
    (4, [1.0, .98, .96, .94, .04, .02, .03], 1.25),
    (3, "sweep13", 1.25),           # reference sweep (Construct_ParameterList 24 x 4), front 13
    (3, "sweep50", 1.25),
    (3, "sweep0", 1.25),            # one of the fronts that were 2.3x / 3.7x (last frame) before the centres advanced by increments
    (3, "sweep95", 1.25),
])
def test_enclosure_widths_within_bound_of_ref(pkg, orc, n, params, bound):
    """north_star: every enclosure contains the reference's values and is at most 2x as wide.  'Reference' = oracle mode ref
    (the reference-structured restatement); checked for the kernel's twin c2 on det at grid points (L, R), the range and
    derivative enclosures (V, D) and the last frame."""
    if isinstance(params, str):
        params = list(pkg.parameter_list(24, 4, 0.02, 0.06)[int(params[5:])])
    desc, _ = pkg.setup_front(n, params)
    rr, rs = orc.run_front(pkg.abi, desc, orc.MODE_REF, series=True, threads=n)
    cr, cs = orc.run_front(pkg.abi, desc, orc.MODE_C2, series=True)
    assert (rr.status, rr.zero_count, rr.failure_count) == (cr.status, cr.zero_count, cr.failure_count)
    for col in (2, 4, 6, 8):
        wr, wc = rs[:, col + 1] - rs[:, col], cs[:, col + 1] - cs[:, col]
        mid = (rs[:, col] + rs[:, col + 1]) / 2
        assert np.all((mid >= cs[:, col]) & (mid <= cs[:, col + 1]))
        assert (wc / wr).max() <= (bound if col in (2, 4) else 1.01), (col, (wc / wr).max())
    lr = np.array([[x.lo, x.hi] for x in rr.last_frame]); lc = np.array([[x.lo, x.hi] for x in cr.last_frame])
    wr, wc = lr[:, 1] - lr[:, 0], lc[:, 1] - lc[:, 0]
    m = wr > 0
    assert np.all((lr[m].mean(axis=1) >= lc[m, 0]) & (lr[m].mean(axis=1) <= lc[m, 1]))
    assert (wc[m] / wr[m]).max() <= bound


def test_sweep1024_counts_and_pass_fail_equal_ref(pkg, orc):
    """north_star: conjugate-point counts and validation pass/fail must match exactly.  All 1,024 fronts of the bench sweep, kernel twin
    c2 against the committed results of the reference-structured oracle (tests/golden/make_sweep_golden.py; 16 CPU-minutes)."""
    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sweep1024_ref.npz"))
    descs, _ = pkg.setup_sweep(3, pkg.parameter_list(128, 8, 0.02, 0.06))
    res, _ = orc.run_batch(pkg.abi, descs, orc.MODE_C2, threads=0)
    st = np.array([r.status for r in res]); z = np.array([r.zero_count for r in res]); f = np.array([r.failure_count for r in res]); steps = np.array([r.steps for r in res])
    assert np.array_equal(st, G["status"]) and np.array_equal(steps, G["steps"])
    assert np.array_equal(z, G["zero_count"])
    assert np.array_equal(f > 0, G["failure_count"] > 0)                 # certified / not certified: identical on every front
    assert np.abs(f - G["failure_count"]).max() <= 1                     # failing fronts: same failing sub-intervals up to one
    assert int((f > 0).sum()) == 23 and sorted(np.unique(z[f == 0])) == [0, 1, 2]
