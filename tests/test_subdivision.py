"""BASELINE config 4: the initial box of a front cut into parts^n sub-boxes (the reference's part() rule, utils.cpp:3-10, 309-321), each
sub-box propagated and counted independently, counts tallied and enclosures hulled per parameter -- on the device in the product
(ccp_frame_count_subdivided), serially in the oracle twin (oracle/subdiv.hpp)."""
import ctypes as C

import numpy as np
import pytest

CONFIG1 = [1.0, 0.98, 0.96, 0.04, 0.02]


def _iv(arr, cnt):
    return np.array([[arr[i].lo, arr[i].hi] for i in range(cnt)])


def _raw(x):
    return np.frombuffer(bytes(x), dtype=np.uint8)


def test_subdivision_rescues_a_front_whose_box_is_too_wide(pkg, orc):
    """Front 42 of the reference's sweep with a neighbourhood 10x the rigorous one (xy_nbd = 1e-11): the undivided box cannot be
    certified (a grid point's determinant enclosure contains 0), 2^3 sub-boxes of the x-part of Uxy can -- all with the same count --
    and the hull of their last frames lies inside the undivided enclosure."""
    abi = pkg.abi
    params = pkg.parameter_list(24, 4, 0.02, 0.06)[42]
    desc, _ = pkg.setup_front(3, params, abi.SetupOpts.default(xy_nbd=1e-11))
    whole, _ = orc.run_front(abi, desc, orc.MODE_C2)
    assert whole.status == 0 and whole.failure_count > 0
    res, sub = orc.subdivided(abi, (abi.FrontDesc * 1)(desc), 2, abi.CCP_SUBDIV_UXY, keep_sub=True)
    r = res[0]
    assert (r.status, r.sub_boxes, r.certified, r.failures) == (0, 8, 8, 0)
    assert r.zero_count == r.zero_min == r.zero_max == 1 and r.steps == whole.steps
    assert {(s.status, s.zero_count, s.failure_count) for s in sub} == {(0, 1, 0)}
    lw, lh = _iv(whole.last_frame, 48), _iv(r.last_frame, 48)
    tol = 1e-9 * np.abs(lw).max(axis=1)
    assert np.all((lh[:, 0] >= lw[:, 0] - tol) & (lh[:, 1] <= lw[:, 1] + tol))
    # the hull IS the hull of the sub-box records
    ls = np.array([_iv(s.last_frame, 48) for s in sub])
    assert np.array_equal(lh[:, 0], ls[:, :, 0].min(axis=0)) and np.array_equal(lh[:, 1], ls[:, :, 1].max(axis=0))


def test_part_rule_covers_the_box(pkg, orc):
    """part(x,k,N) = x.left() + k (x.right()-x.left())/N + (x-x.left())/N in outward-rounded arithmetic: the N pieces cover x, neighbours
    overlap or touch, only the subdivided coordinates change, index rule part_list[q] = (j / N^q) % N."""
    abi = pkg.abi
    desc, _ = pkg.setup_front(3, CONFIG1, abi.SetupOpts.default(L_minus_override=0.25, xy_nbd=3e-12))
    for which, N in ((abi.CCP_SUBDIV_UXY, 5), (abi.CCP_SUBDIV_EU_ERR, 3)):
        res, sub = orc.subdivided(abi, (abi.FrontDesc * 1)(desc), N, which, keep_sub=True)
        assert res[0].sub_boxes == N ** 3 and res[0].certified == N ** 3
    # the pieces themselves, through the first frame of the sub-box records: first_frame = A_u [I; Err] depends on Eu_err only
    N = 3
    res, sub = orc.subdivided(abi, (abi.FrontDesc * 1)(desc), N, abi.CCP_SUBDIV_EU_ERR, keep_sub=True)
    whole, _ = orc.run_front(abi, desc, orc.MODE_C2)
    fw = _iv(whole.first_frame, 40)
    fs = np.array([_iv(s.first_frame, 40) for s in sub])
    hull = np.stack([fs[:, :, 0].min(axis=0), fs[:, :, 1].max(axis=0)], axis=1)
    m = (fw[:, 1] > fw[:, 0]) & (np.arange(40) % 5 < 3)                 # the frame columns (phi' does not depend on Eu_err)
    assert np.all(hull[m, 0] <= fw[m, 0] + 1e-18) and np.all(hull[m, 1] >= fw[m, 1] - 1e-18)          # the pieces cover the box
    assert np.all((hull[m, 1] - hull[m, 0]) <= (fw[m, 1] - fw[m, 0]) * (1 + 1e-9) + 1e-18)           # and nothing more
    assert np.all((fs[:, m, 1] - fs[:, m, 0]).max(axis=0) <= (fw[m, 1] - fw[m, 0]) / N * (1 + 1e-6) + 1e-15)


@pytest.mark.gpu
@pytest.mark.parametrize("which_name,parts", [("UXY", 4), ("EU_ERR", 3)])
def test_subdivided_gpu_bit_identical_to_oracle(engine, pkg, orc, which_name, parts):
    abi = pkg.abi
    which = getattr(abi, "CCP_SUBDIV_" + which_name)
    params = pkg.parameter_list(24, 4, 0.02, 0.06)[[39, 50]]
    descs, _ = pkg.setup_sweep(3, params, abi.SetupOpts.default(xy_nbd=1e-11, L_minus_override=30.0 if which_name == "EU_ERR" else -1.0))
    launches = engine.kernel_launches()
    gpu = engine.frame_count_subdivided(descs, parts, which)
    assert engine.kernel_launches() - launches == 3                     # expand, propagate + count, hull / tally: nothing on the host
    cpu, _ = orc.subdivided(abi, descs, parts, which)
    assert np.array_equal(_raw(gpu), _raw(cpu))
    if which_name == "UXY":                                             # both fronts fail undivided and are certified by 4^3 sub-boxes
        whole = engine.frame_count_batch(descs)
        assert all(w.failure_count > 0 for w in whole)
        assert [(r.certified, r.zero_count) for r in gpu] == [(64, 1), (64, 0)]


@pytest.mark.gpu
def test_config4_at_its_stated_size(engine, pkg, orc):
    """4,096 = 16^3 sub-boxes of the x-part of Uxy of one parameter (config 1, rigorous box), fused propagate + count and device-side
    hull / tally: every sub-box certified with the count of the undivided front, the hull of the frames inside the undivided
    enclosure, and the whole record bit-identical to the oracle twin."""
    abi = pkg.abi
    desc, _ = pkg.setup_front(3, CONFIG1)
    descs = (abi.FrontDesc * 1)(desc)
    gpu = engine.frame_count_subdivided(descs, 16, abi.CCP_SUBDIV_UXY)[0]
    assert (gpu.status, gpu.sub_boxes, gpu.certified, gpu.zero_count, gpu.failures) == (0, 4096, 4096, 2, 0)
    whole = engine.frame_count_batch(descs)[0]
    assert (whole.zero_count, whole.failure_count, whole.steps) == (2, 0, gpu.steps)
    lw, lh = _iv(whole.last_frame, 48), _iv(gpu.last_frame, 48)
    tol = 1e-9 * np.abs(lw).max(axis=1)
    assert np.all((lh[:, 0] >= lw[:, 0] - tol) & (lh[:, 1] <= lw[:, 1] + tol))
    m = (lw[:, 1] > lw[:, 0]) & (np.arange(48) % 6 < 3)                 # frame columns: never wider than undivided (here their width comes
    assert ((lh[m, 1] - lh[m, 0]) / (lw[m, 1] - lw[m, 0])).max() <= 1.0 + 1e-9   # from Eu_err = 2e-5, which this subdivision leaves alone)
    dw, dh = gpu.det_last, whole.det_last
    assert dh.lo - 1e-9 * abs(dh.lo) <= dw.lo and dw.hi <= dh.hi + 1e-9 * abs(dh.hi)
    cpu, _ = orc.subdivided(abi, descs, 16, abi.CCP_SUBDIV_UXY)
    assert np.array_equal(_raw(gpu), _raw(cpu[0]))
