"""SampleParameters driver (heteroclinic.cpp:429-491): chunking, error codes, output format, crash journal + resume.
Host logic only: the GPU call is replaced by a recording stub (the real call is covered by the -m gpu tests)."""
import importlib
import os

import numpy as np
import pytest

pkg = importlib.import_module("computing-conjugate-points_b200")
sweep = importlib.import_module("computing-conjugate-points_b200.sweep")
abi = importlib.import_module("computing-conjugate-points_b200.abi")


def fake_descs(m):
    d = (abi.FrontDesc * m)()
    for i in range(m):
        d[i].n = 3
        d[i].reserved = float(i)          # carries the front id through the stub
    return d


class StubEngine:
    """zero_count = id % 3; id % 7 == 5 -> countZeros failure (-3); id % 11 == 10 -> engine status (-10)."""

    def __init__(self, fail_at_call=None):
        self.calls = []
        self.fail_at_call = fail_at_call

    def frame_count_batch(self, batch):
        if self.fail_at_call is not None and len(self.calls) == self.fail_at_call:
            raise RuntimeError("simulated crash")
        ids = [int(b.reserved) for b in batch]
        self.calls.append(ids)
        res = (abi.FrontResult * len(batch))()
        for k, i in enumerate(ids):
            res[k].zero_count = i % 3
            res[k].steps = i                  # carries the front id into code_of
            res[k].failure_count = 1 if i % 7 == 5 else 0
            res[k].status = abi.CCP_ST_ENCLOSURE if i % 11 == 10 else abi.CCP_OK
        return res


def expected(i):
    return -1 if (i % 7 == 5 or i % 11 == 10) else i % 3


def read_plot(path):
    return [ln.split() for ln in open(path).read().strip().split("\n")]


def test_chunks_codes_and_file_format(tmp_path):
    m = 50
    params = pkg.parameter_list(25, 2)
    eng = StubEngine()
    path = str(tmp_path / "plot_parameters.txt")
    counts = sweep.SampleParameters(eng.frame_count_batch, fake_descs(m), params, path=path, chunk=16, unvalidated=True)
    assert [len(c) for c in eng.calls] == [16, 16, 16, 2]
    assert counts == [expected(i) for i in range(m)]
    rows = read_plot(path)
    assert len(rows) == m
    for i, r in enumerate(rows):
        assert len(r) == 3 and int(r[2]) == expected(i)
        assert float(r[0]) == pytest.approx(params[i][3], rel=1e-7) and float(r[1]) == pytest.approx(params[i][4], rel=1e-7)
        assert r[0] == "%.8g" % params[i][3]                      # file.precision(8), heteroclinic.cpp:476
    assert not os.path.exists(path + ".journal")


def test_crash_keeps_finished_chunks_and_resume_skips_them(tmp_path):
    m = 40
    params = pkg.parameter_list(20, 2)
    path = str(tmp_path / "plot_parameters.txt")
    eng = StubEngine(fail_at_call=2)
    with pytest.raises(RuntimeError, match="simulated crash"):
        sweep.SampleParameters(eng.frame_count_batch, fake_descs(m), params, path=path, chunk=10, unvalidated=True)
    assert not os.path.exists(path)
    key = sweep._journal_key(fake_descs(m), m)
    assert sorted(sweep._read_journal(path + ".journal", key)) == list(range(20))
    assert sweep._read_journal(path + ".journal", sweep._journal_key(fake_descs(m - 1), m - 1)) == {}
    with open(path + ".journal", "a") as f:
        f.write("20 ")                                            # torn line of a crash in the middle of a write
    eng2 = StubEngine()
    counts = sweep.SampleParameters(eng2.frame_count_batch, fake_descs(m), params, path=path, chunk=10, unvalidated=True)
    assert eng2.calls == [list(range(20, 30)), list(range(30, 40))]
    assert counts == [expected(i) for i in range(m)]
    assert [int(r[2]) for r in read_plot(path)] == counts
    assert not os.path.exists(path + ".journal")


def test_resume_false_recomputes_everything(tmp_path):
    m = 12
    params = pkg.parameter_list(6, 2)
    path = str(tmp_path / "plot_parameters.txt")
    with open(path + ".journal", "w") as f:
        f.write("0 7\n1 7\n")
    eng = StubEngine()
    counts = sweep.SampleParameters(eng.frame_count_batch, fake_descs(m), params, path=path, chunk=100, resume=False, unvalidated=True)
    assert eng.calls == [list(range(m))] and counts[0] == expected(0)


def test_argument_errors(tmp_path):
    params = pkg.parameter_list(6, 2)
    eng = StubEngine()
    with pytest.raises(ValueError):
        sweep.SampleParameters(eng.frame_count_batch, fake_descs(5), params, path=str(tmp_path / "p.txt"), unvalidated=True)
    with pytest.raises(ValueError):
        sweep.SampleParameters(eng.frame_count_batch, fake_descs(12), params, path=str(tmp_path / "p.txt"), chunk=0, unvalidated=True)
    short = lambda b: (abi.FrontResult * 1)()
    with pytest.raises(RuntimeError, match="returned 1 records"):
        sweep.SampleParameters(short, fake_descs(12), params, path=str(tmp_path / "p.txt"), unvalidated=True)


def test_journal_of_another_sweep_is_discarded(tmp_path):
    """Resuming with the same path but other descriptors (another list, other boxes, another grid) must not reuse stale counts."""
    m = 20
    params = pkg.parameter_list(10, 2)
    path = str(tmp_path / "plot_parameters.txt")
    eng = StubEngine(fail_at_call=1)
    with pytest.raises(RuntimeError, match="simulated crash"):
        sweep.SampleParameters(eng.frame_count_batch, fake_descs(m), params, path=path, chunk=10, unvalidated=True)
    other = fake_descs(m)
    other[3].grid = 7                                             # same fronts, different computational parameters
    eng2 = StubEngine()
    counts = sweep.SampleParameters(eng2.frame_count_batch, other, params, path=path, chunk=10, unvalidated=True)
    assert eng2.calls == [list(range(10)), list(range(10, 20))]   # nothing was taken from the stale journal
    assert counts == [expected(i) for i in range(m)]
    with open(path + ".journal", "w") as f:                      # a journal without header (older format) is discarded as well
        f.write("0 7\n1 7\n")
    eng3 = StubEngine()
    counts = sweep.SampleParameters(eng3.frame_count_batch, other, params, path=path, chunk=100, unvalidated=True)
    assert eng3.calls == [list(range(m))] and counts[0] == expected(0)


def test_validation_must_be_stated(tmp_path):
    """frame_det_code alone is the frame count on assumed boxes, not the reference's validated code: the caller has to say so
    (unvalidated=True, own file name) or pass a code_of that folds the host proof stages in."""
    params = pkg.parameter_list(6, 2)
    eng = StubEngine()
    with pytest.raises(ValueError, match="unvalidated"):
        sweep.SampleParameters(eng.frame_count_batch, fake_descs(12), params, path=str(tmp_path / "p.txt"))
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        sweep.SampleParameters(eng.frame_count_batch, fake_descs(12), params, unvalidated=True)
        assert os.path.exists("plot_parameters_frame_count_only.txt") and not os.path.exists("plot_parameters.txt")
        host_codes = {i: (-2 if i == 4 else None) for i in range(12)}          # e.g. the Krawczyk proof failed for front 4
        counts = sweep.SampleParameters(eng.frame_count_batch, fake_descs(12), params,
                                        code_of=lambda r: host_codes[r.steps] if host_codes[r.steps] is not None else pkg.frame_det_code(r))
        assert os.path.exists("plot_parameters.txt") and counts[4] == -1
    finally:
        os.chdir(cwd)
