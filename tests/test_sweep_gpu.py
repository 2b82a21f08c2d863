"""End-to-end GPU tests of the callers either side of the hot path: the sweep driver with the REAL engine and
the reference's two output files, plus the multi-GPU entry point of the library."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sample_parameters_with_the_real_engine(engine, pkg, orc, tmp_path):
    """sweep.SampleParameters (heteroclinic.cpp:447-486) driving the GPU engine in chunks that do not divide the sweep, with a crash in
    the middle and a resume: plot_parameters.txt must equal, line for line, the file written from the oracle's counts."""
    from ccp_b200 import sweep
    params = pkg.parameter_list(12, 2)
    descs, _ = pkg.setup_sweep(3, params, pkg.abi.SetupOpts.default(L_minus_override=27.0))     # long enough for the conjugate points
    path = str(tmp_path / "plot_parameters_frame_count_only.txt")
    calls = {"n": 0}

    def flaky(batch):
        calls["n"] += 1
        if calls["n"] == 3:
            raise RuntimeError("simulated crash")
        return engine.frame_count_batch(batch)

    with pytest.raises(RuntimeError, match="simulated crash"):
        sweep.SampleParameters(flaky, descs, params, path=path, chunk=7, unvalidated=True)
    assert os.path.exists(path + ".journal") and not os.path.exists(path)
    launches = engine.kernel_launches()
    counts = sweep.SampleParameters(engine.frame_count_batch, descs, params, path=path, chunk=7, unvalidated=True)
    assert engine.kernel_launches() - launches == 2                      # 24 fronts, 14 done before the crash: chunks of 7 and 3
    cpu, _ = orc.run_batch(pkg.abi, descs, orc.MODE_C2)
    want = [pkg.frame_det_code(r) if pkg.frame_det_code(r) >= 0 else -1 for r in cpu]
    assert counts == want and set(counts) <= {-1, 0, 1, 2} and sum(c >= 0 for c in counts) >= 20
    ref_path = str(tmp_path / "expected.txt")
    sweep.write_plot_parameters(ref_path, params, want)
    assert open(path).read() == open(ref_path).read()
    rows = [ln.split() for ln in open(path).read().strip().split("\n")]
    assert len(rows) == 24 and all(len(r) == 3 for r in rows) and rows[0][0] == "%.8g" % params[0][3]


def test_plot_det_file_through_the_shim_matches_the_reference_file(pkg, tmp_path):
    """topFrame::makePlot through the C++ shim (frameDet with plotting(true), the reference's own call sequence, topFrame.cpp:177-191)
    writes plot_det.txt; compared with the file the reference ships: same 15,176 rows, same time grid, every enclosure inside the
    reference's, the same two sign changes of the midpoints' column."""
    G = np.load(os.path.join(ROOT, "tests", "golden", "plot_det_golden.npz"))
    src = os.path.join(ROOT, "tests", "cpp", "plot_det_gpu.cpp")
    exe = os.path.join(ROOT, "tests", "cpp", "plot_det_gpu")
    cmd = ["g++", "-std=c++17", "-O1", "-o", exe, src, "-L", pkg.PKG_DIR, "-lccp_b200", "-lccp_setup", "-Wl,-rpath," + pkg.PKG_DIR]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True, cwd=str(tmp_path))
    assert r.returncode == 0, r.stdout + r.stderr
    assert "frameDet = 2" in r.stdout
    ours = np.loadtxt(str(tmp_path / "plot_det.txt"))
    assert ours.shape == (15176, 4)
    assert np.max(np.abs(ours[:, 0] - G["t_mid"])) < 1e-11 and np.max(np.abs(ours[:, 2] - G["t_rad"])) < 1e-11
    lo, hi = ours[:, 1] - ours[:, 3], ours[:, 1] + ours[:, 3]
    flo, fhi = G["det_mid"] - G["det_rad"], G["det_mid"] + G["det_rad"]
    tol = 1e-12 * np.maximum(np.abs(flo), np.abs(fhi))                   # both files print rounded midpoint / radius pairs
    assert np.all((lo >= flo - tol) & (hi <= fhi + tol))
    assert (ours[:, 3] / G["det_rad"]).max() <= 1.0 + 1e-9
    sgn = np.sign(ours[:, 1])
    assert list(np.nonzero(sgn[1:] != sgn[:-1])[0]) == [10993, 11463]


def test_library_multi_gpu_entry_point(pkg):
    """ccp_init_multi / ccp_frame_count_batch_multi / frameDetBatch from one C++ host thread over every GPU of the box (one device
    here unless the test runs under `gpurun --gpus N`): records bit-identical to the single-device run."""
    src = os.path.join(ROOT, "tests", "cpp", "multi_gpu.cpp")
    exe = os.path.join(ROOT, "tests", "cpp", "multi_gpu")
    cmd = ["g++", "-std=c++17", "-O1", "-o", exe, src, "-L", pkg.PKG_DIR, "-lccp_b200", "-lccp_setup", "-Wl,-rpath," + pkg.PKG_DIR]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe, "37"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "records bit-identical to the single-device run" in r.stdout
