"""The C++ host shim (reference entry-point names above the C ABI) builds and links; on a GPU it runs the
reference's single-front flow (heteroclinic.cpp:299-334) and returns the reference's codes."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "test_shim")


def _build(pkg):
    cmd = ["g++", "-std=c++17", "-O1", "-o", EXE, os.path.join(ROOT, "tests", "cpp", "test_shim.cpp"),
           "-L", pkg.PKG_DIR, "-lccp_b200", "-lccp_setup", "-Wl,-rpath," + pkg.PKG_DIR]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_shim_compiles_and_links(pkg):
    _build(pkg)
    r = subprocess.run([EXE], capture_output=True, text=True)
    assert r.returncode in (0, 77), r.stdout + r.stderr       # 77 = no GPU: there is no CPU fallback to run


@pytest.mark.gpu
def test_shim_runs_frameDet_on_gpu(pkg):
    _build(pkg)
    r = subprocess.run([EXE], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "frameDet = 2" in r.stdout and "Gxy round trip ok" in r.stdout and "boundDFU ok" in r.stdout and "checkProof = 1" in r.stdout and "FindTime:" in r.stdout


SWEEP = os.path.join(ROOT, "tests", "cpp", "bvp_sweep")


@pytest.mark.gpu
def test_bvp_sweep_validates_reference_sweep_on_gpu(pkg):
    """Boundary-value stage for the reference's own 96-point sweep: every batched Krawczyk step is one launch
    of 4 x 96 validated integrations; all 96 connecting orbits must be verified (heteroclinic.cpp:186-281 loop)."""
    cmd = ["g++", "-std=c++17", "-O1", "-o", SWEEP, os.path.join(ROOT, "tests", "cpp", "bvp_sweep.cpp"),
           "-L", pkg.PKG_DIR, "-lccp_b200", "-lccp_setup", "-Wl,-rpath," + pkg.PKG_DIR]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([SWEEP, "24", "4"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "96 connecting orbits verified" in r.stdout
