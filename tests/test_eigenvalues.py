"""Host mirror of the reference's eigenvalues.cpp (host/ccp_eigenvalues.hpp): Gershgorin bounds and the
Krawczyk validation of eigenvectors.  Pure host C++ on outward-rounded intervals; the checks inside tests/cpp/test_eig.cpp are
containment of eigenpairs refined independently in long double, the reference's failure code (-344) and the normalisation
|pi_1 v|^2 = 1/(2|lambda|) (eigenvalues.cpp:137-167)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "test_eig")


def test_eigenvalue_mirror_host_only(pkg):
    cmd = ["g++", "-std=c++17", "-O1", "-o", EXE, os.path.join(ROOT, "tests", "cpp", "test_eig.cpp"),
           "-L", pkg.PKG_DIR, "-lccp_b200", "-Wl,-rpath," + pkg.PKG_DIR]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([EXE], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    for tag in ("boundEigenvectors ok: 6 columns validated", "krawczykEigenvector ok", "failure codes ok", "boundEigenvalues ok"):
        assert tag in r.stdout
