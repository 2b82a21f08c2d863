"""Host mirror of the reference's eigenvalues.cpp (host/ccp_eigenvalues.hpp): Gershgorin bounds and the
Krawczyk validation of eigenvectors.  Pure host C++ on outward-rounded intervals; the checks inside tests/cpp/test_eig.cpp are
containment of eigenpairs refined independently in long double, the reference's failure code (-344) and the normalisation
|pi_1 v|^2 = 1/(2|lambda|) (eigenvalues.cpp:137-167)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "test_eig")


def test_eigenvalue_mirror_host_only(pkg):
    cmd = ["g++", "-std=c++17", "-O1", "-o", EXE, os.path.join(ROOT, "tests", "cpp", "test_eig.cpp"),
           "-L", pkg.PKG_DIR, "-lccp_b200", "-lccp_setup", "-Wl,-rpath," + pkg.PKG_DIR]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([EXE], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    for tag in ("boundEigenvectors ok: 6 columns validated", "krawczykEigenvector ok", "failure codes ok", "boundEigenvalues ok"):
        assert tag in r.stdout


def test_manifold_conditions_mirror_host_only(pkg):
    """host/ccp_manifold_conditions.hpp: localManifold::constructU / checkIsolatingBlock / checkRateCondition / checkConditions
    (localManifold.cpp:4-41, 180-325) around the GPU call boundDFU.  tests/cpp/test_conditions.cpp checks the certified norm bounds
    (spectral norm, logarithmic norm, ml) against long double on 200 random matrices, and the two conditions on the n = 3 defaults
    at the equilibrium 0 with DF enclosed over the whole neighbourhood on the host (accepted at radius 1e-3 for both manifolds,
    rejected at radius 0.6 and for a cone that is too thin)."""
    exe = os.path.join(ROOT, "tests", "cpp", "test_conditions")
    cmd = ["g++", "-std=c++17", "-O1", "-o", exe, os.path.join(ROOT, "tests", "cpp", "test_conditions.cpp"),
           "-L", pkg.PKG_DIR, "-lccp_b200", "-lccp_setup", "-Wl,-rpath," + pkg.PKG_DIR]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    for tag in ("norm bounds ok", "checkConditions ok", "failing neighbourhoods rejected", "helpers ok"):
        assert tag in r.stdout


def test_local_manifold_eig_mirror_host_only(pkg):
    """host/ccp_local_manifold_eig.hpp: localManifold_Eig (localManifold.cpp:426-681) -- computeK, ErrorEigenfunction,
    computeEigenError_minus/plus_infty, checkConjugatePointsBelowLminus: the producer of the descriptor field Eu_err.
    tests/cpp/test_eig_error.cpp: K against the condition number in long double (1e-6), eps against the formula of Section 3
    evaluated independently (2 %), eps ~ r_u, Eu_m_Error_Final symmetric boxes of the size of eps copied into the descriptor,
    (2.12) accepted at radius 1e-3 and rejected where eps > 1."""
    exe = os.path.join(ROOT, "tests", "cpp", "test_eig_error")
    cmd = ["g++", "-std=c++17", "-O1", "-o", exe, os.path.join(ROOT, "tests", "cpp", "test_eig_error.cpp"),
           "-L", pkg.PKG_DIR, "-lccp_b200", "-lccp_setup", "-Wl,-rpath," + pkg.PKG_DIR]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "localManifold_Eig ok" in r.stdout and "below-L_- test 0" in r.stdout


def test_l_plus_stage_mirror_host_only(pkg):
    """host/ccp_l_plus.hpp: the L_+ stage that consumes the hot path's getLastFrame / getFirstFrame output
    (propagateManifold.cpp:192-558, topFrame.cpp:336-351).  tests/cpp/test_l_plus.cpp: Proposition 2.9's arithmetic against the
    same formulas in plain doubles over a scan of eps_beta (one acceptance threshold), compute_epsilon_beta on frames with known
    answers (orthonormal, thick, rank-deficient), projectionGammaBeta round trip, checkFirstFrame / checkL_plus on synthetic
    result records.  The GPU-backed parts (construct_Manifold_at_LPlus, lastEuFrame) are compiled but not run here."""
    exe = os.path.join(ROOT, "tests", "cpp", "test_l_plus")
    cmd = ["g++", "-std=c++17", "-O1", "-o", exe, os.path.join(ROOT, "tests", "cpp", "test_l_plus.cpp"),
           "-L", pkg.PKG_DIR, "-lccp_b200", "-lccp_setup", "-Wl,-rpath," + pkg.PKG_DIR]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    for tag in ("Proposition 2.9 arithmetic ok", "projectionGammaBeta ok", "checkFirstFrame / checkL_plus ok"):
        assert tag in r.stdout


MANIFOLD_TAGS = ("checkConditions (unstable, GPU boundDFU) ok", "computeEigenError_minus_infty ok", "construct_Manifold_at_LPlus ok",
                 "lastEuFrame = 1", "manifold stages on the GPU ok")


def test_manifold_stages_against_oracle_dfu_backend(pkg, orc):
    """tests/cpp/manifold_gpu.cpp (checkConditions, computeEigenError_*, construct_Manifold_at_LPlus, lastEuFrame around the
    boundDFU call) linked against the ORACLE twin of the boundDFU kernel instead of libccp_b200.so: checks the host mirrors and
    their plumbing through ccp_dfu_desc / ccp_dfu_result on a machine without a GPU.  Test infrastructure only."""
    exe = os.path.join(ROOT, "tests", "cpp", "manifold_dry")
    _build_with_oracle_backend(pkg, "manifold_gpu.cpp", exe)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    for tag in MANIFOLD_TAGS:
        assert tag in r.stdout


@pytest.mark.gpu
def test_manifold_stages_on_gpu(pkg):
    exe = os.path.join(ROOT, "tests", "cpp", "manifold_gpu")
    cmd = ["g++", "-std=c++17", "-O1", "-o", exe, os.path.join(ROOT, "tests", "cpp", "manifold_gpu.cpp"),
           "-L", pkg.PKG_DIR, "-lccp_b200", "-lccp_setup", "-Wl,-rpath," + pkg.PKG_DIR]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    for tag in MANIFOLD_TAGS:
        assert tag in r.stdout


def _build_with_oracle_backend(pkg, src, exe):
    bdir = os.path.join(ROOT, "oracle", "_build")
    cmd = ["g++", "-std=c++17", "-O1", "-o", exe, os.path.join(ROOT, "tests", "cpp", src), os.path.join(ROOT, "tests", "cpp", "oracle_dfu_backend.cpp"),
           "-L", pkg.PKG_DIR, "-lccp_b200", "-lccp_setup", "-L", bdir, "-loracle", "-Wl,-rpath," + bdir, "-Wl,-rpath," + pkg.PKG_DIR]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_compute_front_and_conjugate_points_against_oracle_backend(pkg, orc):
    """host/ccp_driver.hpp: computeFrontAndConjugatePoints (heteroclinic.cpp:23-347) composed from the stage mirrors, for the
    parameters of the reference's main() (heteroclinic.cpp:513-531).  The two GPU entry points it uses (ccp_frame_count_batch,
    ccp_bound_dfu_batch) are served by the oracle twins of the kernels (test infrastructure; the float preparation ccp_setup_front
    comes from the product library): manifolds validated before and after resizing, connecting orbit proven, no conjugate points
    below L_-, L_+ condition, and the reference's answer: 2 conjugate points."""
    exe = os.path.join(ROOT, "tests", "cpp", "driver_dry")
    _build_with_oracle_backend(pkg, "driver_gpu.cpp", exe)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "manifolds 1/1" in r.stdout and "inside 1, proof 1" in r.stdout and "below L_- 1" in r.stdout
    assert "computeFrontAndConjugatePoints = 2" in r.stdout
    # sign variants of the coupling and the n = 2 defaults (heteroclinic.cpp:507-511): counts 1, 1 and 1
    for args, count in ((["1", "3", "1", ".98", ".96", "-.04", ".02"], 1), (["1", "3", "1", ".98", ".96", ".04", "-.02"], 1),
                        (["1", "2", "1", ".97", ".05"], 1)):
        r = subprocess.run([exe] + args, capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        assert "proof 1" in r.stdout and "computeFrontAndConjugatePoints = %d" % count in r.stdout


@pytest.mark.gpu
def test_compute_front_and_conjugate_points_on_gpu(pkg):
    exe = os.path.join(ROOT, "tests", "cpp", "driver_gpu")
    cmd = ["g++", "-std=c++17", "-O1", "-o", exe, os.path.join(ROOT, "tests", "cpp", "driver_gpu.cpp"),
           "-L", pkg.PKG_DIR, "-lccp_b200", "-lccp_setup", "-Wl,-rpath," + pkg.PKG_DIR]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "computeFrontAndConjugatePoints = 2" in r.stdout


def test_batch_driver_against_oracle_backend(pkg, orc):
    """computeFrontAndConjugatePointsBatch (host/ccp_driver.hpp): the complete proof for a sweep with all fronts in lock-step launches
    (SampleParameters, heteroclinic.cpp:429-491), against the oracle twins of the kernels: 6 fronts of Construct_ParameterList(6, 1);
    the batch must reproduce the single-front driver's code on the probed fronts (codes 2 -2 1 0 -2 1: the two fronts with c12 ~ 0 are
    degenerate for the boundary-value stage and end with -2, as in the reference's own sweep where such fronts are plotted as -1; the
    fourth front, (-c12, -c23), ended with -3 before the frame basis was renormalised once per unit of time (oracle/c2.hpp) and is now
    proven to have no conjugate point with the rigorous Krawczyk neighbourhood)."""
    exe = os.path.join(ROOT, "tests", "cpp", "driver_sweep_dry")
    _build_with_oracle_backend(pkg, "driver_sweep_gpu.cpp", exe)
    r = subprocess.run([exe, "6", "1"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "batch == single-front driver on the probed fronts" in r.stdout
    assert "codes: 2 -2 1 0 -2 1" in r.stdout
    # 36 fronts: enough for the per-front host stages to run on several host threads (parallel_for needs >= 32 items); the batch must still
    # reproduce the single-front driver on the probed fronts
    r = subprocess.run([exe, "12", "3"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "batch == single-front driver on the probed fronts" in r.stdout and "36 fronts, 36 connecting orbits proven" in r.stdout
    # the two-component system: a front with one conjugate point, a stable front (0, proven end to end), another with one
    r = subprocess.run([exe, "n2"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "n = 2 codes: 1 0 1" in r.stdout


@pytest.mark.gpu
def test_batch_driver_reference_sweep_on_gpu(pkg):
    exe = os.path.join(ROOT, "tests", "cpp", "driver_sweep_gpu")
    cmd = ["g++", "-std=c++17", "-O1", "-o", exe, os.path.join(ROOT, "tests", "cpp", "driver_sweep_gpu.cpp"),
           "-L", pkg.PKG_DIR, "-lccp_b200", "-lccp_setup", "-Wl,-rpath," + pkg.PKG_DIR]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe, "24", "4"], capture_output=True, text=True)          # the reference's own 96-point sweep
    assert r.returncode == 0, r.stdout + r.stderr
    assert "batch == single-front driver on the probed fronts" in r.stdout


def test_krawczyk_step_with_shared_integrations_against_oracle_backend(pkg, orc):
    """host/ccp_bvp.hpp: in the Krawczyk phase the point image G(x) comes from the integration over the box (image of the centre +
    flow_P * U_pt, mean-value form) -- two validated integrations per front instead of the reference's four
    (boundaryValueProblem.cpp:237-317).  Against the oracle twins: on the same inputs, step by step, the Krawczyk images of both
    arrangements intersect, the shared one is at most marginally wider, the verdicts agree."""
    exe = os.path.join(ROOT, "tests", "cpp", "bvp_share_dry")
    _build_with_oracle_backend(pkg, "bvp_share.cpp", exe)
    r = subprocess.run([exe, "3", "2"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "0 disagreements" in r.stdout


@pytest.mark.gpu
def test_krawczyk_step_with_shared_integrations_on_gpu(pkg):
    exe = os.path.join(ROOT, "tests", "cpp", "bvp_share")
    cmd = ["g++", "-std=c++17", "-O1", "-o", exe, os.path.join(ROOT, "tests", "cpp", "bvp_share.cpp"),
           "-L", pkg.PKG_DIR, "-lccp_b200", "-lccp_setup", "-Wl,-rpath," + pkg.PKG_DIR]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe, "24", "4"], capture_output=True, text=True)          # the reference's own 96-point sweep
    assert r.returncode == 0, r.stdout + r.stderr
    assert "0 disagreements" in r.stdout and "verified 96 (shared) / 96 (four integrations)" in r.stdout
