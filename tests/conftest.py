import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_L_MINUS = 33.84706312634513          # last t_mid + t_rad of the reference's plot_det.txt
CONFIG1 = [1.0, 0.98, 0.96, 0.04, 0.02]      # heteroclinic.cpp:524-529


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def ge():
    import __graft_entry__ as g
    return g


@pytest.fixture(scope="session")
def pkg(ge):
    p = ge.load_package()
    if not (os.path.exists(p.LIB_PATH) and os.path.exists(p.SETUP_LIB_PATH)):
        p.build()
    return p


@pytest.fixture(scope="session")
def orc(ge, pkg):
    o = ge.load_oracle()
    o.lib(pkg.abi)
    return o


@pytest.fixture(scope="session")
def engine(pkg):
    return pkg.Engine(0)


THIN_BOX = 1e-19        # initial box of the connecting orbit at which the reference-structured oracle certifies (see thin_opts)


@pytest.fixture(scope="session")
def thin_opts(pkg):
    """Descriptor options with a (non-rigorous) thin neighbourhood of the connecting orbit, xy_nbd = 1e-19.  The generator's
    default is the rigorous size 1e-12; the reference-structured oracle (mode ref, like the reference itself: no
    renormalisation of the frame basis) cannot certify the last third of a front with it, so every comparison of COUNTS with
    mode ref is made on the thin box, where it certifies."""
    def make(**kw):
        return pkg.abi.SetupOpts.default(xy_nbd=THIN_BOX, **kw)
    return make


@pytest.fixture(scope="session")
def config1_desc(pkg):
    """n=3 reference defaults, L_- pinned to the value recorded in plot_det.txt, decimal-string parameters, rigorous box 1e-12."""
    opts = pkg.abi.SetupOpts.default(L_minus_override=GOLDEN_L_MINUS)
    desc, info = pkg.setup_front(3, CONFIG1, opts)
    pkg.decimal_params(desc)
    return desc, info


@pytest.fixture(scope="session")
def config1_desc_thin(pkg, thin_opts):
    desc, info = pkg.setup_front(3, CONFIG1, thin_opts(L_minus_override=GOLDEN_L_MINUS))
    pkg.decimal_params(desc)
    return desc, info


@pytest.fixture(scope="session")
def config1_ref(pkg, orc, config1_desc_thin):
    """One full reference-structured oracle run of config 1 (a few seconds, shared by the session), thin box."""
    return orc.run_front(pkg.abi, config1_desc_thin[0], orc.MODE_REF, series=True, threads=3)


@pytest.fixture(scope="session")
def config1_c2(pkg, orc, config1_desc):
    """The kernel's twin on config 1 with the rigorous box."""
    return orc.run_front(pkg.abi, config1_desc[0], orc.MODE_C2, series=True)


@pytest.fixture(scope="session")
def config1_c2_thin(pkg, orc, config1_desc_thin):
    return orc.run_front(pkg.abi, config1_desc_thin[0], orc.MODE_C2, series=True)
