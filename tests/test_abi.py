"""The C-ABI library loads without a GPU and exports every symbol include/ccp.h declares."""
import ctypes as C
import os
import re


def test_library_exports_every_declared_symbol(pkg):
    header = open(os.path.join(os.path.dirname(pkg.PKG_DIR), "include", "ccp.h")).read()
    declared = set(re.findall(r"\b(ccp_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found"
    lib = C.CDLL(pkg.LIB_PATH)
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, missing
    assert declared == set(pkg.abi.SYMBOLS), declared ^ set(pkg.abi.SYMBOLS)
    # the host-only descriptor generator has its own header and library (CPU-only consumers never map the CUDA library)
    header = open(os.path.join(os.path.dirname(pkg.PKG_DIR), "include", "ccp_setup.h")).read()
    declared = set(re.findall(r"\b(ccp_[a-z0-9_]+)\s*\(", header))
    lib = C.CDLL(pkg.SETUP_LIB_PATH)
    assert declared == set(pkg.abi.SETUP_SYMBOLS) and all(hasattr(lib, s) for s in declared)
    assert not any(hasattr(C.CDLL(pkg.LIB_PATH), s) for s in declared), "the generator must not live in the CUDA library"


def test_struct_sizes_match_header(pkg):
    abi = pkg.abi
    assert C.sizeof(abi.Interval) == 16
    assert C.sizeof(abi.FrontDesc) == 1168 and C.sizeof(abi.FrontDesc) % 16 == 0
    assert C.sizeof(abi.FrontResult) == 2656
    assert C.sizeof(abi.SeriesRec) == 80


def test_no_cpu_fallback(pkg):
    """Without a device every compute entry point must fail loudly (never route through the oracle)."""
    L = pkg.lib()
    if L.ccp_device_count() > 0:
        return
    assert L.ccp_init(0) == pkg.abi.CCP_E_NO_DEVICE
    d = (pkg.abi.FrontDesc * 1)()
    r = (pkg.abi.FrontResult * 1)()
    assert L.ccp_frame_count_batch(d, 1, r) == pkg.abi.CCP_E_NOT_INIT
    try:
        pkg.Engine(0)
    except pkg.CcpError:
        pass
    else:
        raise AssertionError("Engine() must raise without a GPU")


def test_product_sources_do_not_touch_the_oracle(pkg):
    root = os.path.dirname(pkg.PKG_DIR)
    for dirpath, _, files in os.walk(pkg.PKG_DIR):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "liboracle" not in txt and "oracle/oracle.py" not in txt and '#include "../../oracle' not in txt, f
    assert os.path.isdir(os.path.join(root, "oracle"))
