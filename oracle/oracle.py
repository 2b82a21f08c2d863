"""oracle/oracle.py -- TEST INFRASTRUCTURE: ctypes wrapper of the CPU oracle (oracle/_build/liboracle.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
It is the checker, never the product.  Struct layouts come from the product's ABI mirror (abi.py) because
the oracle consumes/produces the same plain-old-data records as the C ABI.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "liboracle.so")
MODE_REF, MODE_C1, MODE_C2 = 0, 1, 2      # c2 = twin of the CUDA kernel
_lib = None


def build():
    r = subprocess.run(["make", "-C", HERE], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + r.stdout[-3000:] + r.stderr[-3000:])
    return LIB


def lib(abi):
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        L = C.CDLL(LIB)
        L.orc_run_front.restype = C.c_int
        L.orc_run_front.argtypes = [C.POINTER(abi.FrontDesc), C.c_int, C.POINTER(abi.FrontResult), C.POINTER(abi.SeriesRec),
                                    C.c_size_t, C.POINTER(C.c_size_t), C.c_int]
        L.orc_run_batch.restype = C.c_int
        L.orc_run_batch.argtypes = [C.POINTER(abi.FrontDesc), C.c_size_t, C.c_int, C.POINTER(abi.FrontResult), C.c_int, C.POINTER(C.c_double)]
        L.orc_ival_op.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.POINTER(C.c_double)]
        L.orc_mt_dot.argtypes = [C.POINTER(C.c_double)] * 4 + [C.c_int, C.POINTER(C.c_double)]
        L.orc_iv2mt.argtypes = [C.c_double, C.c_double, C.POINTER(C.c_double)]
        L.orc_mt2iv.argtypes = [C.c_double, C.c_double, C.POINTER(C.c_double)]
        L.orc_det.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_double)]
        L.orc_majorant.argtypes = [C.POINTER(abi.FrontDesc), C.POINTER(C.c_double)]
        _lib = L
    return _lib


def run_front(abi, desc, mode, series=False, threads=1):
    L = lib(abi)
    res = abi.FrontResult()
    ln = C.c_size_t(0)
    if not series:
        L.orc_run_front(C.byref(desc), mode, C.byref(res), None, 0, C.byref(ln), threads)
        return res, None
    cap = 1 << 16
    ser = (abi.SeriesRec * cap)()
    L.orc_run_front(C.byref(desc), mode, C.byref(res), ser, cap, C.byref(ln), threads)
    arr = np.frombuffer(ser, dtype=np.float64).reshape(cap, 10)[: ln.value].copy()
    return res, arr


def run_batch(abi, descs, mode, threads=0):
    L = lib(abi)
    m = len(descs)
    res = (abi.FrontResult * m)()
    sec = C.c_double(0)
    L.orc_run_batch(descs, m, mode, res, threads, C.byref(sec))
    return res, sec.value


def bound_dfu(abi, descs):
    """oracle/dfu.hpp: serial restatement of localManifold::boundDFU for an array of abi.DfuDesc."""
    L = lib(abi)
    L.orc_bound_dfu.argtypes = [C.POINTER(abi.DfuDesc), C.c_size_t, C.POINTER(abi.DfuResult)]
    m = len(descs)
    res = (abi.DfuResult * m)()
    L.orc_bound_dfu(descs, m, res)
    return res


def ival_op(abi, op, a, b):
    out = (C.c_double * 2)()
    lib(abi).orc_ival_op(op, a[0], a[1], b[0], b[1], out)
    return out[0], out[1]


def subdivided(abi, descs, parts, which, mode=MODE_C2, threads=0, keep_sub=False):
    """oracle/subdiv.hpp: twin of ccp_frame_count_subdivided (BASELINE config 4).  Returns (SubdivResult array, sub-box records or None)."""
    L = lib(abi)
    L.orc_subdivided.restype = C.c_int
    L.orc_subdivided.argtypes = [C.POINTER(abi.FrontDesc), C.c_size_t, C.c_int, C.c_int, C.c_int, C.POINTER(abi.SubdivResult), C.c_void_p, C.c_int]
    m = len(descs)
    res = (abi.SubdivResult * m)()
    sub = None
    if keep_sub:
        sub = (abi.FrontResult * (m * parts ** descs[0].n))()
    L.orc_subdivided(descs, m, parts, which, mode, res, C.cast(sub, C.c_void_p) if sub is not None else None, threads)
    return res, sub
