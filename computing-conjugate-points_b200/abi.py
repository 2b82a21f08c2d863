"""ctypes mirror of include/ccp.h (the C ABI of libccp_b200.so).

Field order and sizes must match the header exactly; tests/test_abi.py checks sizeof() against the
values the library reports and that every declared symbol is exported.
"""
import ctypes as C

CCP_MAX_N = 4
CCP_MAX_DIM = 8
CCP_MAX_ORDER = 20
CCP_MAX_GRID = 15
CCP_MAX_CONJ = 8

CCP_OK, CCP_ST_BAD_DESC, CCP_ST_ENCLOSURE, CCP_ST_NONFINITE = 0, 1, 2, 3
CCP_E_CUDA, CCP_E_NO_DEVICE, CCP_E_ARG, CCP_E_NOT_INIT = -100, -101, -102, -103


class Interval(C.Structure):
    _fields_ = [("lo", C.c_double), ("hi", C.c_double)]

    def tup(self):
        return (self.lo, self.hi)


class FrontDesc(C.Structure):
    _fields_ = [
        ("n", C.c_int32), ("order", C.c_int32), ("step_log2", C.c_int32), ("grid", C.c_int32),
        ("L_minus", C.c_double), ("reserved", C.c_double),
        ("b", Interval * CCP_MAX_N),
        ("c", Interval * (CCP_MAX_N - 1)),
        ("A_u", C.c_double * (CCP_MAX_DIM * CCP_MAX_DIM)),
        ("p_i", Interval * CCP_MAX_DIM),
        ("Uxy", Interval * CCP_MAX_DIM),
        ("Eu_err", Interval * (CCP_MAX_N * CCP_MAX_N)),
    ]


class FrontResult(C.Structure):
    _fields_ = [
        ("status", C.c_int32), ("zero_count", C.c_int32), ("failure_count", C.c_int32),
        ("series_length", C.c_int32), ("steps", C.c_int32), ("n_conj", C.c_int32),
        ("first_failure", C.c_int32), ("reserved", C.c_int32),
        ("conj_time", Interval * CCP_MAX_CONJ),
        ("conj_index", C.c_int32 * CCP_MAX_CONJ),
        ("last_frame", Interval * (CCP_MAX_DIM * (CCP_MAX_N + 2))),
        ("first_frame", Interval * (CCP_MAX_DIM * (CCP_MAX_N + 1))),
        ("det_first", Interval),
        ("det_last", Interval),
        ("flow_P", Interval * (CCP_MAX_DIM * CCP_MAX_DIM)),
    ]


class SeriesRec(C.Structure):
    _fields_ = [("time", Interval), ("det_L", Interval), ("det_R", Interval), ("det_V", Interval), ("det_D", Interval)]


class SetupOpts(C.Structure):
    _fields_ = [
        ("xy_nbd", C.c_double), ("cone_L", C.c_double), ("eig_err", C.c_double), ("scale", C.c_double),
        ("L_minus_percent", C.c_double), ("L_minus_override", C.c_double),
        ("order", C.c_int32), ("step_log2", C.c_int32), ("grid", C.c_int32), ("newton_steps", C.c_int32),
    ]

    @classmethod
    def default(cls, **kw):
        o = cls(xy_nbd=-1.0, cone_L=-1.0, eig_err=-1.0, scale=-1.0, L_minus_percent=-1.0, L_minus_override=-1.0,
                order=0, step_log2=0, grid=0, newton_steps=0)
        for k, v in kw.items():
            setattr(o, k, v)
        return o


class SetupInfo(C.Structure):
    _fields_ = [
        ("T", C.c_double), ("L_minus", C.c_double), ("newton_residual", C.c_double),
        ("eig_unstable", C.c_double * CCP_MAX_N), ("x_local", C.c_double * CCP_MAX_N), ("y_local", C.c_double * CCP_MAX_N),
        ("converged", C.c_int32), ("reserved", C.c_int32),
    ]


CCP_DFU_MAX_SUBDIV = 32


class DfuDesc(C.Structure):
    """One call of localManifold::boundDFU (include/ccp.h: ccp_dfu_desc)."""
    _fields_ = [
        ("n", C.c_int32), ("subdivision", C.c_int32), ("stable", C.c_int32), ("reserved", C.c_int32),
        ("b", Interval * CCP_MAX_N), ("c", Interval * CCP_MAX_N),
        ("A", C.c_double * (CCP_MAX_DIM * CCP_MAX_DIM)), ("Ainv", Interval * (CCP_MAX_DIM * CCP_MAX_DIM)),
        ("p", Interval * CCP_MAX_DIM), ("U", Interval * CCP_MAX_DIM),
    ]


class DfuResult(C.Structure):
    _fields_ = [("status", C.c_int32), ("parts", C.c_int32), ("DF", Interval * (CCP_MAX_DIM * CCP_MAX_DIM))]


CCP_SUBDIV_UXY, CCP_SUBDIV_EU_ERR = 0, 1


class SubdivResult(C.Structure):
    """Hull / tally of the sub-boxes of one parameter (include/ccp.h: ccp_subdiv_result)."""
    _fields_ = [
        ("status", C.c_int32), ("sub_boxes", C.c_int32), ("certified", C.c_int32), ("zero_count", C.c_int32),
        ("zero_min", C.c_int32), ("zero_max", C.c_int32), ("failures", C.c_int32), ("steps", C.c_int32),
        ("last_frame", Interval * (CCP_MAX_DIM * (CCP_MAX_N + 2))),
        ("first_frame", Interval * (CCP_MAX_DIM * (CCP_MAX_N + 1))),
        ("det_first", Interval), ("det_last", Interval),
        ("flow_P", Interval * (CCP_MAX_DIM * CCP_MAX_DIM)),
    ]


# every symbol include/ccp.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "ccp_device_count": (C.c_int, []),
    "ccp_init": (C.c_int, [C.c_int]),
    "ccp_init_multi": (C.c_int, [C.c_int]),
    "ccp_bound_devices": (C.c_int, []),
    "ccp_shutdown": (None, []),
    "ccp_last_error": (C.c_char_p, []),
    "ccp_version": (C.c_char_p, []),
    "ccp_frame_count_batch": (C.c_int, [C.POINTER(FrontDesc), C.c_size_t, C.POINTER(FrontResult)]),
    "ccp_frame_count_batch_multi": (C.c_int, [C.POINTER(FrontDesc), C.c_size_t, C.POINTER(FrontResult)]),
    "ccp_frame_count_batch_device": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]),
    "ccp_frame_count_batch_device_n": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_int]),
    "ccp_frame_count_subdivided": (C.c_int, [C.POINTER(FrontDesc), C.c_size_t, C.c_int, C.c_int, C.POINTER(SubdivResult)]),
    "ccp_frame_count_series": (C.c_int, [C.POINTER(FrontDesc), C.POINTER(FrontResult), C.POINTER(SeriesRec), C.c_size_t, C.POINTER(C.c_size_t)]),
    "ccp_bound_dfu_batch": (C.c_int, [C.POINTER(DfuDesc), C.c_size_t, C.POINTER(DfuResult)]),
    "ccp_kernel_launches": (C.c_uint64, []),
    "ccp_last_kernel_ms": (C.c_float, []),
    "ccp_debug_primitives": (C.c_int, [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_size_t]),
    "ccp_measure_fp64_peak": (C.c_int, [C.c_int, C.POINTER(C.c_double)]),
    "ccp_front_flops": (C.c_double, [C.c_int, C.c_int, C.c_int, C.c_int]),
}

# every symbol include/ccp_setup.h declares (libccp_setup.so: host-only descriptor generator)
SETUP_SYMBOLS = {
    "ccp_setup_front": (C.c_int, [C.c_int, C.POINTER(C.c_double), C.POINTER(SetupOpts), C.POINTER(FrontDesc), C.POINTER(SetupInfo)]),
    "ccp_setup_sweep": (C.c_int, [C.c_int, C.POINTER(C.c_double), C.c_size_t, C.POINTER(SetupOpts), C.POINTER(FrontDesc), C.POINTER(SetupInfo), C.c_int]),
    "ccp_parameter_list": (C.c_int, [C.c_int, C.c_int, C.c_double, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
}


def bind(lib, symbols=None):
    for name, (res, args) in (SYMBOLS if symbols is None else symbols).items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib
