"""ccp_b200 -- host-side Python mirror of the reference's frameDet call for batches of fronts.

The product is libccp_b200.so (CUDA, sm_100a) behind the C ABI of include/ccp.h; this package is the
thin ctypes binding used by tests/, bench.py and __graft_entry__.py, plus the sweep helpers that
restate the reference's `Construct_ParameterList` / `SampleParameters` loop (heteroclinic.cpp:386-491).
There is no CPU fallback: `Engine()` raises if the library or a GPU is missing.

Import it through `__graft_entry__.load_package()` (the directory name contains '-').
"""
import ctypes as C
import os
import subprocess

import numpy as np

from . import abi
from .abi import FrontDesc, FrontResult, SeriesRec, SetupOpts, SetupInfo, Interval  # noqa: F401

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG_DIR, "libccp_b200.so")
SETUP_LIB_PATH = os.path.join(PKG_DIR, "libccp_setup.so")
_lib = None
_setup_lib = None


class CcpError(RuntimeError):
    pass


def build(verbose=False):
    """Compile libccp_b200.so for sm_100a in-tree (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", os.path.join(PKG_DIR, "csrc")], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout[-4000:])
        print(r.stderr[-4000:])
    if r.returncode != 0:
        raise CcpError("nvcc build of libccp_b200.so failed")
    return LIB_PATH


def lib():
    """The bound shared library; raises loudly if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise CcpError("libccp_b200.so not built: run __graft_entry__.build() (there is no CPU fallback)")
        _lib = abi.bind(C.CDLL(LIB_PATH))
    return _lib


def setup_lib():
    """libccp_setup.so: the host-only descriptor generator (include/ccp_setup.h).  CPU-only consumers (the reference arm of
    bench.py, the oracle tests) use this without ever mapping the CUDA library."""
    global _setup_lib
    if _setup_lib is None:
        if not os.path.exists(SETUP_LIB_PATH):
            raise CcpError("libccp_setup.so not built: run __graft_entry__.build()")
        _setup_lib = abi.bind(C.CDLL(SETUP_LIB_PATH), abi.SETUP_SYMBOLS)
    return _setup_lib


def _check(rc, what):
    if rc != 0:
        detail = lib().ccp_last_error().decode() if _lib is not None else ""
        raise CcpError("%s failed with code %d: %s" % (what, rc, detail))


# ----------------------------------------------------------------------------- descriptors (host only)
def parameter_list(N_c=24, N_r=4, r_min=0.02, r_max=0.06, b=(1.0, 0.975, 0.95)):
    """Construct_ParameterList (heteroclinic.cpp:386-427): N_c*N_r vectors (b1,b2,b3,c12,c23)."""
    out = np.zeros((N_c * N_r, 5), dtype=np.float64)
    b3 = (C.c_double * 3)(*b)
    _check(setup_lib().ccp_parameter_list(N_c, N_r, r_min, r_max, b3, out.ctypes.data_as(C.POINTER(C.c_double))), "ccp_parameter_list")
    return out


def setup_sweep(n, params, opts=None, threads=0):
    """Float shooting for every parameter vector -> (FrontDesc array, SetupInfo array)."""
    params = np.ascontiguousarray(params, dtype=np.float64).reshape(-1, 2 * n - 1)
    m = params.shape[0]
    descs = (FrontDesc * m)()
    infos = (SetupInfo * m)()
    o = opts if opts is not None else SetupOpts.default()
    _check(setup_lib().ccp_setup_sweep(n, params.ctypes.data_as(C.POINTER(C.c_double)), m, C.byref(o), descs, infos, threads), "ccp_setup_sweep")
    return descs, infos


def setup_front(n, params, opts=None):
    d, i = setup_sweep(n, np.asarray(params, dtype=np.float64).reshape(1, -1), opts, threads=1)
    return d[0], i[0]


def decimal_params(desc):
    """Widen point parameters to the 2-ulp hull that contains the decimal literal, mimicking the
    reference's interval("x","x") string constructor used for the defaults in main (heteroclinic.cpp:513-531)."""
    for arr, cnt in ((desc.b, desc.n), (desc.c, desc.n - 1)):
        for i in range(cnt):
            v = arr[i].lo
            arr[i].lo = np.nextafter(v, -np.inf)
            arr[i].hi = np.nextafter(v, np.inf)
    return desc


def subdivide_front(desc, parts):
    """Config 4: split the eigenfunction-error box of one front into parts**n sub-boxes (the reference's
    `part()` rule, utils.cpp:3-10: piece i of [lo,hi] = lo + (hi-lo)*[i,i+1]/parts), each its own descriptor."""
    n = desc.n
    total = parts ** n
    out = (FrontDesc * total)()
    for idx in range(total):
        C.memmove(C.byref(out[idx]), C.byref(desc), C.sizeof(FrontDesc))
        rem = idx
        for j in range(n):           # split row j of every column k the same way (n-dimensional subdivision)
            pj = rem % parts
            rem //= parts
            for k in range(n):
                e = desc.Eu_err[j * abi.CCP_MAX_N + k]
                w = e.hi - e.lo
                lo = e.lo + w * pj / parts
                hi = e.lo + w * (pj + 1) / parts
                out[idx].Eu_err[j * abi.CCP_MAX_N + k].lo = np.nextafter(lo, -np.inf)
                out[idx].Eu_err[j * abi.CCP_MAX_N + k].hi = np.nextafter(hi, np.inf)
    return out


def shard(n_items, rank, world):
    """Interleaved static shard of the sweep (front_id mod world), SURVEY.md 8(e)."""
    return list(range(rank, n_items, world))


def frame_det_code(res):
    """Return value of propagateManifold::frameDet for one result record (propagateManifold.cpp:178-186):
    count >= 0, or -3 when countZeros reports failures; engine-level failures map to -10 like a thrown exception
    (heteroclinic.cpp:460-467)."""
    if res.status != abi.CCP_OK:
        return -10
    return -3 if res.failure_count > 0 else res.zero_count


# ----------------------------------------------------------------------------- boundary-value stage: C^1 flow maps
def flow_desc(n, params, p_i, A_i, Uxy, T, stable=False, order=20, step_log2=5):
    """Descriptor of one validated integration of the boundary-value stage: boundaryValueProblem::Gxy / DGxy
    (boundaryValueProblem.cpp:9-115) integrate C0Rect2Set / C1Rect2Set(p_i, A_i, Uxy) for time T, forwards from the unstable
    manifold or, `stable=True`, with the reversed field f_minus from the stable manifold.  The reversed flow is the forward
    flow of the mirrored set (S = diag(I,-I), host/ccp_bvp.hpp), so one kernel serves both.  `p_i`, `Uxy`: (2n,2) arrays of
    interval end points, `A_i`: (2n,2n) point matrix, `params`: b_1..b_n, c_12.. as numbers or (lo,hi) pairs."""
    d = abi.FrontDesc()
    D = 2 * n
    d.n, d.order, d.step_log2, d.grid, d.L_minus = n, order, step_log2, 0, float(T)      # grid 0 = flow map only
    pr = [(float(q), float(q)) if np.isscalar(q) else (float(q[0]), float(q[1])) for q in params]
    for i in range(n):
        d.b[i].lo, d.b[i].hi = pr[i]
    for i in range(n - 1):
        d.c[i].lo, d.c[i].hi = pr[n + i]
    p_i = np.asarray(p_i, dtype=np.float64).reshape(D, 2); Uxy = np.asarray(Uxy, dtype=np.float64).reshape(D, 2)
    A_i = np.asarray(A_i, dtype=np.float64).reshape(D, D)
    for i in range(D):
        flip = stable and i >= n
        d.p_i[i].lo, d.p_i[i].hi = (-p_i[i, 1], -p_i[i, 0]) if flip else (p_i[i, 0], p_i[i, 1])
        d.Uxy[i].lo, d.Uxy[i].hi = Uxy[i, 0], Uxy[i, 1]
        for j in range(D):
            d.A_u[i * abi.CCP_MAX_DIM + j] = -A_i[i, j] if flip else A_i[i, j]
    return d


def flow_result(res, n, stable=False):
    """(image, deriv_A) of one flow-map result record: image[2n,2] encloses Phi_T(set) (Gxy), deriv_A[2n,2n,2] encloses
    [D Phi_T] A_i over the set (DGxy before the factor DW); mirrored back for `stable=True`."""
    D = 2 * n
    img = np.zeros((D, 2)); P = np.zeros((D, D, 2))
    for i in range(D):
        flip = stable and i >= n
        a = res.last_frame[i * (abi.CCP_MAX_N + 2) + n + 1]
        img[i] = (-a.hi, -a.lo) if flip else (a.lo, a.hi)
        for j in range(D):
            q = res.flow_P[i * abi.CCP_MAX_DIM + j]
            P[i, j] = (-q.hi, -q.lo) if flip else (q.lo, q.hi)
    return img, P


def dfu_desc(n, params, A, p, U, subdivision=15, stable=False, ainv_rel_radius=1e-13):
    """Descriptor of one call of localManifold::boundDFU.  `A`: (2n,2n) point matrix ((*pF).A), `p`: equilibrium (2n numbers),
    `U`: (2n,2) box in local coordinates.  The reference hands over CAPD's gaussInverseMatrix(A); for tests and synthetic sweeps the
    float inverse is inflated by `ainv_rel_radius` * max|A^-1| per entry (far above its actual error for these 2n x 2n matrices)."""
    D = 2 * n
    d = abi.DfuDesc()
    d.n, d.subdivision, d.stable = n, int(subdivision), 1 if stable else 0
    pr = [(float(q), float(q)) if np.isscalar(q) else (float(q[0]), float(q[1])) for q in params]
    for i in range(n):
        d.b[i].lo, d.b[i].hi = pr[i]
    for i in range(n - 1):
        d.c[i].lo, d.c[i].hi = pr[n + i]
    A = np.asarray(A, dtype=np.float64).reshape(D, D)
    Ai = np.linalg.inv(A); rad = ainv_rel_radius * np.abs(Ai).max()
    U = np.asarray(U, dtype=np.float64).reshape(D, 2)
    for i in range(D):
        d.p[i].lo = d.p[i].hi = float(p[i])
        d.U[i].lo, d.U[i].hi = U[i, 0], U[i, 1]
        for j in range(D):
            d.A[i * abi.CCP_MAX_DIM + j] = A[i, j]
            d.Ainv[i * abi.CCP_MAX_DIM + j].lo, d.Ainv[i * abi.CCP_MAX_DIM + j].hi = Ai[i, j] - rad, Ai[i, j] + rad
    return d


# ----------------------------------------------------------------------------- the engine (GPU only)
class Engine:
    """One process, one GPU.  `frame_count_batch` is the drop-in for the reference's serial loop over
    `computeFrontAndConjugatePoints -> E_u.frameDet(...)` (heteroclinic.cpp:334,449-471)."""

    def __init__(self, device=0, n_devices=None):
        """`device`: the GPU this process is bound to (one process per GPU, as under torchrun).  `n_devices` (instead): ONE process
        drives GPUs 0..n_devices-1 (0 = all visible) through ccp_init_multi; `frame_count_batch_multi` then shards a batch over them."""
        L = lib()
        if L.ccp_device_count() <= 0:
            raise CcpError("no CUDA device visible: ccp_b200 has no CPU fallback")
        if n_devices is None:
            _check(L.ccp_init(device), "ccp_init")
        else:
            _check(L.ccp_init_multi(n_devices), "ccp_init_multi")
        self.L = L

    def bound_devices(self):
        return int(self.L.ccp_bound_devices())

    def frame_count_batch(self, descs):
        m = len(descs)
        res = (FrontResult * m)()
        _check(self.L.ccp_frame_count_batch(descs, m, res), "ccp_frame_count_batch")
        return res

    def frame_count_batch_multi(self, descs):
        """The batch over every bound GPU (front f on device f mod n_devices); records in front order."""
        m = len(descs)
        res = (FrontResult * m)()
        _check(self.L.ccp_frame_count_batch_multi(descs, m, res), "ccp_frame_count_batch_multi")
        return res

    def frame_count_subdivided(self, descs, parts, which=abi.CCP_SUBDIV_UXY):
        """BASELINE config 4: every front's box cut into parts**n sub-boxes (the reference's part() rule), each propagated and counted
        independently, hull / tally per parameter on the device.  Returns one abi.SubdivResult per front."""
        m = len(descs)
        res = (abi.SubdivResult * m)()
        _check(self.L.ccp_frame_count_subdivided(descs, m, parts, which, res), "ccp_frame_count_subdivided")
        return res

    def bound_dfu_batch(self, descs):
        """localManifold::boundDFU (localManifold.cpp:347-409) for a batch of calls: one abi.DfuDesc each."""
        m = len(descs)
        res = (abi.DfuResult * m)()
        _check(self.L.ccp_bound_dfu_batch(descs, m, res), "ccp_bound_dfu_batch")
        return res

    def frame_count_batch_device(self, d_descs_ptr, m, d_results_ptr, stream_ptr=None, n=None):
        """Device-resident batch.  With `n` (components, uniform over the batch) the call only enqueues; without it the library reads
        n from the first descriptor (a 4-byte copy and one stream synchronisation)."""
        st = C.c_void_p(stream_ptr) if stream_ptr else None
        if n is None:
            _check(self.L.ccp_frame_count_batch_device(C.c_void_p(d_descs_ptr), m, C.c_void_p(d_results_ptr), st), "ccp_frame_count_batch_device")
        else:
            _check(self.L.ccp_frame_count_batch_device_n(C.c_void_p(d_descs_ptr), m, C.c_void_p(d_results_ptr), st, n), "ccp_frame_count_batch_device_n")

    def frame_count_series(self, desc, cap=None):
        res = FrontResult()
        ln = C.c_size_t(0)
        if cap is None:
            _check(self.L.ccp_frame_count_series(C.byref(desc), C.byref(res), None, 0, C.byref(ln)), "ccp_frame_count_series")
            cap = ln.value
        ser = (SeriesRec * max(cap, 1))()
        _check(self.L.ccp_frame_count_series(C.byref(desc), C.byref(res), ser, cap, C.byref(ln)), "ccp_frame_count_series")
        arr = np.frombuffer(ser, dtype=np.float64).reshape(-1, 10)[: min(cap, ln.value)].copy()
        return res, arr

    def fp64_peak_tflops(self, iters=4096):
        t = C.c_double(0)
        _check(self.L.ccp_measure_fp64_peak(iters, C.byref(t)), "ccp_measure_fp64_peak")
        return t.value

    def kernel_launches(self):
        return int(self.L.ccp_kernel_launches())

    def last_kernel_ms(self):
        return float(self.L.ccp_last_kernel_ms())

    def close(self):
        self.L.ccp_shutdown()


def front_flops(n, order, grid, steps):
    return float(lib().ccp_front_flops(n, order, grid, steps))


def results_to_numpy(res):
    """View an array of FrontResult as raw float64 rows (for bit-exact comparisons)."""
    m = len(res)
    return np.frombuffer(res, dtype=np.uint8).reshape(m, C.sizeof(FrontResult)).copy()
