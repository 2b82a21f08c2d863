"""Throughput of localManifold::boundDFU on the GPU (calls/s, sub-boxes/s) with the serial oracle timed beside it."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as ge
pkg = ge.load_package(); orc = ge.load_oracle(); abi = pkg.abi; orc.lib(abi)
if os.environ.get("CCP_LIB"):                      # experiment builds (e.g. libccp_b200_x.so)
    import ctypes as C
    pkg._lib = abi.bind(C.CDLL(os.path.join(pkg.PKG_DIR, os.environ["CCP_LIB"])))
eng = pkg.Engine(0)
m = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
n, params, N = 3, [1.0, .98, .96, .04, .02], 15
desc, _ = pkg.setup_front(n, params)
D = 2 * n
A = np.array([[desc.A_u[i * 8 + j] for j in range(D)] for i in range(D)])
U = np.zeros((D, 2)); U[:n] = [-1.2e-6, 1.2e-6]; U[n:] = [-2e-11, 2e-11]
d = pkg.dfu_desc(n, params, A, np.zeros(D), U, subdivision=N, stable=False)
arr = (abi.DfuDesc * m)(*[d] * m)
eng.bound_dfu_batch(arr)
ms = []
for _ in range(3):
    r = eng.bound_dfu_batch(arr); ms.append(eng.last_kernel_ms())
best = min(ms)
t = time.time(); c = orc.bound_dfu(abi, (abi.DfuDesc * 8)(*[d] * 8)); cpu = (time.time() - t) / 8
same = bytes(r[0]) == bytes(c[0])
print("boundDFU n=%d N=%d (%d sub-boxes per call): %d calls in %.3f ms -> %.0f calls/s, %.3g sub-boxes/s; serial oracle %.1f ms per call (%.0f calls/s/core); GPU == oracle: %s"
      % (n, N, N ** n, m, best, m / best * 1e3, m * N ** n / best * 1e3, cpu * 1e3, 1 / cpu, same))
