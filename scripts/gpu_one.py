"""One launch of the front kernel on identical fronts (for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
pkg = ge.load_package(); abi = pkg.abi
if os.environ.get("CCP_LIB"):                      # experiment builds (e.g. libccp_b200_x.so)
    import ctypes as C, ccp_b200
    ccp_b200._lib = abi.bind(C.CDLL(os.path.join(pkg.PKG_DIR, os.environ["CCP_LIB"])))
eng = pkg.Engine(0)
nf = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
L = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
opts = abi.SetupOpts.default(L_minus_override=L) if L > 0 else None
n = int(sys.argv[3]) if len(sys.argv) > 3 else 3
par = {2: [1, .97, .05], 3: [1, .98, .96, .04, .02], 4: [1, .98, .96, .94, .04, .02, .03]}[n]
desc, _ = pkg.setup_front(n, par, opts)
descs = (abi.FrontDesc * nf)(*[desc] * nf)
reps = int(os.environ.get("CCP_REPS", "1"))
ms = []
for _ in range(reps):
    out = eng.frame_count_batch(descs); ms.append(eng.last_kernel_ms())
best = min(ms)
print("n", n, "fronts", nf, "kernel ms %.3f" % best, "(of %d)" % reps, "steps", out[0].steps, "zeros", out[0].zero_count, "fronts/s %.0f" % (nf / best * 1e3), "model TFLOP/s %.2f" % (nf * pkg.front_flops(n, 20, 14, out[0].steps) / best / 1e9))
