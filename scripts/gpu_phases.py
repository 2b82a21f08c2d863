"""Per-phase cycle breakdown of the front kernel (profiling build, block 0 only)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as ge
pkg = ge.load_package(); abi = pkg.abi
import ccp_b200
lib = abi.bind(C.CDLL(os.path.join(pkg.PKG_DIR, "libccp_b200_prof.so")))
ccp_b200._lib = lib
eng = pkg.Engine(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
par = {2: [1, .97, .05], 3: [1, .98, .96, .04, .02], 4: [1, .98, .96, .94, .04, .02, .03]}[n]
nf = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
desc, _ = pkg.setup_front(n, par)
descs = (abi.FrontDesc * nf)(*[desc] * nf)
eng.frame_count_batch(descs)
lib.ccp_debug_profile(None, 1)
out = eng.frame_count_batch(descs)
prof = (C.c_ulonglong * 16)()
lib.ccp_debug_profile(prof, 0)
names = ["hull+order0", "tables", "y,J eval", "Ws+Fk tables", "grid Horner+point dets", "range dets+decide+frames", "set update"]
tot = sum(prof[:7])
print("n=%d fronts=%d kernel %.2f ms steps %d" % (n, nf, eng.last_kernel_ms(), out[0].steps))
for i, nm in enumerate(names):
    print("  %-28s %10.0f cycles/step  %5.1f%%" % (nm, prof[i] / out[0].steps, 100.0 * prof[i] / tot))
print("  total %.0f cycles/step" % (tot / out[0].steps))
for i, nm in ((7, "  phi warp: A tasks"), (8, "  phi warp: B tasks"), (9, "  phi warp: combine"), (10, "  phi warp: wait for Psi warp"), (12, "  (loop overhead)")):
    print("  %-32s %10.0f cycles/step" % (nm, prof[i] / out[0].steps))
