"""Cycle timeline of one integrator step (step 500 of block 0), both warps (trace build)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as ge
pkg = ge.load_package(); abi = pkg.abi
import ccp_b200
lib = abi.bind(C.CDLL(os.path.join(pkg.PKG_DIR, "libccp_b200_trace.so")))
ccp_b200._lib = lib
eng = pkg.Engine(0)
nf = int(sys.argv[1]) if len(sys.argv) > 1 else 148
desc, _ = pkg.setup_front(3, [1, .98, .96, .04, .02])
descs = (abi.FrontDesc * nf)(*[desc] * nf)
eng.frame_count_batch(descs)
out = eng.frame_count_batch(descs)
buf = (C.c_longlong * 1024)(); n = (C.c_int * 2)()
lib.ccp_debug_trace(buf, n)
a = np.array(buf).reshape(2, 256, 2)
print("fronts", nf, "kernel ms", eng.last_kernel_ms())
t0 = a[0, 0, 0]
names = {1: "step start", 2: "after hull+order0", 3: "after table loop", 4: "after y,J,Ws", 5: "after Fk", 6: "after eval", 7: "after dets+decide", 8: "after set update", 9: "after boxF", 10: "after X1 (d2,d3)", 11: "after X2 (N_k)", 12: "after X3 (Xi)", 13: "after X4 (dJ)"}
for w in range(2):
    print("warp", w)
    prev = None
    for i in range(n[w]):
        t, m = int(a[w, i, 0] - t0), int(a[w, i, 1])
        if m >= 100:
            kind, k = m // 100, m % 100
            if k not in (0, 1, 4, 5, 10, 11, 14, 15, 18, 19, 20): prev = t; continue
            nm = {1: "phi pass start", 2: "phi pass done", 3: "phi sums stored (combine start)", 4: "combine done (at bar)", 5: "after bar", 6: "psi start", 7: "psi pass start", 8: "psi sums stored (combine start)"}[kind] + " k=%d" % k
        else:
            nm = names[m]
        print("  %8d  (+%6d)  %s" % (t, t - (prev if prev is not None else t), nm))
        prev = t
