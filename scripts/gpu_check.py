"""GPU vs oracle (kernel twin, mode c2) bit comparison with first-mismatch diagnostics: python scripts/gpu_check.py [n] [L_minus]"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
pkg = ge.load_package(); orc = ge.load_oracle(); abi = pkg.abi; orc.lib(abi)
eng = pkg.Engine(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
L = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
par = {2: [1, .97, .05], 3: [1, .98, .96, .04, .02], 4: [1, .98, .96, .94, .04, .02, .03]}[n]
opts = abi.SetupOpts.default(L_minus_override=L) if L > 0 else None
desc, _ = pkg.setup_front(n, par, opts)
res, ser = eng.frame_count_series(desc)
cr, cs = orc.run_front(abi, desc, orc.MODE_C2, series=True)
ser = np.asarray(ser)
print("gpu status %d zeros %d fail %d steps %d | c2 status %d zeros %d fail %d" % (res.status, res.zero_count, res.failure_count, res.steps, cr.status, cr.zero_count, cr.failure_count))
print("series shapes", ser.shape, cs.shape)
m = min(len(ser), len(cs))
neq = (ser[:m].view(np.uint64) != cs[:m].view(np.uint64))
if neq.any():
    rows = np.where(neq.any(axis=1))[0]
    print("series mismatch rows: %d, first row %d (step %d, sub %d) cols %s" % (len(rows), rows[0], rows[0] // desc.grid, rows[0] % desc.grid, np.where(neq[rows[0]])[0]))
    r = rows[0]
    print(" gpu", ser[r]); print(" c2 ", cs[r]); print(" rel diff", np.abs(ser[r] - cs[r]) / np.maximum(np.abs(cs[r]), 1e-300))
else:
    print("series bit-identical (%d rows)" % m)
a = np.frombuffer(bytes(res), dtype=np.uint64); b = np.frombuffer(bytes(cr), dtype=np.uint64)
if np.array_equal(a, b):
    print("result record bit-identical")
else:
    idx = np.where(a != b)[0]
    print("result record differs at 8-byte words", idx[:40], "of", len(a))
    fa = np.frombuffer(bytes(res), dtype=np.float64); fb = np.frombuffer(bytes(cr), dtype=np.float64)
    for i in idx[:8]: print("  word", i, fa[i], fb[i])
