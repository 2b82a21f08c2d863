import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as ge
pkg = ge.load_package(); orc = ge.load_oracle(); abi = pkg.abi
eng = pkg.Engine(0)
L = float(sys.argv[1]) if len(sys.argv) > 1 else 2.0
opts = abi.SetupOpts.default(L_minus_override=L)
desc, info = pkg.setup_front(3, [1.0, .98, .96, .04, .02], opts)
res, ser = eng.frame_count_series(desc)
c1, cser = orc.run_front(abi, desc, orc.MODE_C1, series=True)
for nm in ("status","zero_count","failure_count","series_length","steps","n_conj","first_failure"):
    print(nm, getattr(res,nm), getattr(c1,nm))
print("series shapes", ser.shape, cser.shape)
if ser.shape == cser.shape:
    bad = np.argwhere(ser != cser)
    print("series ndiff", len(bad), "first", bad[:8].tolist())
    if len(bad):
        i = bad[0][0]; print(ser[i]); print(cser[i]); print((ser[i]-cser[i]))
a = np.frombuffer(bytes(res), dtype=np.float64); b = np.frombuffer(bytes(c1), dtype=np.float64)
d = np.argwhere(~((a == b) | (np.isnan(a) & np.isnan(b)))).ravel()
print("result diff idx", d[:40], "n", len(d))
for i in d[:10]: print(i, a[i], b[i], a[i]-b[i])
