"""Scratch GPU probe: smoke, config-1 parity vs oracle c1, quick sweep timing, FP64 peak."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as ge
pkg = ge.load_package(); orc = ge.load_oracle(); abi = pkg.abi
ge.smoke()
eng = pkg.Engine(0)
print("fp64 peak TF/s:", eng.fp64_peak_tflops())
opts = abi.SetupOpts.default(L_minus_override=33.84706312634513)
desc, info = pkg.setup_front(3, [1.0, .98, .96, .04, .02], opts)
t = time.time(); res, ser = eng.frame_count_series(desc); print("gpu series time", time.time() - t)
c1, cser = orc.run_front(abi, desc, orc.MODE_C1, series=True)
print("gpu: status", res.status, "zeros", res.zero_count, "fail", res.failure_count, "steps", res.steps)
print("c1 : status", c1.status, "zeros", c1.zero_count, "fail", c1.failure_count, "steps", c1.steps)
print("series equal:", np.array_equal(ser, cser), " result equal:", bytes(res) == bytes(c1))
if not np.array_equal(ser, cser):
    bad = np.argwhere(ser != cser); print("first diffs", bad[:10]); i = bad[0][0]; print(ser[i], cser[i])
if bytes(res) != bytes(c1):
    a = np.frombuffer(bytes(res), dtype=np.float64); b = np.frombuffer(bytes(c1), dtype=np.float64); print("result diff idx", np.argwhere(a != b)[:20].ravel())
# sweep timing
for (Nc, Nr) in ((24, 4), (128, 8)):
    params = pkg.parameter_list(Nc, Nr)
    t = time.time(); descs, infos = pkg.setup_sweep(3, params); ts = time.time() - t
    t = time.time(); out = eng.frame_count_batch(descs); tg = time.time() - t
    t = time.time(); out = eng.frame_count_batch(descs); tg2 = time.time() - t
    zc = np.array([o.zero_count for o in out]); fc = np.array([o.failure_count for o in out]); st = np.array([o.status for o in out])
    print("sweep %d fronts: setup %.2fs gpu e2e %.4fs (2nd %.4fs) kernel %.3f ms; zeros hist %s failures>0: %d status!=0: %d" %
          (len(descs), ts, tg, tg2, eng.last_kernel_ms(), np.bincount(zc).tolist(), int((fc > 0).sum()), int((st != 0).sum())))
    fl = sum(pkg.front_flops(3, 20, 14, o.steps) for o in out)
    print("   fronts/s (kernel) %.0f ; model TFLOP/s %.2f" % (len(descs) / (eng.last_kernel_ms() * 1e-3), fl / (eng.last_kernel_ms() * 1e-3) / 1e12))
