"""Builds tests/golden/sweep1024_ref.npz: status, zero_count, failure_count, steps of oracle mode `ref` (the reference-structured
restatement: n separate 4n-dimensional C^0 Lohner integrations of f_linearize, topFrame series, countZeros) on all 1,024 fronts of the
bench sweep (Construct_ParameterList 128 x 8, heteroclinic.cpp:386-427; BASELINE configs[1]).  About 16 minutes on 8 cores, which
is why the result is committed instead of being recomputed by the test-suite:  python tests/golden/make_sweep_golden.py
"""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as g

pkg = g.load_package(); orc = g.load_oracle(); orc.lib(pkg.abi)
# thin box (xy_nbd = 1e-19): the box on which the reference-structured algorithm certifies (tests/conftest.py:thin_opts)
descs, _ = pkg.setup_sweep(3, pkg.parameter_list(128, 8, 0.02, 0.06), pkg.abi.SetupOpts.default(xy_nbd=1e-19))
res, sec = orc.run_batch(pkg.abi, descs, orc.MODE_REF, threads=0)
dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "sweep1024_ref.npz")
np.savez_compressed(dst, status=np.array([r.status for r in res], np.int32), zero_count=np.array([r.zero_count for r in res], np.int32),
                    failure_count=np.array([r.failure_count for r in res], np.int32), steps=np.array([r.steps for r in res], np.int32))
print("wrote", dst, "in %.0f s" % sec)
