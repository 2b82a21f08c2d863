"""Builds tests/golden/sweep96_ref_2p5e-13.npz: oracle mode `ref` (the reference-structured restatement) on the reference's own 96-point sweep
(Construct_ParameterList 24 x 4, heteroclinic.cpp:437-441) with the neighbourhood of the connecting orbit a proof pipeline delivers
(Krawczyk radius 2.5e-13).  It certifies only part of the sweep -- the frame basis Dphi_t A_u is ill-conditioned late in the
run -- which is what tests/test_oracle_golden.py::test_rigorous_box_certifies_where_the_unnormalised_frame_cannot compares the kernel's
algorithm with.  About 3 minutes on 8 cores:  python tests/golden/make_sweep96_golden.py
"""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as g

pkg = g.load_package(); orc = g.load_oracle(); orc.lib(pkg.abi)
descs, _ = pkg.setup_sweep(3, pkg.parameter_list(24, 4, 0.02, 0.06), pkg.abi.SetupOpts.default(xy_nbd=2.5e-13))
res, sec = orc.run_batch(pkg.abi, descs, orc.MODE_REF, threads=0)
st = np.array([r.status for r in res], np.int32); f = np.array([r.failure_count for r in res], np.int32)
dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "sweep96_ref_2p5e-13.npz")
np.savez_compressed(dst, status=st, zero_count=np.array([r.zero_count for r in res], np.int32), failure_count=f,
                    steps=np.array([r.steps for r in res], np.int32), certified=np.int32(((st == 0) & (f == 0)).sum()))
print("wrote", dst, "in %.0f s;" % sec, int(((st == 0) & (f == 0)).sum()), "of 96 certified by mode ref")
