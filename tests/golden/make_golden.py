"""Builds tests/golden/plot_det_golden.npz from the reference's only recorded output,
/root/reference/plot_det.txt (written by topFrame::makePlot, topFrame.cpp:177-191 through plot(),
utils.cpp:126-135): one row per sub-interval, `t_mid det_mid t_rad det_rad` of det_series[i][2] for the
n=3 default parameters b=(1,.98,.96), c=(+.04,+.02) (heteroclinic.cpp:524-529).
Run in the build container only (the GPU box has no /root/reference):  python tests/golden/make_golden.py
"""
import os
import numpy as np

src = "/root/reference/plot_det.txt"
dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "plot_det_golden.npz")
a = np.loadtxt(src)
assert a.shape == (15176, 4)
np.savez_compressed(dst, t_mid=a[:, 0], det_mid=a[:, 1], t_rad=a[:, 2], det_rad=a[:, 3],
                    L_minus=np.float64(a[-1, 0] + a[-1, 2]), note="reference plot_det.txt, n=3 defaults, grid 14, step 2^-5, order 20")
print("wrote", dst, os.path.getsize(dst), "bytes; L_minus =", repr(float(a[-1, 0] + a[-1, 2])))
