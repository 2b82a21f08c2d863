#!/bin/bash
# compute-sanitizer on the smallest case that exercises every phase of front_kernel (39 steps: one basis renormalisation, a partial
# last step), the config-4 kernels and boundDFU.  One tool per call:  gpurun -- 'bash scripts/gpu_sanitize.sh racecheck|memcheck|synccheck'
tool=${1:-racecheck}
mkdir -p gpurun_out
cat > /tmp/san_case.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np
import __graft_entry__ as ge
pkg = ge.load_package(); abi = pkg.abi
eng = pkg.Engine(0)
opts = abi.SetupOpts.default(L_minus_override=1.2)
for n, par in ((3, [1.0, .98, .96, .04, .02]), (2, [1.0, .97, .05]), (4, [1.0, .98, .96, .94, .04, .02, .03])):
    desc, _ = pkg.setup_front(n, par, opts)
    out = eng.frame_count_batch((abi.FrontDesc * 3)(desc, desc, desc))
    print("n", n, "steps", out[0].steps, "status", out[0].status, "failures", out[0].failure_count)
desc, _ = pkg.setup_front(3, [1.0, .98, .96, .04, .02], opts)
res, ser = eng.frame_count_series(desc)
print("series", ser.shape)
sub = eng.frame_count_subdivided((abi.FrontDesc * 2)(desc, desc), 2, abi.CCP_SUBDIV_UXY)
print("subdivided", sub[0].certified, sub[0].zero_count)
A = np.array([[desc.A_u[i * abi.CCP_MAX_DIM + j] for j in range(6)] for i in range(6)])
U = np.array([[-1e-6, 1e-6]] * 3 + [[-1e-11, 1e-11]] * 3)
d = pkg.dfu_desc(3, [1.0, .98, .96, .04, .02], A, [0.0] * 6, U, subdivision=4)
print("dfu", eng.bound_dfu_batch((abi.DfuDesc * 2)(d, d))[0].parts)
PY
timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san_case.py > gpurun_out/sanitizer_${tool}.log 2>&1; echo "rc=$?" >> gpurun_out/sanitizer_${tool}.log
tail -25 gpurun_out/sanitizer_${tool}.log
