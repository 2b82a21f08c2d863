#!/bin/bash
# session 4, run f: stage times of the complete-proof sweep driver; boundDFU kernel at 2 / 3 / 4 CTAs per SM
tag=${1:-s4f}
mkdir -p gpurun_out
g++ -std=c++17 -O1 -o tests/cpp/driver_sweep_gpu tests/cpp/driver_sweep_gpu.cpp -L computing-conjugate-points_b200 -lccp_b200 -lccp_setup -Wl,-rpath,$PWD/computing-conjugate-points_b200 && for a in "128 8"; do timeout 900 ./tests/cpp/driver_sweep_gpu $a | cut -c1-300 | tee -a gpurun_out/${tag}_driver_sweep.log; done
nproc
( python scripts/gpu_dfu.py 65536; for v in dfu2 dfu3 dfu4; do echo -n "$v: "; CCP_LIB=libccp_b200_$v.so python scripts/gpu_dfu.py 65536; done; python scripts/gpu_dfu.py 65536 ) 2>&1 | tee gpurun_out/${tag}_dfu.txt
