#!/bin/bash
# boundary-value stage timing: batched Krawczyk steps of 1,024 fronts at several first rungs of the step ladder, and the reference's 96-point sweep
mkdir -p gpurun_out
g++ -std=c++17 -O1 -o tests/cpp/bvp_sweep tests/cpp/bvp_sweep.cpp -L computing-conjugate-points_b200 -lccp_b200 -lccp_setup -Wl,-rpath,$PWD/computing-conjugate-points_b200 || exit 1
tag=${1:-bvp}; shift
for st in "$@"; do ./tests/cpp/bvp_sweep 128 8 $st | tee -a gpurun_out/${tag}_bvp_sweep.log; done
./tests/cpp/bvp_sweep 24 4 | tee -a gpurun_out/${tag}_bvp_sweep.log
