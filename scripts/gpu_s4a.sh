#!/bin/bash
# session 4, run a: full GPU test-suite, BVP sweep with / without shared integrations, A/B of kernel variants
tag=${1:-s4a}; shift
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -6 gpurun_out/${tag}_pytest.log
g++ -std=c++17 -O1 -o tests/cpp/bvp_sweep tests/cpp/bvp_sweep.cpp -L computing-conjugate-points_b200 -lccp_b200 -lccp_setup -Wl,-rpath,$PWD/computing-conjugate-points_b200 || exit 1
for sh in 1 0 1; do ./tests/cpp/bvp_sweep 128 8 3 $sh | tee -a gpurun_out/${tag}_bvp_sweep.log; done
for sh in 1 0; do ./tests/cpp/bvp_sweep 24 4 3 $sh | tee -a gpurun_out/${tag}_bvp_sweep.log; done
( for rep in 1 2; do
for n in 1036 1184; do echo -n "main: "; CCP_REPS=3 python scripts/gpu_one.py $n | cut -d' ' -f1-14; done
for v in "$@"; do for n in 1036 1184; do echo -n "$v: "; CCP_LIB=libccp_b200_$v.so CCP_REPS=3 python scripts/gpu_one.py $n | cut -d' ' -f1-14; done; done
done ) 2>&1 | tee gpurun_out/${tag}_sizes.txt
