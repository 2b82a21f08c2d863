#!/bin/bash
# session 4, run b: BVP sweep with the chunked host pipeline (1, 2, 3 chunks), first rung 2^-2, the shared-integration GPU test
tag=${1:-s4b}; shift
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_host_mirrors.py tests/test_cpp_shim.py tests/test_flow_map.py -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -6 gpurun_out/${tag}_pytest.log
g++ -std=c++17 -O1 -o tests/cpp/bvp_sweep tests/cpp/bvp_sweep.cpp -L computing-conjugate-points_b200 -lccp_b200 -lccp_setup -Wl,-rpath,$PWD/computing-conjugate-points_b200 || exit 1
for ch in 1 2 3 2; do ./tests/cpp/bvp_sweep 128 8 3 1 $ch | tee -a gpurun_out/${tag}_bvp_sweep.log; done
./tests/cpp/bvp_sweep 128 8 2 1 2 | tee -a gpurun_out/${tag}_bvp_sweep.log
./tests/cpp/bvp_sweep 128 8 3 0 2 | tee -a gpurun_out/${tag}_bvp_sweep.log
./tests/cpp/bvp_sweep 24 4 2 1 2 | tee -a gpurun_out/${tag}_bvp_sweep.log
./tests/cpp/bvp_sweep 24 4 3 1 2 | tee -a gpurun_out/${tag}_bvp_sweep.log
