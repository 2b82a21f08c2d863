#!/bin/bash
# session 4, run g: full GPU test-suite after the host-side changes, complete-proof sweep driver with stage times
tag=${1:-s4g}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -5 gpurun_out/${tag}_pytest.log
g++ -std=c++17 -O1 -o tests/cpp/driver_sweep_gpu tests/cpp/driver_sweep_gpu.cpp -L computing-conjugate-points_b200 -lccp_b200 -lccp_setup -Wl,-rpath,$PWD/computing-conjugate-points_b200 && for a in "24 4" "128 8" "128 8"; do timeout 900 ./tests/cpp/driver_sweep_gpu $a | cut -c1-300 | tee -a gpurun_out/${tag}_driver_sweep.log; done
