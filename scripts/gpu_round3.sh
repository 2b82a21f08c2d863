#!/bin/bash
# full GPU test-suite, strong-scaling point N=1 of config 3, BVP sweep timing
tag=${1:-r2b}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -6 gpurun_out/${tag}_pytest.log
timeout 900 python bench.py --total-fronts 65536 --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/bench_${tag}_65536_1gpu.json 2> gpurun_out/bench_${tag}_65536_1gpu.err; echo "strong N=1 rc=$?"
cut -c1-330 gpurun_out/bench_${tag}_65536_1gpu.json
g++ -std=c++17 -O1 -o tests/cpp/bvp_sweep tests/cpp/bvp_sweep.cpp -L computing-conjugate-points_b200 -lccp_b200 -lccp_setup -Wl,-rpath,$PWD/computing-conjugate-points_b200
for st in 3 5; do ./tests/cpp/bvp_sweep 128 8 $st | tee -a gpurun_out/${tag}_bvp_sweep.log; done
./tests/cpp/bvp_sweep 24 4 | tee -a gpurun_out/${tag}_bvp_sweep.log
