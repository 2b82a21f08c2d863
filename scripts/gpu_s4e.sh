#!/bin/bash
# session 4, run e: evidence set of the final build (1 GPU): default bench line, reference arm, ncu launch list + full capture,
# other workloads (sub-boxes, n = 2, n = 4, 65,536 parameters), BVP sweep, complete-proof sweep driver
tag=${1:-v26}
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err; echo "bench rc=$?"; cut -c1-160 gpurun_out/bench_${tag}.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${tag}_reference_arm.json 2> gpurun_out/bench_${tag}_reference_arm.err; echo "reference arm rc=$?"; cut -c1-200 gpurun_out/bench_${tag}_reference_arm.json
bash scripts/gpu_prof.sh ${tag}
timeout 600 python bench.py --workload subbox4096 --steps 3 --warmup 1 > gpurun_out/bench_${tag}_subbox4096.json 2> gpurun_out/bench_${tag}_subbox4096.err; echo "subbox rc=$?"
timeout 600 python bench.py --n 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${tag}_n2.json 2> gpurun_out/bench_${tag}_n2.err; echo "n2 rc=$?"
timeout 600 python bench.py --n 4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${tag}_n4.json 2> gpurun_out/bench_${tag}_n4.err; echo "n4 rc=$?"
timeout 900 python bench.py --total-fronts 65536 --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/bench_${tag}_65536_1gpu.json 2> gpurun_out/bench_${tag}_65536_1gpu.err; echo "65536 rc=$?"
python - <<PY
import json
for t in ("subbox4096","n2","n4","65536_1gpu"):
    try:
        d=json.loads(open("gpurun_out/bench_${tag}_%s.json"%t).read().strip().splitlines()[-1]); print(t, round(d["value"]), round(d["ms_per_step"],2), d.get("certified"), d["roofline"]["frac"])
    except Exception as e: print(t, "failed", e)
PY
CCP_REPS=3 python scripts/gpu_one.py 148 | tail -1
CCP_REPS=2 python scripts/gpu_one.py 8192 | tail -1
g++ -std=c++17 -O1 -o tests/cpp/bvp_sweep tests/cpp/bvp_sweep.cpp -L computing-conjugate-points_b200 -lccp_b200 -lccp_setup -Wl,-rpath,$PWD/computing-conjugate-points_b200 && for a in "128 8" "24 4"; do ./tests/cpp/bvp_sweep $a | tee -a gpurun_out/${tag}_bvp_sweep.log; done
g++ -std=c++17 -O1 -o tests/cpp/driver_sweep_gpu tests/cpp/driver_sweep_gpu.cpp -L computing-conjugate-points_b200 -lccp_b200 -lccp_setup -Wl,-rpath,$PWD/computing-conjugate-points_b200 && for a in "24 4" "128 8"; do timeout 900 ./tests/cpp/driver_sweep_gpu $a | cut -c1-400 | tee -a gpurun_out/${tag}_driver_sweep.log; done
