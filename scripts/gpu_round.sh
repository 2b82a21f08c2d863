#!/bin/bash
# One gpurun call of a kernel iteration: GPU tests, bench, ncu launch list, one ncu --set full capture of front_kernel<3>.
#   gpurun --timeout 1500 -- 'bash scripts/gpu_round.sh v21 [notests]'
tag=${1:-vX}; mode=${2:-all}
mkdir -p gpurun_out
if [ "$mode" != "notests" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
  tail -5 gpurun_out/${tag}_pytest.log
fi
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err; echo "bench rc=$?"
cat gpurun_out/bench_${tag}.json | cut -c1-1500
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_${tag}.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_${tag}_l.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:front_kernel -c 1 -o gpurun_out/${tag}_full -f python scripts/gpu_one.py 1024 > gpurun_out/ncu_${tag}.log 2>&1; echo "ncu rc=$?"
CCP_REPS=3 python scripts/gpu_one.py 1024 | tail -1
CCP_REPS=3 python scripts/gpu_one.py 148 | tail -1
CCP_REPS=2 python scripts/gpu_one.py 8192 | tail -1
CCP_REPS=3 python scripts/gpu_one.py 1024 0 2 | tail -1
CCP_REPS=3 python scripts/gpu_one.py 1024 0 4 | tail -1
