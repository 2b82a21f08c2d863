#!/bin/bash
# evidence set of a kernel version: full GPU test-suite, default bench line, ncu launch list + full capture, a few sizes
tag=${1:-v27}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -4 gpurun_out/${tag}_pytest.log
python bench.py > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err; echo "bench rc=$?"; cut -c1-160 gpurun_out/bench_${tag}.json
bash scripts/gpu_prof.sh ${tag}
CCP_REPS=3 python scripts/gpu_one.py 148 | tail -1
CCP_REPS=2 python scripts/gpu_one.py 8192 | tail -1
CCP_REPS=3 python scripts/gpu_one.py 1024 0 2 | tail -1
CCP_REPS=3 python scripts/gpu_one.py 1024 0 4 | tail -1
timeout 900 python bench.py --total-fronts 65536 --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/bench_${tag}_65536_1gpu.json 2> gpurun_out/bench_${tag}_65536_1gpu.err; echo "65536 rc=$?"; cut -c1-160 gpurun_out/bench_${tag}_65536_1gpu.json
