#!/bin/bash
# quick kernel iteration: parity tests of the front kernel, then timings at several loads.  gpurun -- 'bash scripts/gpu_iter.sh tag'
tag=${1:-it}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_flow_map.py tests/test_subdivision.py -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -4 gpurun_out/${tag}_pytest.log
for n in 148 1036 1184 8192; do CCP_REPS=3 python scripts/gpu_one.py $n | tail -1; done | tee gpurun_out/${tag}_sizes.txt
CCP_REPS=3 python scripts/gpu_one.py 1024 0 2 | tail -1 | tee -a gpurun_out/${tag}_sizes.txt
CCP_REPS=3 python scripts/gpu_one.py 1024 0 4 | tail -1 | tee -a gpurun_out/${tag}_sizes.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_${tag}.json')); print('bench', d['value'], d['ms_per_step'], 'certified', d['certified'], 'frac', d['roofline']['frac'])"
