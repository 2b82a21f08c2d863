#!/bin/bash
# session 4, run d: parity tests of the new build, same-box A/B against variants, default bench line, ncu launch list + full capture
tag=${1:-s4d}; shift
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_subdivision.py tests/test_flow_map.py -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -4 gpurun_out/${tag}_pytest.log
( for rep in 1 2; do
for n in 1036 1184; do echo -n "main: "; CCP_REPS=3 python scripts/gpu_one.py $n | cut -d' ' -f1-14; done
for v in "$@"; do for n in 1036 1184; do echo -n "$v: "; CCP_LIB=libccp_b200_$v.so CCP_REPS=3 python scripts/gpu_one.py $n | cut -d' ' -f1-14; done; done
done ) 2>&1 | tee gpurun_out/${tag}_sizes.txt
python bench.py > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/bench_${tag}.json
bash scripts/gpu_prof.sh ${tag}
python scripts/gpu_one.py 8192; python scripts/gpu_one.py 1024 0 2; python scripts/gpu_one.py 1024 0 4
