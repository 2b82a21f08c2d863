#!/bin/bash
# session 4, run c: full GPU test-suite of the new build, then same-box A/B against variant builds (libccp_b200_<v>.so)
tag=${1:-s4c}; shift
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -6 gpurun_out/${tag}_pytest.log
( for rep in 1 2; do
for n in 148 1036 1184; do echo -n "main: "; CCP_REPS=3 python scripts/gpu_one.py $n | cut -d' ' -f1-14; done
for v in "$@"; do for n in 148 1036 1184; do echo -n "$v: "; CCP_LIB=libccp_b200_$v.so CCP_REPS=3 python scripts/gpu_one.py $n | cut -d' ' -f1-14; done; done
done
echo -n "main n=2: "; CCP_REPS=3 python scripts/gpu_one.py 1024 0 2 | cut -d' ' -f1-14
for v in "$@"; do echo -n "$v n=2: "; CCP_LIB=libccp_b200_$v.so CCP_REPS=3 python scripts/gpu_one.py 1024 0 2 | cut -d' ' -f1-14; done
) 2>&1 | tee gpurun_out/${tag}_sizes.txt
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/bench_${tag}.json
