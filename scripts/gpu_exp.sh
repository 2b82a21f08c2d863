#!/bin/bash
# quick check of a kernel change: bit-parity tests, timings at several loads (main build, then variant builds libccp_b200_<v>.so), cycle trace of one step
tag=${1:-exp}; shift
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -4 gpurun_out/${tag}_pytest.log
( for n in 148 592 1036 1184; do echo -n "main: "; CCP_REPS=3 python scripts/gpu_one.py $n | cut -d' ' -f1-14; done
for v in "$@"; do for n in 148 1036 1184; do echo -n "$v: "; CCP_LIB=libccp_b200_$v.so CCP_REPS=3 python scripts/gpu_one.py $n | cut -d' ' -f1-14; done; done ) 2>&1 | tee gpurun_out/${tag}_sizes.txt
python scripts/gpu_trace.py 148 > gpurun_out/${tag}_trace_148.txt 2>&1; bash scripts/ncu_quick.sh 1036 > /dev/null 2>&1; cp gpurun_out/ncu_quick.csv gpurun_out/${tag}_ncu_quick.csv
python scripts/gpu_trace.py 1184 > gpurun_out/${tag}_trace_1184.txt 2>&1
