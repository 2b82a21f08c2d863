#!/bin/bash
# A/B of kernel builds in ONE box: main build and variant builds libccp_b200_<v>.so; parity test of the main build first
tag=${1:-ab}; shift
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
( for rep in 1 2; do
for n in 148 1036 1184; do echo -n "main: "; CCP_REPS=3 python scripts/gpu_one.py $n | cut -d' ' -f1-14; done
for v in "$@"; do for n in 148 1036 1184; do echo -n "$v: "; CCP_LIB=libccp_b200_$v.so CCP_REPS=3 python scripts/gpu_one.py $n | cut -d' ' -f1-14; done; done
done ) 2>&1 | tee gpurun_out/${tag}_sizes.txt
