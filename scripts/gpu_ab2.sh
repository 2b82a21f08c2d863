#!/bin/bash
# same-box A/B of kernel variant builds (libccp_b200_<v>.so) against the main build: 1,036 and 1,184 identical fronts, two rounds
tag=${1:-ab}; shift
mkdir -p gpurun_out
( for rep in 1 2; do
for n in 1036 1184; do echo -n "main: "; CCP_REPS=3 python scripts/gpu_one.py $n | cut -d' ' -f1-14; done
for v in "$@"; do for n in 1036 1184; do echo -n "$v: "; CCP_LIB=libccp_b200_$v.so CCP_REPS=3 python scripts/gpu_one.py $n | cut -d' ' -f1-14; done; done
done ) 2>&1 | tee gpurun_out/${tag}_sizes.txt
