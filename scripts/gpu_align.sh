#!/bin/bash
# per-SM alignment variants (CCP_DBG_ALIGN = 0 none | 1 all fronts of an SM in phase | 1 + (mode << 1): half of them offset by half a step)
mkdir -p gpurun_out
for rep in 1 2; do
for a in ${ALIGNS:-0 1 3 5 7}; do
  for n in 592 1036 1184; do echo -n "align=$a: "; CCP_DBG_ALIGN=$a CCP_REPS=3 python scripts/gpu_one.py $n | cut -d' ' -f1-14; done
done
done 2>&1 | tee gpurun_out/${1:-align}.txt
