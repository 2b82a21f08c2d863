#!/bin/bash
# A/B timing in ONE box: current build vs ab/<name>.so (boxes differ by a few per cent, so only same-box pairs compare)
L=computing-conjugate-points_b200/libccp_b200.so
cp $L /tmp/new.so
for rep in 1 2; do
  for v in "$@" new; do
    if [ $v = new ]; then cp /tmp/new.so $L; else cp ab/$v.so $L; fi
    echo -n "$v: "; CCP_REPS=3 python scripts/gpu_one.py 1024 | cut -d' ' -f1-14
  done
done
for v in "$@" new; do
  if [ $v = new ]; then cp /tmp/new.so $L; else cp ab/$v.so $L; fi
  echo -n "$v 148: "; CCP_REPS=3 python scripts/gpu_one.py 148 | cut -d' ' -f1-14
done
cp /tmp/new.so $L
