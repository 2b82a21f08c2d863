#!/bin/bash
# ncu evidence for one kernel version: launch list of a short bench run + one --set full capture of front_kernel<3> (1,024 identical fronts)
tag=${1:-vX}
mkdir -p gpurun_out
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1 || exit 1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_${tag}.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_${tag}_l.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:front_kernel -c 1 -o gpurun_out/${tag}_full -f python scripts/gpu_one.py ${2:-1024} > gpurun_out/ncu_${tag}.log 2>&1; echo "ncu rc=$?"
