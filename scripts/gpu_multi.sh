#!/bin/bash
# N-GPU call: library multi-GPU test (one process, all devices), torchrun bench weak + strong (config 3: 65,536 parameters in total)
N=${1:-2}; tag=${2:-r2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_sweep_gpu.py -m gpu -q -k multi_gpu > gpurun_out/${tag}_pytest_${N}gpu.log 2>&1; tail -3 gpurun_out/${tag}_pytest_${N}gpu.log
./tests/cpp/multi_gpu 4099 | tee gpurun_out/${tag}_multi_gpu_${N}.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_${tag}_${N}gpu.json 2> gpurun_out/bench_${tag}_${N}gpu.err; echo "weak rc=$?"
timeout 900 $TR bench.py --gpus $N --total-fronts 65536 --steps 3 --warmup 1 > gpurun_out/bench_${tag}_65536_${N}gpu.json 2> gpurun_out/bench_${tag}_65536_${N}gpu.err; echo "strong rc=$?"
python - <<PY
import json
for f in ("gpurun_out/bench_${tag}_${N}gpu.json","gpurun_out/bench_${tag}_65536_${N}gpu.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["n_gpus"], d["scaling"], round(d["value"]), round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"]), "certified", d["certified"], "/", d["fronts_total"])
    except Exception as e: print(f, "failed", e)
PY
