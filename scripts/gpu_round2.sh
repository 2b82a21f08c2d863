#!/bin/bash
# full GPU test-suite + the additional bench workloads of round 2
tag=${1:-r2}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -15 gpurun_out/${tag}_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err; echo "bench rc=$?"
cut -c1-600 gpurun_out/bench_${tag}.json
timeout 600 python bench.py --workload subbox4096 --steps 3 --warmup 1 > gpurun_out/bench_${tag}_subbox4096.json 2> gpurun_out/bench_${tag}_subbox4096.err; echo "subbox rc=$?"
cut -c1-900 gpurun_out/bench_${tag}_subbox4096.json
timeout 600 python bench.py --n 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${tag}_n2.json 2> gpurun_out/bench_${tag}_n2.err; echo "n2 rc=$?"
timeout 600 python bench.py --n 4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${tag}_n4.json 2> gpurun_out/bench_${tag}_n4.err; echo "n4 rc=$?"
python - <<'PY'
import json
for t in ("n2","n4"):
    try:
        d=json.load(open("gpurun_out/bench_%s_%s.json"%("",t)))
        print(t, d["value"], d["ms_per_step"], d["certified"], d["roofline"]["frac"])
    except Exception as e: print(t, "failed", e)
PY
