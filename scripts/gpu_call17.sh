#!/bin/bash
set -u
O=gpurun_out/c17; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/summary.txt
B="python bench.py --horizon 300 --steps 2 --warmup 2 --no-e2e --no-cpu --no-parity --no-slab"
timeout 600 $B > $O/bench_h300.json 2> $O/bench_h300.err; echo "bench rc=$?" >> $O/summary.txt
cat $O/summary.txt; tail -5 $O/pytest.log
python - <<'PY'
import json
d = json.loads(open("gpurun_out/c17/bench_h300.json").read().strip().splitlines()[-1])
print("it/s", round(d["value"], 4), "ms/step", round(d["ms_per_step"], 1), d["solver"], "launches", d["gpu_launches"])
PY
