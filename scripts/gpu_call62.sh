#!/bin/bash
set -u
O=gpurun_out/c62; mkdir -p $O
( time timeout 900 python bench.py --horizon 100 --steps 3 --warmup 3 --no-e2e --no-cpu --no-parity --no-slab --no-ensemble ) > $O/bench_conc.json 2> $O/bench_conc.err; echo rc=$?
python - <<'PY'
import json
d=json.loads(open("gpurun_out/c62/bench_conc.json").read().strip().splitlines()[-1])
print(d["value"], json.dumps(d["concurrent_problems"], indent=1))
PY
tail -5 gpurun_out/c62/bench_conc.err
