#!/bin/bash
set -u
mkdir -p gpurun_out/c11
O=gpurun_out/c11
timeout 900 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/summary.txt
timeout 600 python bench.py --workload ensemble1d --steps 5 --warmup 2 > $O/ensemble1d.json 2> $O/ensemble1d.err; echo "ens rc=$?" >> $O/summary.txt
timeout 1200 python scripts/parity_vs_strict.py 1024 1000 study > $O/parity_study_1024x1000.jsonl 2> $O/parity_study.err; echo "study rc=$?" >> $O/summary.txt
timeout 900 python bench.py --horizon 300 --steps 2 --warmup 2 --no-e2e --no-cpu --no-parity --no-slab > $O/bench_h300.json 2> $O/bench_h300.err; echo "bench rc=$?" >> $O/summary.txt
cat $O/summary.txt; tail -6 $O/pytest.log; cat $O/ensemble1d.json; cat $O/parity_study_1024x1000.jsonl
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c11/bench_h300.json').read().strip().splitlines()[-1])
print("h300 it/s", d['value'], "ms/step", d['ms_per_step'], d['solver'])
PY
