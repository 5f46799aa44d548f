#!/bin/bash
set -u
mkdir -p gpurun_out/c9
O=gpurun_out/c9
timeout 900 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/summary.txt
timeout 900 python bench.py --horizon 300 --steps 2 --warmup 2 --no-e2e --no-cpu --no-parity --no-slab > $O/bench_h300.json 2> $O/bench_h300.err; echo "bench rc=$?" >> $O/summary.txt
VCH_GRAPH_UNROLL=0 timeout 900 python bench.py --horizon 300 --steps 2 --warmup 2 --no-e2e --no-cpu --no-parity --no-slab > $O/bench_h300_u0.json 2> $O/bench_h300_u0.err; echo "bench u0 rc=$?" >> $O/summary.txt
VCH_GRAPH_UNROLL=3 timeout 900 python bench.py --horizon 300 --steps 2 --warmup 2 --no-e2e --no-cpu --no-parity --no-slab > $O/bench_h300_u3.json 2> $O/bench_h300_u3.err; echo "bench u3 rc=$?" >> $O/summary.txt
cat $O/summary.txt; tail -8 $O/pytest.log
python - <<'PY'
import json
for v in ("h300","h300_u0","h300_u3"):
    try:
        d=json.loads(open(f'gpurun_out/c9/bench_{v}.json').read().strip().splitlines()[-1])
        print(v, "it/s", d['value'], "ms/step", d['ms_per_step'], d['solver'])
    except Exception as e: print(v, e, open(f'gpurun_out/c9/bench_{v}.err').read()[-800:])
PY
