#!/bin/bash
set -u
mkdir -p gpurun_out/c6
O=gpurun_out/c6
./scripts/fft_scaling 2>&1 | tail -19 > $O/fft_scaling.txt
./scripts/kernel_bench > $O/kernel_bench.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/summary.txt
timeout 900 python bench.py --horizon 300 --steps 2 --warmup 2 --no-e2e --no-cpu > $O/bench_h300.json 2> $O/bench_h300.err; echo "bench rc=$?" >> $O/summary.txt
cat $O/fft_scaling.txt $O/kernel_bench.txt $O/summary.txt; tail -3 $O/pytest.log
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c6/bench_h300.json').read().strip().splitlines()[-1])
print("it/s", d['value'], "ms/step", d['ms_per_step'], d['solver'])
PY
