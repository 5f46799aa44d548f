#!/bin/bash
set -u
mkdir -p gpurun_out/c12
O=gpurun_out/c12
timeout 900 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/summary.txt
for b in 1024 1024; do timeout 600 python bench.py --workload ensemble1d --steps 20 --warmup 5 --batch $b >> $O/ensemble1d.json 2>> $O/ensemble1d.err; done
cat $O/summary.txt; tail -6 $O/pytest.log; cat $O/ensemble1d.json
