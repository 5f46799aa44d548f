#!/bin/bash
set -u
mkdir -p gpurun_out/c8
O=gpurun_out/c8
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/summary.txt
timeout 1500 python bench.py --steps 3 --warmup 3 --no-cpu > $O/bench_full.json 2> $O/bench_full.err; echo "bench rc=$?" >> $O/summary.txt
cat $O/summary.txt; tail -5 $O/pytest.log
