#!/bin/bash
set -u
mkdir -p gpurun_out/c10
O=gpurun_out/c10
timeout 900 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/summary.txt
( time timeout 1700 python bench.py ) > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?" >> $O/summary.txt
cat $O/summary.txt; tail -8 $O/pytest.log; tail -5 $O/bench_default.err
