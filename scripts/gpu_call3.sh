#!/bin/bash
set -u
mkdir -p gpurun_out/c3
O=gpurun_out/c3
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/summary.txt
timeout 900 python bench.py --horizon 300 --steps 2 --warmup 2 --no-e2e --no-cpu > $O/bench_h300.json 2> $O/bench_h300.err; echo "bench rc=$?" >> $O/summary.txt
VCH_BICG6=0 timeout 900 python bench.py --horizon 300 --steps 2 --warmup 2 --no-e2e --no-cpu > $O/bench_h300_7launch.json 2> $O/bench_h300_7launch.err; echo "bench7 rc=$?" >> $O/summary.txt
VCH_DEBUG=1 timeout 600 python scripts/newton_floor_probe.py 1024 1000 > $O/floor_probe.out 2> $O/floor_probe.err; echo "floor probe rc=$?" >> $O/summary.txt
grep "newton:" $O/floor_probe.err | awk '{print $3, $7, $10, $15}' | tr -d ',' > $O/floor_probe.txt; rm -f $O/floor_probe.err
cat $O/summary.txt; tail -3 $O/pytest.log
