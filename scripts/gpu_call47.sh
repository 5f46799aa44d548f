#!/bin/bash
# reference's own pytest files against the final build's drop-ins + concurrent-problems probe
set -u
O=gpurun_out/c47; mkdir -p $O
( time timeout 900 scripts/run_reference_tests.sh run ) > $O/reference_own_tests.log 2>&1; echo "reftests rc=$?" >> $O/summary.txt
timeout 600 python scripts/concurrent_problems_probe.py 1024 100 3 3 > $O/concurrent_1024.txt 2>&1; echo "conc1024 rc=$?" >> $O/summary.txt
timeout 600 python scripts/concurrent_problems_probe.py 128 100 5 4 > $O/concurrent_128.txt 2>&1; echo "conc128 rc=$?" >> $O/summary.txt
timeout 600 python scripts/concurrent_problems_probe.py 512 100 3 3 > $O/concurrent_512.txt 2>&1; echo "conc512 rc=$?" >> $O/summary.txt
cat $O/summary.txt; grep -E "passed|failed|error" $O/reference_own_tests.log | tail -4; tail -3 $O/concurrent_1024.txt $O/concurrent_128.txt $O/concurrent_512.txt
