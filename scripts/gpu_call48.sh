#!/bin/bash
set -u
O=gpurun_out/c48; mkdir -p $O
PROBE_VERBOSE=1 timeout 300 python scripts/concurrent_problems_probe.py 128 100 5 3 > $O/concurrent_128.txt 2>&1
PROBE_VERBOSE=1 timeout 300 python scripts/concurrent_problems_probe.py 1024 100 3 2 > $O/concurrent_1024.txt 2>&1
cat $O/concurrent_128.txt $O/concurrent_1024.txt
