#!/bin/bash
set -u
O=gpurun_out/c49; mkdir -p $O
timeout 300 python scripts/concurrency_bisect_probe.py 128 100 3 3 > $O/bisect_128.txt 2>&1
timeout 300 python scripts/concurrency_bisect_probe.py 64 50 3 3 > $O/bisect_64.txt 2>&1
timeout 300 python scripts/concurrency_bisect_probe.py 1024 30 2 2 > $O/bisect_1024.txt 2>&1
cat $O/bisect_128.txt $O/bisect_64.txt $O/bisect_1024.txt
