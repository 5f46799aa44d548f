#!/bin/bash
set -u
O=gpurun_out/c54; mkdir -p $O
PROBE_REPS=16 timeout 200 python scripts/determinism_probe.py 1024 8 4 > $O/det_1024.txt 2>&1
VCH_COLS_TMA=0 PROBE_REPS=16 timeout 200 python scripts/determinism_probe.py 1024 8 4 > $O/det_1024_notma.txt 2>&1
timeout 300 python scripts/concurrency_bisect_probe.py 128 100 3 4 > $O/bisect_128.txt 2>&1
timeout 300 python scripts/concurrency_bisect_probe.py 1024 30 2 3 > $O/bisect_1024.txt 2>&1
PROBE_VERBOSE=1 timeout 300 python scripts/concurrent_problems_probe.py 1024 100 3 3 > $O/concurrent_1024.txt 2>&1
for f in det_1024 det_1024_notma; do echo "== $f"; grep -E "jacobian_solve|newton step|vs run" $O/$f.txt; done
for f in bisect_128 bisect_1024; do echo "== $f"; grep -c "False" $O/$f.txt; grep "False" $O/$f.txt | head -5; done
cat $O/concurrent_1024.txt
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest.log
