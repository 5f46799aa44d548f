#!/bin/bash
set -u
O=gpurun_out/c25; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/summary.txt
cat $O/summary.txt; tail -40 $O/pytest.log | cut -c1-250
