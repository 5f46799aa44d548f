#!/bin/bash
set -u
mkdir -p gpurun_out/c2
O=gpurun_out/c2
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/summary.txt
timeout 1200 python scripts/parity_vs_strict.py 1024 1000 study > $O/parity_study_1024x1000.jsonl 2> $O/parity_study.err; echo "study rc=$?" >> $O/summary.txt
timeout 600 python scripts/parity_vs_strict.py 512 1000 study > $O/parity_study_512x1000.jsonl 2>> $O/parity_study.err; echo "study512 rc=$?" >> $O/summary.txt
cat $O/summary.txt; tail -3 $O/pytest.log
