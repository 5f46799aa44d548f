#!/bin/bash
set -u
O=gpurun_out/c50; mkdir -p $O
timeout 300 python scripts/determinism_probe.py 1024 30 4 > $O/det_1024.txt 2>&1
timeout 300 python scripts/determinism_probe.py 512 60 4 > $O/det_512.txt 2>&1
VCH_COLS_TMA=0 timeout 300 python scripts/determinism_probe.py 1024 30 4 > $O/det_1024_notma.txt 2>&1
cat $O/det_1024.txt; echo ----; cat $O/det_512.txt; echo ---- no TMA; cat $O/det_1024_notma.txt
