#!/bin/bash
set -u
O=gpurun_out/c51; mkdir -p $O
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in racecheck initcheck memcheck; do
  timeout 280 $CS --tool $tool --print-limit 20 python scripts/sanitizer_target.py 128 1 > $O/${tool}_128.txt 2>&1; echo "$tool 128 rc=$?" >> $O/summary.txt
  VCH_COLS_TMA=0 timeout 280 $CS --tool $tool --print-limit 20 python scripts/sanitizer_target.py 512 1 > $O/${tool}_512_notma.txt 2>&1; echo "$tool 512 notma rc=$?" >> $O/summary.txt
done
timeout 280 $CS --tool racecheck --print-limit 20 python scripts/sanitizer_target.py 512 1 > $O/racecheck_512_tma.txt 2>&1; echo "racecheck 512 tma rc=$?" >> $O/summary.txt
cat $O/summary.txt
for f in $O/*_*.txt; do echo "== $f"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Uninitialized|Invalid|ok " $f | sort | uniq -c | head -12; done
