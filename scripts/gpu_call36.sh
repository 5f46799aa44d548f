#!/bin/bash
set -u
O=gpurun_out/c36; mkdir -p $O
CMD="python bench.py --workload ensemble1d --steps 2 --warmup 1"
timeout 300 $CMD > $O/plain.json 2> $O/plain.err; echo "plain rc=$?" >> $O/summary.txt
timeout 420 ncu --set full --clock-control none -k 'regex:1d' -s 4 -c 8 -o $O/k1d -f $CMD > $O/ncu_1d.log 2>&1; echo "ncu rc=$?" >> $O/summary.txt
ncu -i $O/k1d.ncu-rep --page raw --csv > $O/k1d.raw.csv 2>/dev/null; rm -f $O/k1d.ncu-rep
cat $O/summary.txt; cat $O/plain.json | cut -c1-600; du -sh $O
