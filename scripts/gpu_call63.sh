#!/bin/bash
# ncu --set full of the 1D time-loop kernels after the occupancy change
set -u
O=gpurun_out/c63; mkdir -p $O
CMD="python bench.py --workload ensemble1d --steps 2 --warmup 1"
timeout 200 $CMD > $O/plain.json 2> $O/plain.err; echo "plain rc=$?"
timeout 500 ncu --set full --clock-control none -k regex:1d -s 4 -c 8 -o $O/full_1d -f $CMD > $O/ncu.log 2>&1; echo "ncu rc=$?"
ncu -i $O/full_1d.ncu-rep --page raw --csv > $O/raw.csv 2>/dev/null
python scripts/ncu_full_summarise.py $O/raw.csv "ncu --set full --clock-control none -k regex:1d -s 4 -c 8 $CMD (1024 problems, N = 128, 100 steps; build with 96 threads / 80 registers per CTA)" $O/r02_ncu_full_1d_after.txt $O/traffic_1d.json
rm -f $O/full_1d.ncu-rep
head -60 $O/r02_ncu_full_1d_after.txt
