#!/bin/bash
# r02 evidence for the final build: ncu launch list + --set full captures (exported to CSV on the box; the .ncu-rep files stay there)
set -u
O=gpurun_out/c24; mkdir -p $O
CMD="python bench.py --horizon 10 --steps 1 --warmup 1 --no-e2e --no-cpu --no-parity --no-slab --profile-steps 5"
VCH_NO_GRAPHS=1 timeout 300 $CMD > $O/plain.json 2> $O/plain.err; echo "plain rc=$?" >> $O/summary.txt
VCH_NO_GRAPHS=1 timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/launches.csv $CMD > $O/ncu_list.log 2>&1; echo "ncu list rc=$?" >> $O/summary.txt
VCH_NO_GRAPHS=1 timeout 420 ncu --set full --clock-control none -k 'regex:cols16|rows16' -s 300 -c 12 -o $O/fft -f $CMD > $O/ncu_fft.log 2>&1; echo "ncu fft rc=$?" >> $O/summary.txt
VCH_NO_GRAPHS=1 timeout 420 ncu --set full --clock-control none -k 'regex:tile_kernel|clip_mass|mass_shift' -s 20 -c 14 -o $O/tiles -f $CMD > $O/ncu_tiles.log 2>&1; echo "ncu tiles rc=$?" >> $O/summary.txt
for r in fft tiles; do ncu -i $O/$r.ncu-rep --page raw --csv > $O/$r.raw.csv 2>/dev/null; rm -f $O/$r.ncu-rep; done
cat $O/summary.txt; du -sh $O
