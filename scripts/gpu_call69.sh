#!/bin/bash
# ncu evidence of the FINAL build: launch list + --set full of the transform kernels (same commands as the earlier round-2 captures)
set -u
O=gpurun_out/c69; mkdir -p $O
CMD="python bench.py --horizon 10 --steps 1 --warmup 1 --no-e2e --no-cpu --no-parity --no-slab --no-ensemble --no-concurrent --profile-steps 5"
VCH_NO_GRAPHS=1 timeout 200 $CMD > $O/plain.json 2> $O/plain.err; echo "plain rc=$?"
VCH_NO_GRAPHS=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/launches.csv $CMD > $O/ncu_list.log 2>&1; echo "list rc=$?"
python scripts/ncu_summarise.py $O/launches.csv "VCH_NO_GRAPHS=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv $CMD (final build)" > $O/r02_ncu_launch_list_final.txt
VCH_NO_GRAPHS=1 timeout 300 ncu --set full --clock-control none -k 'regex:cols16|rows16' -s 300 -c 12 -o $O/fft -f $CMD > $O/ncu_full.log 2>&1; echo "full rc=$?"
ncu -i $O/fft.ncu-rep --page raw --csv > $O/fft.raw.csv 2>/dev/null; rm -f $O/fft.ncu-rep
python scripts/ncu_full_summarise.py $O/fft.raw.csv "VCH_NO_GRAPHS=1 ncu --set full --clock-control none -k regex:cols16|rows16 -s 300 -c 12 $CMD (final build)" $O/r02_ncu_full_fft_final.txt $O/traffic_fft.json
head -12 $O/r02_ncu_launch_list_final.txt; grep -E "^==|time_duration|dram__bytes" $O/r02_ncu_full_fft_final.txt | head -30
