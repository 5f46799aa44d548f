#!/bin/bash
set -u
O=gpurun_out/c19; mkdir -p $O
./scripts/fft_scaling | tail -24 > $O/fft_scaling_rot.txt 2>&1
./scripts/fft_scaling_norot | tail -24 > $O/fft_scaling_norot.txt 2>&1
timeout 300 ncu --set full --clock-control none -k regex:cols16 -s 40 -c 6 -o $O/cols16 -f ./scripts/fft_scaling > $O/ncu_cols.log 2>&1
timeout 300 ncu --set full --clock-control none -k 'regex:residual_tile|dmu_close|residual_kernel|dmu_ceiling|step_setup_tile|adj_rhs_tile' -s 12 -c 12 -o $O/tiles -f ./scripts/kernel_bench > $O/ncu_tiles.log 2>&1
for r in cols16 tiles; do ncu -i $O/$r.ncu-rep --page raw --csv > $O/$r.raw.csv 2>/dev/null; done
ls -la $O; du -sh $O
cat $O/fft_scaling_rot.txt; echo ----; cat $O/fft_scaling_norot.txt
