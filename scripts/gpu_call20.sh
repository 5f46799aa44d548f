#!/bin/bash
set -u
O=gpurun_out/c20; mkdir -p $O
./scripts/kernel_bench > $O/kernel_bench.txt 2>&1
for k in residual_tile_kernel dmu_close_tile_kernel step_setup_tile_kernel adj_rhs_tile_kernel adj_qr_tile_kernel; do
  timeout 200 ncu --set full --clock-control none -k $k -s 25 -c 2 -o $O/$k -f ./scripts/kernel_bench > $O/ncu_$k.log 2>&1
  ncu -i $O/$k.ncu-rep --page raw --csv > $O/$k.raw.csv 2>/dev/null
  rm -f $O/$k.ncu-rep
done
timeout 600 python -m pytest tests -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/summary.txt
cat $O/kernel_bench.txt $O/summary.txt; tail -3 $O/pytest.log; du -sh $O
