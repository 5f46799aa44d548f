#!/bin/bash
# r02 evidence: ncu launch list + one --set full capture of the hot kernels (current build), after the plain command exited 0
set -u
O=gpurun_out/c15; mkdir -p $O
CMD="python bench.py --horizon 10 --steps 1 --warmup 1 --no-e2e --no-cpu --no-parity --no-slab --profile-steps 5"
VCH_NO_GRAPHS=1 timeout 300 $CMD > $O/plain.json 2> $O/plain.err; echo "plain rc=$?" >> $O/summary.txt
VCH_NO_GRAPHS=1 timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/launches.csv $CMD > $O/ncu_list.log 2>&1; echo "ncu list rc=$?" >> $O/summary.txt
VCH_NO_GRAPHS=1 timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:rows16|cols16|residual_kernel|dmu_ceiling|schur_rhs|bicg_init|bicg_close|step_setup|adj_rhs|clip_mass|mass_shift' -s 400 -c 60 -o $O/full_r02 -f $CMD > $O/ncu_full.log 2>&1; echo "ncu full rc=$?" >> $O/summary.txt
./scripts/kernel_bench > $O/kernel_bench.txt 2>&1
cat $O/summary.txt; tail -3 $O/ncu_full.log; ls -la $O
