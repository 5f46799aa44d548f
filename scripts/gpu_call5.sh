#!/bin/bash
set -u
mkdir -p gpurun_out/c5
O=gpurun_out/c5
./scripts/fft_scaling > $O/fft_scaling.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/summary.txt
timeout 900 python bench.py --horizon 300 --steps 2 --warmup 2 --no-e2e --no-cpu > $O/bench_h300.json 2> $O/bench_h300.err; echo "bench rc=$?" >> $O/summary.txt
VCH_FFT16=0 timeout 900 python bench.py --horizon 300 --steps 2 --warmup 2 --no-e2e --no-cpu > $O/bench_h300_old.json 2> $O/bench_h300_old.err; echo "bench old rc=$?" >> $O/summary.txt
cat $O/fft_scaling.txt $O/summary.txt; tail -3 $O/pytest.log
