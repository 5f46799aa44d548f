#!/bin/bash
# Round 2, GPU call 1: sanity of the default build, measurement of the three blind round-1 variants, latency probe,
# full-horizon default-vs-strict parity.
set -u
mkdir -p gpurun_out/c1
O=gpurun_out/c1
PKG=sparse-optimal-control-of-viscous-chan-hilliard-via-gradient-descent--1d-2d_b200
cp $PKG/libvch_b200.so /tmp/lib_default.so
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/smi.txt 2>&1
./scripts/latency_probe > $O/latency_probe.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest_default.log 2>&1; echo "pytest default rc=$?" >> $O/summary.txt
for v in default v2 fs v2fs b6 b6fs; do
  if [ $v = default ]; then cp /tmp/lib_default.so $PKG/libvch_b200.so; else cp variants/lib_$v.so $PKG/libvch_b200.so; fi
  export VCH_BICG6=0; case $v in b6*) export VCH_BICG6=1;; esac
  if [ $v != default ]; then
    timeout 600 python -m pytest tests/test_gpu_2d.py tests/test_gpu_dropin_2d.py tests/test_gpu_edge_cases.py -m gpu -x -q > $O/pytest_$v.log 2>&1; echo "pytest $v rc=$?" >> $O/summary.txt
  fi
  timeout 900 python bench.py --horizon 300 --steps 2 --warmup 2 --no-e2e --no-cpu > $O/bench_$v.json 2> $O/bench_$v.err; echo "bench $v rc=$?" >> $O/summary.txt
done
unset VCH_BICG6
cp /tmp/lib_default.so $PKG/libvch_b200.so
timeout 900 python scripts/parity_vs_strict.py 1024 1000 > $O/parity_vs_strict_1024x1000.json 2> $O/parity_vs_strict.err; echo "parity_vs_strict rc=$?" >> $O/summary.txt
cat $O/summary.txt
