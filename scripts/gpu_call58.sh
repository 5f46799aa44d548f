#!/bin/bash
# 1D time-loop kernels: register cap -> CTAs per SM -> waves for the 1024-problem ensemble
set -u
O=gpurun_out/c58; mkdir -p $O
PKG="sparse-optimal-control-of-viscous-chan-hilliard-via-gradient-descent--1d-2d_b200"
for v in none 96 80 64 none 96; do
  cp scripts/libvar_$v.so $PKG/libvch_b200.so
  timeout 300 python bench.py --workload ensemble1d --steps 6 --warmup 3 > $O/ens_$v.json 2> $O/ens_$v.err
  python - <<PY
import json
try:
    d=json.loads(open("$O/ens_$v.json").read().strip().splitlines()[-1])
    print("regs $v:", round(d["value"]), d["unit"], "ms/step", round(d["ms_per_step"],3), {k: d["config"].get(k) for k in ("problems_per_gpu",)}, d.get("sum_J", d["config"].get("sum_J")))
except Exception as e:
    print("regs $v: ERR", e); print(open("$O/ens_$v.err").read()[-1500:])
PY
done
for v in 96 80; do
  cp scripts/libvar_$v.so $PKG/libvch_b200.so
  timeout 600 python -m pytest tests/test_gpu_1d.py tests/test_gpu_dropin_1d.py tests/test_gpu_edge_cases.py -m gpu -q > $O/pytest_$v.log 2>&1; echo "pytest $v rc=$?"; tail -2 $O/pytest_$v.log
done
