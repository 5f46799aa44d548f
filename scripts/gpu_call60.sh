#!/bin/bash
set -u
O=gpurun_out/c60; mkdir -p $O
PKG="sparse-optimal-control-of-viscous-chan-hilliard-via-gradient-descent--1d-2d_b200"
cp $PKG/libvch_b200.so scripts/libvar_intree.so
for v in intree 80 72 64 80 72; do
  cp scripts/libvar_$v.so $PKG/libvch_b200.so
  timeout 300 python bench.py --workload ensemble1d --steps 6 --warmup 3 > $O/ens_$v.json 2> $O/ens_$v.err
  python - <<PY
import json
try:
    d=json.loads(open("$O/ens_$v.json").read().strip().splitlines()[-1])
    print("regs $v threads 96:", round(d["value"]), d["unit"], "ms/step", round(d["ms_per_step"],3), d.get("sum_J", d["config"].get("sum_J")))
except Exception as e:
    print("regs $v: ERR", e); print(open("$O/ens_$v.err").read()[-1500:])
PY
done
