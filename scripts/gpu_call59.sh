#!/bin/bash
# 1D time-loop kernels: threads per CTA x register cap
set -u
O=gpurun_out/c59; mkdir -p $O
PKG="sparse-optimal-control-of-viscous-chan-hilliard-via-gradient-descent--1d-2d_b200"
for cfg in "96 0" "96 96" "none 96" "80 96" "96 128" "96 64" "80 64" "96 96"; do
  set -- $cfg; v=$1; t=$2
  cp scripts/libvar_$v.so $PKG/libvch_b200.so
  VCH_1D_THREADS=$t timeout 300 python bench.py --workload ensemble1d --steps 6 --warmup 3 > $O/ens_${v}_$t.json 2> $O/ens_${v}_$t.err
  python - <<PY
import json
try:
    d=json.loads(open("$O/ens_${v}_$t.json").read().strip().splitlines()[-1])
    print("regs $v threads $t:", round(d["value"]), d["unit"], "ms/step", round(d["ms_per_step"],3), d.get("sum_J", d["config"].get("sum_J")))
except Exception as e:
    print("regs $v threads $t: ERR", e); print(open("$O/ens_${v}_$t.err").read()[-1500:])
PY
done
cp scripts/libvar_96.so $PKG/libvch_b200.so
VCH_1D_THREADS=96 timeout 600 python -m pytest tests/test_gpu_1d.py tests/test_gpu_dropin_1d.py tests/test_gpu_edge_cases.py -m gpu -q > $O/pytest_96_96.log 2>&1; echo "pytest regs 96 threads 96 rc=$?"; tail -2 $O/pytest_96_96.log
