#!/bin/bash
# 1D kernels with the two-barrier block reduction
set -u
O=gpurun_out/c67; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_1d.py tests/test_gpu_dropin_1d.py tests/test_gpu_edge_cases.py -m gpu -q > $O/pytest_1d.log 2>&1; echo "pytest 1d rc=$?"; tail -2 $O/pytest_1d.log
for k in 1 2; do
timeout 300 python bench.py --workload ensemble1d --steps 6 --warmup 3 > $O/ens_$k.json 2> $O/ens_$k.err
python - <<PY
import json
d=json.loads(open("$O/ens_$k.json").read().strip().splitlines()[-1])
print("ensemble1d:", round(d["value"]), d["unit"], "ms/step", round(d["ms_per_step"],3), d.get("sum_J", d["config"].get("sum_J")))
PY
done
