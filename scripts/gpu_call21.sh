#!/bin/bash
set -u
O=gpurun_out/c21; mkdir -p $O
( time timeout 1700 python bench.py ) > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?" >> $O/summary.txt
cat $O/summary.txt; tail -5 $O/bench_default.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/c21/bench_default.json").read().strip().splitlines()[-1])
print("it/s", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"] if d.get("e2e") else None, d["solver"], "launches", d["gpu_launches"])
print("parity", d.get("parity_vs_strict"))
print("cpu", {k: v for k, v in (d.get("cpu_baseline") or {}).items() if k != "detail"})
print("roof", {k: v for k, v in d["roofline"].items() if k != "kernels"})
PY
