#!/bin/bash
set -u
O=gpurun_out/c28; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/summary.txt
( time timeout 1700 python bench.py ) > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?" >> $O/summary.txt
cat $O/summary.txt; tail -4 $O/pytest.log; tail -5 $O/bench_default.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/c28/bench_default.json").read().strip().splitlines()[-1])
print("it/s", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"] if d.get("e2e") else None, d["solver"], "launches", d["gpu_launches"])
print("parity", d.get("parity_vs_strict", {}).get("pass"), "ens1d", d.get("ensemble1d"))
print("roof", {k: v for k, v in d["roofline"].items() if k != "kernels"})
print("cost_kernel", d["roofline"]["kernels"].get("cost_kernel"), "slab", d.get("slab_4096"))
PY
