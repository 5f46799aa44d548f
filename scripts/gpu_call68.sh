#!/bin/bash
set -u
O=gpurun_out/c68; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/summary.txt
timeout 300 python __graft_entry__.py --smoke > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/summary.txt
( time timeout 1700 python bench.py ) > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?" >> $O/summary.txt
cat $O/summary.txt; tail -3 $O/pytest.log; tail -2 $O/smoke.log | cut -c1-300
python - <<'PY'
import json
d = json.loads(open("gpurun_out/c68/bench_default.json").read().strip().splitlines()[-1])
print("it/s", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"], "parity", d["parity_vs_strict"]["pass"])
print("ens1d", d["ensemble1d"]["problem_it_per_s"], "roof", d["roofline"]["frac"], d["roofline"]["traffic"], "cpu same_config", d["cpu_baseline"]["same_config"])
PY
