#!/bin/bash
set -u
O=gpurun_out/c18; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/summary.txt
B="python bench.py --horizon 300 --steps 2 --warmup 2 --no-e2e --no-cpu --no-parity --no-slab"
timeout 600 $B > $O/bench_h300.json 2> $O/bench_h300.err; echo "bench rc=$?" >> $O/summary.txt
VCH_COLS_TMA=0 timeout 600 $B > $O/bench_h300_notma.json 2> $O/bench_h300_notma.err; echo "bench notma rc=$?" >> $O/summary.txt
cat $O/summary.txt; tail -5 $O/pytest.log
python - <<'PY'
import json
for f in ("bench_h300", "bench_h300_notma"):
    try:
        d = json.loads(open(f"gpurun_out/c18/{f}.json").read().strip().splitlines()[-1])
        print(f, "it/s", round(d["value"], 4), "ms/step", round(d["ms_per_step"], 1), d["solver"], "launches", d["gpu_launches"])
        for k, v in list(d["roofline"]["kernels"].items())[:8]: print("   ", k, v)
    except Exception as e:
        print(f, "ERR", e); print(open(f"gpurun_out/c18/{f}.err").read()[-2000:])
PY
