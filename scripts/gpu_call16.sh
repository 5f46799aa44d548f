#!/bin/bash
# tiled stencil kernels + fused solve start/close: microbench, parity suite, A/B bench at M = 300
set -u
O=gpurun_out/c16; mkdir -p $O
./scripts/kernel_bench > $O/kernel_bench.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/summary.txt
B="python bench.py --horizon 300 --steps 2 --warmup 2 --no-e2e --no-cpu --no-parity --no-slab"
timeout 600 $B > $O/bench_h300.json 2> $O/bench_h300.err; echo "bench rc=$?" >> $O/summary.txt
VCH_FUSED_SOLVE=0 timeout 600 $B > $O/bench_h300_nofuse.json 2> $O/bench_h300_nofuse.err; echo "bench nofuse rc=$?" >> $O/summary.txt
VCH_TILED=0 timeout 600 $B > $O/bench_h300_old.json 2> $O/bench_h300_old.err; echo "bench old rc=$?" >> $O/summary.txt
cat $O/kernel_bench.txt $O/summary.txt; tail -5 $O/pytest.log
python - <<'PY'
import json
for f in ("bench_h300", "bench_h300_nofuse", "bench_h300_old"):
    try:
        d = json.loads(open(f"gpurun_out/c16/{f}.json").read().strip().splitlines()[-1])
        print(f, "it/s", round(d["value"], 4), "ms/step", round(d["ms_per_step"], 1), d["solver"], "launches", d["gpu_launches"])
        if f == "bench_h300":
            for k, v in d["roofline"]["kernels"].items(): print("   ", k, v)
    except Exception as e:
        print(f, "ERR", e); print(open(f"gpurun_out/c16/{f}.err").read()[-2000:])
PY
