#!/bin/bash
set -u
O=gpurun_out/c43; mkdir -p $O
B="python bench.py --horizon 300 --steps 2 --warmup 2 --no-e2e --no-cpu --no-parity --no-slab --no-ensemble --profile-steps 2"
VCH_DEBUG=1 timeout 300 python bench.py --horizon 3 --steps 1 --warmup 1 --no-e2e --no-cpu --no-parity --no-slab --no-ensemble --profile-steps 1 2>&1 >/dev/null | grep "L2 window" | head -2
timeout 600 $B > $O/persist.json 2> $O/persist.err; echo "rc=$?"
VCH_L2_PERSIST=0 timeout 600 $B > $O/nopersist.json 2> $O/nopersist.err; echo "rc=$?"
timeout 600 $B > $O/persist2.json 2> $O/persist2.err; echo "rc=$?"
python - <<'PY'
import json
for f in ("persist", "nopersist", "persist2"):
    try:
        d = json.loads(open(f"gpurun_out/c43/{f}.json").read().strip().splitlines()[-1])
        print(f, "it/s", round(d["value"], 4), "ms/step", round(d["ms_per_step"], 1), d["solver"]["J_last"])
    except Exception as e:
        print(f, "ERR", e); print(open(f"gpurun_out/c43/{f}.err").read()[-1500:])
PY
