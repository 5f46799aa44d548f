#!/bin/bash
set -u
O=gpurun_out/c22; mkdir -p $O
B="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-parity --no-slab --profile-steps 2"
timeout 600 $B > $O/new.json 2> $O/new.err
VCH_FUSED_SOLVE=0 timeout 600 $B > $O/nofuse.json 2> $O/nofuse.err
VCH_TILED=0 timeout 600 $B > $O/old.json 2> $O/old.err
VCH_COLS_TMA=0 timeout 600 $B > $O/notma.json 2> $O/notma.err
python - <<'PY'
import json
for f in ("new", "nofuse", "old", "notma"):
    try:
        d = json.loads(open(f"gpurun_out/c22/{f}.json").read().strip().splitlines()[-1])
        print(f, "it/s", round(d["value"], 4), "ms/step", round(d["ms_per_step"], 1), d["solver"], "launches", d["gpu_launches"])
    except Exception as e:
        print(f, "ERR", e); print(open(f"gpurun_out/c22/{f}.err").read()[-2000:])
PY
