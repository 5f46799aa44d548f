#!/bin/bash
set -u
O=gpurun_out/c45; mkdir -p $O
B="python bench.py --horizon 300 --steps 2 --warmup 2 --no-e2e --no-cpu --no-parity --no-slab --no-ensemble --profile-steps 2"
for k in 1 2 3 1 2; do
  VCH_WHILE_UNROLL=$k timeout 600 $B > $O/u$k.json 2> $O/u$k.err
  python -c "
import json; d=json.loads(open('$O/u$k.json').read().strip().splitlines()[-1]); print('unroll $k: it/s', round(d['value'],4), 'ms/step', round(d['ms_per_step'],1), d['solver'])"
done
