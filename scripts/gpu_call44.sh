#!/bin/bash
set -u
O=gpurun_out/c44; mkdir -p $O
B="python bench.py --horizon 300 --steps 2 --warmup 2 --no-e2e --no-cpu --no-parity --no-slab --no-ensemble --profile-steps 2"
for k in 0 2 3 5 0; do
  VCH_L2_PERSIST=$k timeout 600 $B > $O/p$k.json 2> $O/p$k.err
  python -c "
import json; d=json.loads(open('$O/p$k.json').read().strip().splitlines()[-1]); print('persist vecs $k: it/s', round(d['value'],4), 'ms/step', round(d['ms_per_step'],1))"
done
