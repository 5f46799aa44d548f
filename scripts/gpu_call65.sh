#!/bin/bash
# the driver's exact N = 2 command (defaults: full horizon, e2e leg, slab leg, ensemble leg)
set -u
O=gpurun_out/c65; mkdir -p $O
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 ) > $O/bench_n2_default.json 2> $O/bench_n2_default.err; echo "rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/c65/bench_n2_default.json').read().strip().splitlines()[-1])
    print("N=2 value", d['value'], "ms/step", d['ms_per_step'], "e2e", d['e2e'].get('value'), d['e2e'].get('error'))
    s=d['slab_4096']; print({k: s.get(k) for k in ('speedup_vs_1gpu','efficiency','J_rel_diff_vs_1gpu','error')})
    print(d['ensemble1d']['problem_it_per_s'], d.get('concurrent_problems'), d['clocks'])
except Exception as e:
    print("ERR", e)
PY
tail -5 $O/bench_n2_default.err
