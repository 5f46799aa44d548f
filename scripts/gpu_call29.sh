#!/bin/bash
set -u
N=${1:-2}
O=gpurun_out/c41; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_slab.py -m gpu -q > $O/pytest_slab_n$N.log 2>&1; echo "pytest slab rc=$?" >> $O/summary_n$N.txt
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 2 --warmup 1 --no-e2e --horizon 200 > $O/bench_n$N.json 2> $O/bench_n$N.err; echo "bench rc=$?" >> $O/summary_n$N.txt
cat $O/summary_n$N.txt; tail -4 $O/pytest_slab_n$N.log
python - $N <<'PY'
import json, sys
N=sys.argv[1]
try:
    d=json.loads(open(f'gpurun_out/c41/bench_n{N}.json').read().strip().splitlines()[-1])
    print("value", d['value'], "ms/step", d['ms_per_step'])
    print(json.dumps(d['slab_4096'], indent=1)); print(d.get('ensemble1d'))
except Exception as e:
    print("ERR", e); print(open(f'gpurun_out/c41/bench_n{N}.err').read()[-3000:])
PY
