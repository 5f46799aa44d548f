#!/bin/bash
set -u
mkdir -p gpurun_out/c13
O=gpurun_out/c13
nvidia-smi -L > $O/gpus.txt
timeout 1200 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/summary.txt
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 1 --no-e2e --horizon 200 > $O/bench_n2.json 2> $O/bench_n2.err; echo "bench n2 rc=$?" >> $O/summary.txt
cat $O/summary.txt; tail -8 $O/pytest.log
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/c13/bench_n2.json').read().strip().splitlines()[-1])
    print("N=2 value", d['value'], "ms/step", d['ms_per_step'])
    print(json.dumps(d['slab_4096'], indent=1))
except Exception as e:
    print("ERR", e); print(open('gpurun_out/c13/bench_n2.err').read()[-3000:])
PY
