#!/bin/bash
# 2 GPUs: slab tests + the N = 2 arm of the default bench (slab 4096^2 leg) on the build with the reproducibility fix
set -u
O=gpurun_out/c56; mkdir -p $O
nvidia-smi -L > $O/gpus.txt
timeout 600 python -m pytest tests/test_gpu_slab.py -m gpu -q > $O/pytest_slab.log 2>&1; echo "pytest slab rc=$?" >> $O/summary.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 --no-e2e --horizon 100 > $O/bench_n2.json 2> $O/bench_n2.err; echo "bench n2 rc=$?" >> $O/summary.txt
cat $O/summary.txt; tail -4 $O/pytest_slab.log
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/c56/bench_n2.json').read().strip().splitlines()[-1])
    print("N=2 value", d['value'], "ms/step", d['ms_per_step'])
    print(json.dumps(d['slab_4096'], indent=1)[:1500])
    print(d.get('ensemble1d'))
except Exception as e:
    print("ERR", e); print(open('gpurun_out/c56/bench_n2.err').read()[-3000:])
PY
