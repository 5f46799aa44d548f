#!/bin/bash
# 4 GPUs: all slab tests (incl. the 4-rank one) + the N = 4 arm of the default bench on the final build
set -u
O=gpurun_out/c57; mkdir -p $O
nvidia-smi -L > $O/gpus.txt
timeout 600 python -m pytest tests/test_gpu_slab.py -m gpu -q > $O/pytest_slab.log 2>&1; echo "pytest slab rc=$?" >> $O/summary.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 2 --warmup 3 --no-e2e --horizon 100 > $O/bench_n4.json 2> $O/bench_n4.err; echo "bench n4 rc=$?" >> $O/summary.txt
cat $O/summary.txt; tail -4 $O/pytest_slab.log
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/c57/bench_n4.json').read().strip().splitlines()[-1])
    print("N=4 value", d['value'], "ms/step", d['ms_per_step'])
    s=d['slab_4096']; print({k: s[k] for k in ('speedup_vs_1gpu','efficiency','J_rel_diff_vs_1gpu')}, s['slab'], s['one_gpu'])
    print(d.get('ensemble1d'))
except Exception as e:
    print("ERR", e); print(open('gpurun_out/c57/bench_n4.err').read()[-3000:])
PY
