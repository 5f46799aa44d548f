#!/bin/bash
# N = 2 replicas: where does the loss against one GPU come from?  per-rank times with distinct and with identical problems
set -u
O=gpurun_out/c66; mkdir -p $O
B="bench.py --gpus 2 --horizon 300 --steps 3 --warmup 3 --no-e2e --no-cpu --no-parity --no-slab --no-ensemble --profile-steps 2"
for mode in distinct same; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 $B --rank-seeds $mode > $O/n2_$mode.json 2> $O/n2_$mode.err; echo "$mode rc=$?"
done
timeout 300 python bench.py --gpus 1 --horizon 300 --steps 3 --warmup 3 --no-e2e --no-cpu --no-parity --no-slab --no-ensemble --no-concurrent --profile-steps 2 > $O/n1.json 2> $O/n1.err; echo "n1 rc=$?"
CUDA_VISIBLE_DEVICES=1 timeout 300 python bench.py --gpus 1 --horizon 300 --steps 3 --warmup 3 --no-e2e --no-cpu --no-parity --no-slab --no-ensemble --no-concurrent --profile-steps 2 > $O/n1_gpu1.json 2> $O/n1_gpu1.err; echo "n1 on GPU 1 rc=$?"
python - <<'PY'
import json
for f in ("n2_distinct","n2_same","n1","n1_gpu1"):
    try:
        d=json.loads(open(f"gpurun_out/c66/{f}.json").read().strip().splitlines()[-1])
        print(f, "value", round(d["value"],4), "ms/step", round(d["ms_per_step"],1), d.get("per_rank"), d["clocks"])
    except Exception as e:
        print(f, "ERR", e)
PY
