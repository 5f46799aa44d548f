#!/bin/bash
set -u
N=${1:-2}
O=gpurun_out/c40; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_slab.py -m gpu -q > $O/pytest_slab_n$N.log 2>&1; echo "pytest slab (chain) rc=$?" >> $O/summary_n$N.txt
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --workload slab2d --grid 4096 --horizon 20 --steps 2 --warmup 1"
timeout 600 $T > $O/slab_chain_n$N.json 2> $O/slab_chain_n$N.err; echo "slab chain rc=$?" >> $O/summary_n$N.txt
VCH_SLAB_CHAIN=0 timeout 600 $T > $O/slab_barrier_n$N.json 2> $O/slab_barrier_n$N.err; echo "slab barrier rc=$?" >> $O/summary_n$N.txt
cat $O/summary_n$N.txt; tail -3 $O/pytest_slab_n$N.log
for f in slab_chain slab_barrier; do echo "== $f"; tail -1 $O/${f}_n$N.json | cut -c1-900; tail -3 $O/${f}_n$N.err | cut -c1-300; done
