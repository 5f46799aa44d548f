"""World-size-2 CPU test (gloo) of the N>1 host logic: the ensemble/benchmark path shards INDEPENDENT control problems
across ranks with no data-path collective; only scalars (J per problem, timings) are gathered.  The compute itself
needs a GPU, so here each rank stands in the oracle for its shard (tests may use the oracle) and the gathered result is
compared with the single-rank run."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import PKG, ROOT


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    for p in (PKG, os.path.join(PKG, "Vch_control_1D"), os.path.join(ROOT, "oracle")):
        sys.path.insert(0, p)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import GD_1D as G
    import vch_oracle as O
    B = 5
    ens = G.make_ensemble(B, fwd_config=G.ForwardSolverConfig(N=32, T=0.05))
    lo, hi = G.shard_range(B, rank, world)
    P = O.Phys1D(N=32, T=0.05)
    J_local = torch.zeros(B, dtype=torch.float64)
    for b in range(lo, hi):                      # this rank's shard only; no exchange of fields
        fw = O.forward_1d(P, None, ens["phi_init"][b])
        J, _ = O.cost_1d(fw["phi"], np.zeros_like(fw["phi"]), ens["phi_Q"][b], ens["phi_T"][b], ens["x"], ens["t_hist"],
                         ens["b1"][b], ens["b2"][b], ens["b3"][b], ens["ksp"][b])
        J_local[b] = J
    dist.all_reduce(J_local, op=dist.ReduceOp.SUM)      # scalar gather
    t = torch.tensor([1.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)            # bench.py's max-over-ranks timing rule
    if rank == 0:
        np.save(os.path.join(out_dir, "J.npy"), J_local.numpy())
        np.save(os.path.join(out_dir, "t.npy"), t.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_equals_single_rank(tmp_path):
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    J2 = np.load(tmp_path / "J.npy")
    assert np.load(tmp_path / "t.npy")[0] == 2.0
    for p in (PKG, os.path.join(PKG, "Vch_control_1D"), os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from conftest import load_dropin
    G = load_dropin("1D")["GD_1D"]
    import vch_oracle as O
    ens = G.make_ensemble(5, fwd_config=G.ForwardSolverConfig(N=32, T=0.05))
    P = O.Phys1D(N=32, T=0.05)
    for b in range(5):
        fw = O.forward_1d(P, None, ens["phi_init"][b])
        J, _ = O.cost_1d(fw["phi"], np.zeros_like(fw["phi"]), ens["phi_Q"][b], ens["phi_T"][b], ens["x"], ens["t_hist"],
                         ens["b1"][b], ens["b2"][b], ens["b3"][b], ens["ksp"][b])
        assert J2[b] == J
