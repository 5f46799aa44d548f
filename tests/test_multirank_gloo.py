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


# ---------------------------------------------------------------------------------------------- slab decomposition (config 5)
def _slab_worker(rank, world, port, out_dir):
    """Host-side model of the slab mode's data movement on CPU/gloo: the row partition rule, the ghost-row exchange the
    stencils rely on (mirror rule only at the global boundary) and the chunking of the row<->column transposes."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    for p in (PKG, os.path.join(ROOT, "oracle")):
        sys.path.insert(0, p)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import vch_b200_native as nat
    import vch_oracle as O
    N = 64
    parts = nat.slab_partition(N, world)
    r0, nr = parts[rank]
    rng = np.random.default_rng(7)
    v = rng.standard_normal((N + 1, N + 1))                 # same global field on every rank
    mine = torch.from_numpy(v[r0:r0 + nr].copy())
    # ghost rows: neighbour's boundary row, or the mirror image at the global boundary
    lo = torch.empty(N + 1, dtype=torch.float64); hi = torch.empty(N + 1, dtype=torch.float64)
    reqs = []
    if rank > 0: reqs += [dist.isend(mine[0].clone(), rank - 1), dist.irecv(lo, rank - 1)]
    else: lo = mine[1].clone()
    if rank < world - 1: reqs += [dist.isend(mine[-1].clone(), rank + 1), dist.irecv(hi, rank + 1)]
    else: hi = mine[-2].clone()
    for q in reqs: q.wait()
    ext = torch.cat([lo[None], mine, hi[None]]).numpy()
    h = 1.0 / N
    colp = np.pad(ext, ((0, 0), (1, 1)), mode="reflect")
    lap = (ext[2:] + ext[:-2] - 2 * ext[1:-1]) / h**2 + (colp[1:-1, 2:] + colp[1:-1, :-2] - 2 * ext[1:-1]) / h**2
    L = O.neumann_2d(N, N, h, h)
    ref = (L @ v.ravel()).reshape(N + 1, N + 1)[r0:r0 + nr]
    ok_lap = np.allclose(lap, ref, rtol=1e-12, atol=1e-9)
    # transposes: rank r receives columns [c0, c0+nc) of every rank's rows -> (all rows, owned columns)
    send = [torch.from_numpy(np.ascontiguousarray(v[r0:r0 + nr, c0:c0 + nc])) for c0, nc in parts]
    recv = [torch.empty(parts[r][1], parts[rank][1], dtype=torch.float64) for r in range(world)]
    ops = [dist.P2POp(dist.isend, send[r], r) for r in range(world) if r != rank] + \
          [dist.P2POp(dist.irecv, recv[r], r) for r in range(world) if r != rank]
    recv[rank] = send[rank]
    for q in dist.batch_isend_irecv(ops): q.wait()
    cols = torch.cat(recv).numpy()
    ok_tr = np.array_equal(cols, v[:, r0:r0 + nr])
    flag = torch.tensor([float(ok_lap), float(ok_tr)], dtype=torch.float64)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        np.save(os.path.join(out_dir, "slab_ok.npy"), flag.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_slab_partition_halo_and_transpose_model(tmp_path):
    import vch_b200_native as nat
    assert nat.slab_partition(4096, 8) == [(512 * r, 512 + (r == 7)) for r in range(8)]      # 7 x 512 + 513
    assert nat.slab_partition(128, 1) == [(0, 129)]
    for bad in ((100, 2), (128, 3), (64, 16), (32, 8)):
        with pytest.raises(ValueError):
            nat.slab_partition(*bad)
    world = 2
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_slab_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert np.load(tmp_path / "slab_ok.npy").tolist() == [1.0, 1.0]
