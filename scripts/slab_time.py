"""Slab-mode timing: forward sweep of ONE N^2 problem on WORLD_SIZE GPUs (1 = the ordinary single-GPU context).
    python -m torch.distributed.run --nnodes=1 --nproc-per-node R --master-addr 127.0.0.1 --master-port 29535 scripts/slab_time.py N M [profile]
"""
import os, sys, time
import numpy as np, torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sparse-optimal-control-of-viscous-chan-hilliard-via-gradient-descent--1d-2d_b200"))
import vch_b200_native as nat
N = int(sys.argv[1]); M = int(sys.argv[2]); prof = len(sys.argv) > 3 and sys.argv[3] == "profile"
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
if world > 1: dist.init_process_group("gloo", rank=rank, world_size=world)
h = 1.0 / N
if world > 1: c = nat.SlabCtx2D.create_distributed(N, h, 1.0, 0.05, 10.0, 0.75, 1.0, 1e-4, 1e-2)
else: c = nat.Ctx2D(N, N, h, h, 1.0, 1.0, 0.05, 10.0, 0.75, 1.0, 1e-4, 1e-2); c.row0, c.rows = 0, N + 1
rng = np.random.default_rng(42)
phi0 = 0.1 * rng.standard_normal((N + 1, N + 1)); phi0 -= phi0.mean()
p0 = torch.from_numpy(np.ascontiguousarray(phi0[c.row0:c.row0 + c.rows])).cuda()
dts = np.full(M, 1e-2)
for rep in range(3):
    if prof and rep == 2: c.profile(True)
    if world > 1: dist.barrier()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    hist, _, _ = c.forward(p0, None, dts)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    if rank == 0:
        s = c.last_stats
        print(f"N={N} ranks={world} rep {rep}: {1e3*(t1-t0)/M:.3f} ms/step  newton evals {s['newton_residual_evals']} solves {s['newton_linear_solves']} "
              f"krylov its {s['krylov_iterations']} launches {s['kernel_launches']}", flush=True)
if prof and rank == 0:
    rep = c.profile_report(); tot = sum(v[0] for v in rep.values())
    for k, v in sorted(rep.items(), key=lambda kv: -kv[1][0]):
        print(f"  {k:26s} n={v[1]:6d} avg {1e3*v[0]/v[1]:8.2f} us  share {v[0]/tot:.3f}")
if world > 1: dist.barrier(); dist.destroy_process_group()
