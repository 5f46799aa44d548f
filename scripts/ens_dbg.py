import os, sys, time
import numpy as np, torch, torch.distributed as dist
ROOT = "/root/repo"
PKG = os.path.join(ROOT, "sparse-optimal-control-of-viscous-chan-hilliard-via-gradient-descent--1d-2d_b200")
sys.path.insert(0, PKG); sys.path.insert(0, os.path.join(PKG, "Vch_control_1D"))
import vch_b200_native as nat, GD_1D as G
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
use_dist = "--nodist" not in sys.argv
if world > 1 and use_dist:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
B = 512
cfg = G.ForwardSolverConfig(); ens = G.make_ensemble(B)
ctx = nat.Ctx1D(cfg.N, cfg.Lx / cfg.N, cfg.Lx, cfg.tau, cfg.gamma, cfg.c1, cfg.c2, cfg.kappa, device=local)
dev = torch.device("cuda", local)
up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
phi_init, phiQ, phiT = up(ens["phi_init"]), up(ens["phi_Q"]), up(ens["phi_T"])
hist, _, _ = ctx.forward(phi_init, None, ens["dts"])
u = torch.zeros_like(hist)
for rep in range(6):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    _, _, r = ctx.adjoint(hist, ens["t_hist"], ens["b1"], ens["b2"], phiQ, phiT); torch.cuda.synchronize(); t1 = time.perf_counter()
    u1, red = ctx.grad_prox(u, r, ens["b3"], 100.0, ens["ksp"], -1.0, 1.0); torch.cuda.synchronize(); t2 = time.perf_counter()
    h1, _, _ = ctx.forward(phi_init, u1, ens["dts"]); torch.cuda.synchronize(); t3 = time.perf_counter()
    J = ctx.cost(h1, u1, phiQ, phiT, ens["x"], ens["t_hist"], ens["b1"], ens["b2"], ens["b3"], ens["ksp"]); torch.cuda.synchronize(); t4 = time.perf_counter()
    print(f"rank {rank} rep {rep}: adjoint {1e3*(t1-t0):.1f} prox {1e3*(t2-t1):.1f} forward {1e3*(t3-t2):.1f} cost {1e3*(t4-t3):.1f} ms", flush=True)
if world > 1 and use_dist:
    dist.destroy_process_group()
