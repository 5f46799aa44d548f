import os, sys, time
import numpy as np, torch
ROOT = "/root/repo"
PKG = os.path.join(ROOT, "sparse-optimal-control-of-viscous-chan-hilliard-via-gradient-descent--1d-2d_b200")
sys.path.insert(0, PKG); sys.path.insert(0, os.path.join(PKG, "Vch_control_1D"))
import vch_b200_native as nat
import GD_1D as G
B = 1024
cfg = G.ForwardSolverConfig()
ens = G.make_ensemble(B)
c = nat.ctx1d(cfg.N, cfg.Lx / cfg.N, cfg.Lx, cfg.tau, cfg.gamma, cfg.c1, cfg.c2, cfg.kappa)
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
phi_init, phiQ, phiT = dev(ens["phi_init"]), dev(ens["phi_Q"]), dev(ens["phi_T"])
hist, _, _ = c.forward(phi_init, None, ens["dts"])
u = torch.zeros_like(hist)
for it in range(14):
    ts = []
    torch.cuda.synchronize(); t0 = time.perf_counter()
    _, _, r = c.adjoint(hist, ens["t_hist"], ens["b1"], ens["b2"], phiQ, phiT); torch.cuda.synchronize(); ts.append(time.perf_counter())
    u1, red = c.grad_prox(u, r, ens["b3"], 100.0, ens["ksp"], -1.0, 1.0); torch.cuda.synchronize(); ts.append(time.perf_counter())
    hist1, _, _ = c.forward(phi_init, u1, ens["dts"]); torch.cuda.synchronize(); ts.append(time.perf_counter())
    J = c.cost(hist1, u1, phiQ, phiT, ens["x"], ens["t_hist"], ens["b1"], ens["b2"], ens["b3"], ens["ksp"]); torch.cuda.synchronize(); ts.append(time.perf_counter())
    d = np.diff([t0] + ts) * 1e3
    print(f"iter {it}: adjoint {d[0]:.2f} prox {d[1]:.2f} forward {d[2]:.2f} cost {d[3]:.2f} ms | sumJ {float(J[:,0].sum()):.4f} max|u| {float(u1.abs().max()):.3f} max|phi| {float(hist1.abs().max()):.4f}", flush=True)
    u, hist = u1, hist1
