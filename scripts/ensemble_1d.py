"""Scratch: BASELINE config 4 — ensemble of 1024 independent 1D control problems, one optimistic PGD iteration."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "sparse-optimal-control-of-viscous-chan-hilliard-via-gradient-descent--1d-2d_b200")
sys.path.insert(0, PKG); sys.path.insert(0, os.path.join(PKG, "Vch_control_1D"))
import vch_b200_native as nat
import GD_1D as G
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
cfg = G.ForwardSolverConfig()
ens = G.make_ensemble(B)
c = nat.ctx1d(cfg.N, cfg.Lx / cfg.N, cfg.Lx, cfg.tau, cfg.gamma, cfg.c1, cfg.c2, cfg.kappa)
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
phi_init, phiQ, phiT = dev(ens["phi_init"]), dev(ens["phi_Q"]), dev(ens["phi_T"])
hist, _, _ = c.forward(phi_init, None, ens["dts"])
u = torch.zeros_like(hist)
for rep in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    u1, hist1, J, red, r = G.optimistic_iteration_ensemble(c, c, u, hist, phiQ, phiT, ens["x"], ens["t_hist"], ens["dts"], phi_init,
                                                           ens["b1"], ens["b2"], ens["b3"], ens["ksp"], 100.0)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"ensemble B={B}: one PGD iteration for all problems {dt*1e3:.2f} ms -> {B/dt:.0f} problem-iterations/s; J[0]={J[0,0]:.6f}", flush=True)
for name, f in (("adjoint", lambda: c.adjoint(hist, ens["t_hist"], ens["b1"], ens["b2"], phiQ, phiT)),
                ("forward", lambda: c.forward(phi_init, u1, ens["dts"])),
                ("cost", lambda: c.cost(hist1, u1, phiQ, phiT, ens["x"], ens["t_hist"], ens["b1"], ens["b2"], ens["b3"], ens["ksp"])),
                ("prox", lambda: c.grad_prox(u, r, ens["b3"], 100.0, ens["ksp"], -1.0, 1.0))):
    f(); torch.cuda.synchronize(); t0 = time.perf_counter(); f(); torch.cuda.synchronize()
    print(f"  {name}: {(time.perf_counter()-t0)*1e3:.2f} ms")
