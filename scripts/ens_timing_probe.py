"""Scratch probe: per-call wall time of the four calls of one ensemble PGD iteration over chained iterates."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "sparse-optimal-control-of-viscous-chan-hilliard-via-gradient-descent--1d-2d_b200")
sys.path.insert(0, PKG); sys.path.insert(0, os.path.join(PKG, "Vch_control_1D"))
import vch_b200_native as nat, GD_1D as G
B = 1024
cfg = G.ForwardSolverConfig(); ens = G.make_ensemble(B); dev = torch.device("cuda", 0)
ctx = nat.Ctx1D(cfg.N, cfg.Lx / cfg.N, cfg.Lx, cfg.tau, cfg.gamma, cfg.c1, cfg.c2, cfg.kappa)
up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
phi_init, phiQ, phiT = up(ens["phi_init"]), up(ens["phi_Q"]), up(ens["phi_T"])
hist, _, _ = ctx.forward(phi_init, None, ens["dts"]); u = torch.zeros_like(hist)
def timed(f, *a):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = f(*a); torch.cuda.synchronize(); return r, (time.perf_counter() - t0) * 1e3
for it in range(4):
    (_, _, r), ta = timed(ctx.adjoint, hist, ens["t_hist"], ens["b1"], ens["b2"], phiQ, phiT)
    (un, red), tp = timed(ctx.grad_prox, u, r, ens["b3"], 100.0, ens["ksp"], -1.0, 1.0)
    (hn, _, _), tf = timed(ctx.forward, phi_init, un, ens["dts"])
    J, tc = timed(ctx.cost, hn, un, phiQ, phiT, ens["x"], ens["t_hist"], ens["b1"], ens["b2"], ens["b3"], ens["ksp"])
    print(f"iterate {it + 1}: adjoint {ta:.2f} ms, prox {tp:.2f} ms, forward {tf:.2f} ms, cost {tc:.2f} ms; |u|max {float(un.abs().max()):.3f}, nonzero {float((un != 0).double().mean()):.3f}, sum J {float(J[:, 0].sum()):.6f}", flush=True)
    u, hist = un, hn
