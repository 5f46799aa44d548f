"""Tiny target for compute-sanitizer: one residual evaluation + one Jacobian solve + one short forward solve on an N x N grid.
    compute-sanitizer --tool racecheck|initcheck|memcheck python scripts/sanitizer_target.py [N=128] [steps=2]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "sparse-optimal-control-of-viscous-chan-hilliard-via-gradient-descent--1d-2d_b200")
sys.path.insert(0, PKG); sys.path.insert(0, os.path.join(PKG, "Vch_control_2D"))
import torch
import vch_b200_native as nat
import Forward2_solver as F2
from config import ForwardSolverConfig
N = int(sys.argv[1]) if len(sys.argv) > 1 else 128
M = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda", 0)
P = ForwardSolverConfig(Nx=N, Ny=N, T=M * 1e-2)
ctx = nat.Ctx2D(N, N, 1.0 / N, 1.0 / N, 1.0, 1.0, P.tau, P.gamma, P.c1, P.c2, P.kappa, device=0)
phi0 = torch.from_numpy(F2.init_phi_random(N, N, 1e-2, amp=0.1, seed=42)).to(dev)
w = torch.zeros_like(phi0)
mu = ctx.initialize_mu(phi0, w)
rp, rm = ctx.residual(phi0, phi0, mu, mu, w, w, 1e-2)
d1, d2, its = ctx.jacobian_solve(phi0, 1e-2, rp, rm)
h, _, _ = ctx.forward(phi0, None, np.full(M, 1e-2))
torch.cuda.synchronize()
print("ok", N, its, float(h[-1].abs().max()))
