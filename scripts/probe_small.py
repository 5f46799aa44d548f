"""Scratch: where does the time go on the default 128^2 grid (latency-bound regime)?"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sparse-optimal-control-of-viscous-chan-hilliard-via-gradient-descent--1d-2d_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import vch_b200_native as nat, vch_oracle as O
N, M = 128, 100
c = nat.Ctx2D(N, N, 1 / N, 1 / N, 1, 1, 0.05, 10.0, 0.75, 1.0, 1e-4)
phi0 = torch.from_numpy(O.init_phi_2d(N, N)).cuda()
dts = np.full(M, 1e-2)
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    hist, _, _ = c.forward(phi0, None, dts)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    s = c.last_stats
    print(f"forward: {dt*1e3:.1f} ms; solves {s['newton_linear_solves']}, its {s['krylov_iterations']}, resid evals {s['newton_residual_evals']}, launches {s['kernel_launches']} -> {dt*1e6/s['newton_linear_solves']:.0f} us per Newton iteration all-in")
x = np.linspace(0, 1, N + 1); t = np.concatenate([[0], np.cumsum(dts)])
phiT = torch.zeros_like(phi0); phiQ = torch.zeros_like(hist)
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    p, q, r = c.adjoint(hist, t, 5.0, 10.0, phiQ, phiT, want_pq=False)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    s = c.last_stats
    print(f"adjoint: {dt*1e3:.1f} ms; its {s['krylov_iterations']}, launches {s['kernel_launches']} -> {dt*1e6/M:.0f} us per step")
# one isolated linear solve timing
Rp, Rm = torch.randn_like(phi0), torch.randn_like(phi0)
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(50):
        d1, d2, its = c.jacobian_solve(hist[50], 1e-2, Rp, Rm)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"jacobian_solve: {dt/50*1e6:.0f} us per call, {its} its")
