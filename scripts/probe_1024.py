"""Scratch probe: 2D N^2 forward/adjoint timing + solver statistics on one GPU (device-resident)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sparse-optimal-control-of-viscous-chan-hilliard-via-gradient-descent--1d-2d_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import vch_b200_native as nat
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
M = int(sys.argv[2]) if len(sys.argv) > 2 else 20
dt = 1e-2
c = nat.Ctx2D(N, N, 1.0 / N, 1.0 / N, 1.0, 1.0, 0.05, 10.0, 0.75, 1.0, 1e-4)
rng = np.random.default_rng(42)
phi0 = 0.1 * rng.standard_normal((N + 1, N + 1)); phi0 -= phi0.mean()
phi0_d = torch.from_numpy(phi0).cuda()
dts = np.full(M, dt)
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    hist, _, _ = c.forward(phi0_d, None, dts)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    s = c.last_stats
    print(f"forward N={N} M={M}: {t1-t0:.3f}s  {1e3*(t1-t0)/M:.2f} ms/step  stats={s}", flush=True)
x = np.linspace(0, 1, N + 1); t = dt * np.arange(M + 1)
xx, yy = np.meshgrid(x, x, indexing="ij")
phiT = torch.from_numpy(0.7 * np.sin(2 * np.pi * xx) * np.cos(np.pi * yy)).cuda()
s_ = torch.from_numpy(t / t[-1]).cuda()[:, None, None]
phiQ = (1 - s_) * hist[0] + s_ * phiT
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    p, q, r = c.adjoint(hist, t, 5.0, 10.0, phiQ, phiT, want_pq=False)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"adjoint: {t1-t0:.3f}s  {1e3*(t1-t0)/M:.2f} ms/step stats={c.last_stats}", flush=True)
u = torch.zeros_like(hist)
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    u1, h1, J, red, st = c.pgd_iteration(u, hist, phiQ, phiT, t, dts, x, x, 5.0, 10.0, 1e-4, 1e-4, -1.0, 1.0, 50.0)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"pgd iteration: {t1-t0:.3f}s J={J} red={red} stats={st}", flush=True)
print("phi max", float(h1.abs().max()), "u1 frac at bounds", float((u1.abs() == 1).double().mean()))
