"""Scratch probe: effect of the BiCGStab relative tolerance on time and on the iterates (1 GPU, device-resident).
Reference result = rtol 1e-11 (the default, parity-proven against the oracle)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sparse-optimal-control-of-viscous-chan-hilliard-via-gradient-descent--1d-2d_b200"))
import vch_b200_native as nat
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
M = int(sys.argv[2]) if len(sys.argv) > 2 else 100
T = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
dt = T / M
c = nat.Ctx2D(N, N, 1.0 / N, 1.0 / N, 1.0, 1.0, 0.05, 10.0, 0.75, 1.0, 1e-4)
rng = np.random.default_rng(42)
phi0 = 0.1 * rng.standard_normal((N + 1, N + 1)); phi0 -= phi0.mean()
phi0_d = torch.from_numpy(phi0).cuda()
dts = np.full(M, dt)
x = np.linspace(0, 1, N + 1); t = dt * np.arange(M + 1)
xx, yy = np.meshgrid(x, x, indexing="ij")
phiT = torch.from_numpy(0.7 * np.sin(2 * np.pi * xx) * np.cos(np.pi * yy)).cuda()
s_ = torch.from_numpy(t / t[-1]).cuda()[:, None, None]
base = None
def rel(a, b): return float((a - b).norm() / b.norm())
for rtol in (1e-11, 1e-10, 1e-9, 1e-8, 1e-7):
    c.set_krylov(rtol, 200)
    hist0, _, _ = c.forward(phi0_d, None, dts)
    phiQ = (1 - s_) * hist0[0] + s_ * phiT
    u = torch.zeros_like(hist0)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    u1, h1, J, red, st = c.pgd_iteration(u, hist0, phiQ, phiT, t, dts, x, x, 5.0, 10.0, 1e-4, 1e-4, -1.0, 1.0, 50.0)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    u2, h2, J2, red2, st2 = c.pgd_iteration(u1, h1, phiQ, phiT, t, dts, x, x, 5.0, 10.0, 1e-4, 1e-4, -1.0, 1.0, 50.0)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    res = dict(hist0=hist0, u1=u1, h1=h1, J=J[0], u2=u2, h2=h2, J2=J2[0])
    line = f"rtol={rtol:g}: it1 {t1-t0:.3f}s it2 {t2-t1:.3f}s stats2={st2}"
    if base is None:
        base = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in res.items()}
    else:
        line += (f" | rel hist0 {rel(hist0, base['hist0']):.2e} u1 {rel(u1, base['u1']):.2e} h1 {rel(h1, base['h1']):.2e}"
                 f" u2 {rel(u2, base['u2']):.2e} h2 {rel(h2, base['h2']):.2e} dJ {abs(J[0]-base['J'])/abs(base['J']):.2e}"
                 f" dJ2 {abs(J2[0]-base['J2'])/abs(base['J2']):.2e}"
                 f" support1 {int(((u1 != 0) != (base['u1'] != 0)).sum())} support2 {int(((u2 != 0) != (base['u2'] != 0)).sum())}")
    print(line, flush=True)
    del hist0, u1, h1, u2, h2, res, phiQ, u
