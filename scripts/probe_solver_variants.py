"""Scratch probe (one GPU): what the forcing term (first Newton solve to 1e-6) and the BiCGStab half-step exit change at
full size.  Runs M forward steps + the adjoint sweep on an N^2 grid under
  strict  (VCH_NO_HALF_EXIT=1, set_krylov_first(0))      the pre-change solver
  forced  (VCH_NO_HALF_EXIT=1)                             forcing term only
  default                                                  forcing term + half-step exit
and prints time per step, solver statistics and the relative differences of phi / r to the strict run.
Usage: python scripts/probe_solver_variants.py [N=1024] [M=150]"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sparse-optimal-control-of-viscous-chan-hilliard-via-gradient-descent--1d-2d_b200"))
import vch_b200_native as nat
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
M = int(sys.argv[2]) if len(sys.argv) > 2 else 150
dt = 1e-2
rng = np.random.default_rng(42)
phi0 = 0.1 * rng.standard_normal((N + 1, N + 1)); phi0 -= phi0.mean()
phi0_d = torch.from_numpy(phi0).cuda()
dts = np.full(M, dt); t = dt * np.arange(M + 1)
x = np.linspace(0, 1, N + 1); xx, yy = np.meshgrid(x, x, indexing="ij")
phiT = torch.from_numpy(0.7 * np.sin(2 * np.pi * xx) * np.cos(np.pi * yy)).cuda()
u = torch.from_numpy(0.5 * np.sin(3 * np.pi * xx) * np.cos(2 * np.pi * yy)).cuda()[None].repeat(M + 1, 1, 1).contiguous()
rel = lambda a, b: float((a - b).norm() / b.norm())
ref = None
for name, env, first in (("strict", True, 0.0), ("forced", True, 1e-6), ("default", False, 1e-6)):
    os.environ.pop("VCH_NO_HALF_EXIT", None)
    if env:
        os.environ["VCH_NO_HALF_EXIT"] = "1"
    c = nat.Ctx2D(N, N, 1.0 / N, 1.0 / N, 1.0, 1.0, 0.05, 10.0, 0.75, 1.0, 1e-4)
    c.set_krylov_first(first)
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        hist, _, _ = c.forward(phi0_d, u, dts)
        torch.cuda.synchronize(); t1 = time.perf_counter()
    sf = dict(c.last_stats)
    s_ = torch.from_numpy(t / t[-1]).cuda()[:, None, None]
    phiQ = (1 - s_) * hist[0] + s_ * phiT
    for rep in range(2):
        torch.cuda.synchronize(); t2 = time.perf_counter()
        _, _, r = c.adjoint(hist, t, 5.0, 10.0, phiQ, phiT, want_pq=False)
        torch.cuda.synchronize(); t3 = time.perf_counter()
    sa = dict(c.last_stats)
    if ref is None:
        ref = (hist.clone(), r.clone())
    print(f"{name:8s} forward {1e3*(t1-t0)/M:.3f} ms/step its {sf['krylov_iterations']} solves {sf['newton_linear_solves']} "
          f"half {sf['krylov_half_exits']} stalls {sf['krylov_stalls']} | adjoint {1e3*(t3-t2)/M:.3f} ms/step its {sa['krylov_iterations']} "
          f"half {sa['krylov_half_exits']} | rel phi {rel(hist, ref[0]):.2e} (last level {rel(hist[-1], ref[0][-1]):.2e}) rel r {rel(r, ref[1]):.2e}", flush=True)
    del c, hist, r
