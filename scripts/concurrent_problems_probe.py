"""Probe: K independent 2D control problems on ONE GPU at the same time (one library context, one stream, one host thread each)
against the same K problems one after the other.  The kernels of the Krylov loop are single-wave latency chains,
so a second problem can only use what the first leaves idle: SMs without a CTA in the column solve (19 of 148), issue slots while
a wave drains.  Aggregate PGD iterations / s, wall clock around all threads with a device synchronisation on both sides.

    python scripts/concurrent_problems_probe.py [N=1024] [M=100] [iters=3] [Kmax=3]
"""
import os, sys, threading, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "sparse-optimal-control-of-viscous-chan-hilliard-via-gradient-descent--1d-2d_b200")
sys.path.insert(0, PKG); sys.path.insert(0, os.path.join(PKG, "Vch_control_2D"))
import torch
import vch_b200_native as nat
import Forward2_solver as F2
from config import ForwardSolverConfig, OptimizationConfig

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
M = int(sys.argv[2]) if len(sys.argv) > 2 else 100
ITERS = int(sys.argv[3]) if len(sys.argv) > 3 else 3
KMAX = int(sys.argv[4]) if len(sys.argv) > 4 else 3
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
nat.require_device()
dt = 1e-2
P, Op = ForwardSolverConfig(Nx=N, Ny=N, T=M * dt), OptimizationConfig()
dts = np.full(M, dt)
t_hist = np.concatenate([[0.0], np.minimum(np.cumsum(dts), P.T)])
x = np.linspace(0.0, 1.0, N + 1)


class Problem:
    def __init__(self, seed):
        self.stream = torch.cuda.Stream()
        with torch.cuda.stream(self.stream):
            self.ctx = nat.Ctx2D(N, N, 1.0 / N, 1.0 / N, 1.0, 1.0, P.tau, P.gamma, P.c1, P.c2, P.kappa, device=0)
            phi0 = torch.from_numpy(F2.init_phi_random(N, N, 1e-2, amp=0.1, seed=seed)).to(dev)
            self.h, _, _ = self.ctx.forward(phi0, None, dts)
            xx, yy = torch.meshgrid(torch.from_numpy(x).to(dev), torch.from_numpy(x).to(dev), indexing="ij")
            self.phiT = (0.7 * torch.sin(2 * np.pi * xx) * torch.cos(np.pi * yy)).contiguous()
            s = torch.from_numpy(t_hist / P.T).to(dev)[:, None, None]
            self.phiQ = ((1 - s) * self.h[0] + s * self.phiT).contiguous()
            self.u = torch.zeros_like(self.h)
            self.un, self.hn, self.r = torch.empty_like(self.h), torch.empty_like(self.h), torch.empty_like(self.h)
            self.J = None
        self.stream.synchronize()
        self.snapshot = (self.u.clone(), self.h.clone())

    def reset(self):
        self.u.copy_(self.snapshot[0]); self.h.copy_(self.snapshot[1])

    def run(self, iters):
        with torch.cuda.stream(self.stream):
            for _ in range(iters):
                _, _, J, _, _ = self.ctx.pgd_iteration(self.u, self.h, self.phiQ, self.phiT, t_hist, dts, x, x, Op.b1, Op.b2, Op.b3,
                                                       Op.kappa_sparsity, Op.u_min, Op.u_max, Op.alpha_max, u_out=self.un,
                                                       phi_out=self.hn, r_out=self.r)
                self.u, self.un = self.un, self.u
                self.h, self.hn = self.hn, self.h
                self.J = float(J[0])
            self.stream.synchronize()


probs = [Problem(42 + k) for k in range(KMAX)]
for p in probs:                      # warm-up (graphs, staging) from the same start
    p.run(1); p.reset()
torch.cuda.synchronize()
for K in range(1, KMAX + 1):
    def sequential():
        for p in probs[:K]:
            p.reset()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for p in probs[:K]:
            p.run(ITERS)
        torch.cuda.synchronize()
        return time.perf_counter() - t0, [p.J for p in probs[:K]]
    seq, Js = sequential()
    _, Js2 = sequential()                                   # a second time: separates history dependence from concurrency
    for p in probs[:K]:
        p.reset()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    th = [threading.Thread(target=p.run, args=(ITERS,)) for p in probs[:K]]
    [t.start() for t in th]; [t.join() for t in th]
    torch.cuda.synchronize(); con = time.perf_counter() - t0
    Jc = [p.J for p in probs[:K]]
    rel = lambda A, B: max(abs(a - b) / abs(a) for a, b in zip(A, B))
    print(f"{N}^2 x {M} steps, {K} problem(s) x {ITERS} PGD iterations: one after the other {K * ITERS / seq:.3f} it/s, "
          f"concurrent {K * ITERS / con:.3f} it/s ({seq / con:.3f}x); J rel. diff: repeat {rel(Js, Js2):.2e}, concurrent {rel(Js, Jc):.2e}", flush=True)
    if os.environ.get("PROBE_VERBOSE"):
        print("   J seq ", [repr(j) for j in Js]); print("   J seq2", [repr(j) for j in Js2]); print("   J conc", [repr(j) for j in Jc], flush=True)
