"""Probe: which stage of the path gives different bits when K library contexts run at the same time on one GPU?
Each stage (uncontrolled forward, adjoint, controlled forward, cost) runs for K problems one after the other (twice) and then
concurrently (one thread + one torch stream per problem) on IDENTICAL inputs; outputs are compared bit for bit and the Krylov
iteration counts are printed.      python scripts/concurrency_bisect_probe.py [N=128] [M=100] [K=3] [reps=3]
"""
import os, sys, threading
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "sparse-optimal-control-of-viscous-chan-hilliard-via-gradient-descent--1d-2d_b200")
sys.path.insert(0, PKG); sys.path.insert(0, os.path.join(PKG, "Vch_control_2D"))
import torch
import vch_b200_native as nat
import Forward2_solver as F2
from config import ForwardSolverConfig, OptimizationConfig

N = int(sys.argv[1]) if len(sys.argv) > 1 else 128
M = int(sys.argv[2]) if len(sys.argv) > 2 else 100
K = int(sys.argv[3]) if len(sys.argv) > 3 else 3
REPS = int(sys.argv[4]) if len(sys.argv) > 4 else 3
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
            self.phi0 = torch.from_numpy(F2.init_phi_random(N, N, 1e-2, amp=0.1, seed=seed)).to(dev)
            xx, yy = torch.meshgrid(torch.from_numpy(x).to(dev), torch.from_numpy(x).to(dev), indexing="ij")
            self.phiT = (0.7 * torch.sin(2 * np.pi * xx) * torch.cos(np.pi * yy)).contiguous()
        self.stream.synchronize()
        self.inp = {}
        self.out = None
        self.its = None

    def stage(self, name):
        with torch.cuda.stream(self.stream):
            c = self.ctx
            if name == "forward0":
                out = c.forward(self.phi0, None, dts)[0]
            elif name == "adjoint":
                out = c.adjoint(self.inp["h0"], t_hist, Op.b1, Op.b2, self.inp["Q"], self.phiT, want_pq=False)[2]
            elif name == "forward1":
                out = c.forward(self.phi0, self.inp["u1"], dts)[0]
            elif name == "cost":
                out = torch.from_numpy(np.asarray(c.cost(self.inp["h1"], self.inp["u1"], self.inp["Q"], self.phiT, x, x, t_hist, Op.b1, Op.b2,
                                                         Op.b3, Op.kappa_sparsity)))
            self.its = dict(c.last_stats).get("krylov_iterations") if name != "cost" else None
            self.stream.synchronize()
        self.out = out


probs = [Problem(42 + k) for k in range(K)]


def run_all(name, concurrent):
    if concurrent:
        th = [threading.Thread(target=p.stage, args=(name,)) for p in probs]
        [t.start() for t in th]; [t.join() for t in th]
    else:
        for p in probs:
            p.stage(name)
    torch.cuda.synchronize()
    return [p.out.clone() for p in probs], [p.its for p in probs]


def compare(name):
    ref, its_ref = run_all(name, False)
    rows = []
    for label, conc in [("repeat", False)] + [(f"concurrent#{i}", True) for i in range(REPS)]:
        o, its = run_all(name, conc)
        eq = [bool(torch.equal(a, b)) for a, b in zip(ref, o)]
        md = [float((a - b).abs().max() / a.abs().max()) for a, b in zip(ref, o)]
        rows.append(f"   {label:13s} bit-identical {eq}  max rel diff {['%.1e' % d for d in md]}  krylov its {its} (ref {its_ref})")
    print(f"{name} ({N}^2 x {M}, {K} problems):"); print("\n".join(rows), flush=True)
    return ref


h0 = compare("forward0")
s = torch.from_numpy(t_hist / P.T).to(dev)[:, None, None]
for p, h in zip(probs, h0):
    p.inp["h0"] = h
    p.inp["Q"] = ((1 - s) * h[0] + s * p.phiT).contiguous()
r = compare("adjoint")
for p, rr in zip(probs, r):
    u0 = torch.zeros_like(rr)
    p.inp["u1"] = nat.grad_prox(u0, rr, Op.b3, Op.alpha_max, Op.kappa_sparsity, Op.u_min, Op.u_max)[0]
torch.cuda.synchronize()
h1 = compare("forward1")
for p, h in zip(probs, h1):
    p.inp["h1"] = h
compare("cost")
