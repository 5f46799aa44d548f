"""Probe: run-to-run reproducibility of ONE context on one GPU (no concurrency).  Repeats the uncontrolled forward solve, finds the
first trajectory level whose bits differ between repeats, then repeats the primitives (Newton step, Jacobian solve, residual) from
the last identical state.      python scripts/determinism_probe.py [N=1024] [M=30] [repeats=4]
"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "sparse-optimal-control-of-viscous-chan-hilliard-via-gradient-descent--1d-2d_b200")
sys.path.insert(0, PKG); sys.path.insert(0, os.path.join(PKG, "Vch_control_2D"))
import torch
import vch_b200_native as nat
import Forward2_solver as F2
from config import ForwardSolverConfig

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
M = int(sys.argv[2]) if len(sys.argv) > 2 else 30
R = int(sys.argv[3]) if len(sys.argv) > 3 else 4
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
nat.require_device()
dt = 1e-2
P = ForwardSolverConfig(Nx=N, Ny=N, T=M * dt)
dts = np.full(M, dt)
ctx = nat.Ctx2D(N, N, 1.0 / N, 1.0 / N, 1.0, 1.0, P.tau, P.gamma, P.c1, P.c2, P.kappa, device=0)
phi0 = torch.from_numpy(F2.init_phi_random(N, N, 1e-2, amp=0.1, seed=42)).to(dev)

runs = []
for i in range(R):
    h, mh, wh = ctx.forward(phi0, None, dts, want_mu=True, want_w=True)
    torch.cuda.synchronize()
    st = dict(ctx.last_stats)
    runs.append((h.clone(), mh.clone(), wh.clone(), st))
    print(f"forward run {i}: krylov its {st.get('krylov_iterations')}, linear solves {st.get('newton_linear_solves')}, "
          f"residual evals {st.get('newton_residual_evals')}, half exits {st.get('krylov_half_exits')}", flush=True)
first = None
for i in range(1, R):
    lv = [k for k in range(M + 1) if not torch.equal(runs[0][0][k], runs[i][0][k])]
    mlv = [k for k in range(M) if not torch.equal(runs[0][1][k], runs[i][1][k])]
    print(f"run 0 vs run {i}: first differing phi level {lv[0] if lv else None} ({len(lv)} differ), first differing mu step {mlv[0] if mlv else None}"
          + (f", max |dphi| at that level {float((runs[0][0][lv[0]] - runs[i][0][lv[0]]).abs().max()):.2e}" if lv else ""), flush=True)
    if lv and (first is None or lv[0] < first):
        first = lv[0]
for a in range(1, R):
    for b in range(a + 1, R):
        same = torch.equal(runs[a][0], runs[b][0])
        print(f"run {a} vs run {b}: trajectories bit-identical {same}")

k = (first - 1) if first else M // 2          # last identical level (or the middle when every repeat agreed)
print(f"primitives from level {k}")
h, mh, wh, _ = runs[0]
phi_old = h[k].contiguous()
if k == 0:
    w_old = torch.zeros_like(phi_old); mu_old = ctx.initialize_mu(phi_old, w_old)
else:
    mu_old = mh[k - 1].contiguous(); w_old = wh[k - 1].contiguous()
w_new = wh[k].contiguous()


def rep(name, fn, n=int(os.environ.get("PROBE_REPS", 8))):
    outs = []
    for _ in range(n):
        o = fn(); torch.cuda.synchronize(); outs.append(o)
    ref = outs[0]
    flags = []
    for o in outs[1:]:
        eq = all(bool(torch.equal(x, y)) for x, y in zip(ref[0], o[0]))
        md = max(float((x - y).abs().max()) for x, y in zip(ref[0], o[0]))
        flags.append(("=" if eq else f"{md:.1e}"))
    print(f"{name}: vs first call {flags}; extra {[o[1] for o in outs]}", flush=True)


def newton():
    pn, mn, hist = ctx.newton(phi_old, mu_old, w_old, w_new, dt)
    return (pn, mn), (len(hist), ctx.last_stats.get("krylov_iterations"))


rp, rm = ctx.residual(phi_old, phi_old, mu_old, mu_old, w_new, w_old, dt)
rep("residual", lambda: (ctx.residual(phi_old, phi_old, mu_old, mu_old, w_new, w_old, dt), None))


def jsolve():
    d1, d2, its = ctx.jacobian_solve(phi_old, dt, rp, rm)
    return (d1, d2), its


rep("jacobian_solve", jsolve)
rep("newton step", newton)
rep("apply_laplacian", lambda: ((ctx.apply_laplacian(phi_old),), None))
