#!/usr/bin/env python
"""Golden vectors for BASELINE config 4 (ensemble of independent 1D control problems): 16 randomly chosen members of the
1024-problem ensemble, each taken through one optimistic PGD iteration by the UNMODIFIED 1D reference from /root/reference
(GD_1D.py:359-376: run_backward -> calculate_gradient -> perform_gradient_step -> perform_proximal_and_projection ->
run_main_simulation -> calculate_cost).  Test infrastructure only; ~1 min.

The ensemble is SURVEY.md §8(d) config 4: physics fixed at the 1D defaults (the reference's adjoint cannot vary them,
backward_solver.py:29-33); per-problem draws from np.random.default_rng(1234), in this order: choice_t ~ U{1,2,3},
A_T ~ U[0.3, 0.8], kappa_sp ~ logU[1e-5, 1e-3], b1 ~ U[0.1, 1], b2 ~ U[5, 20], b3 ~ logU[1e-4, 1e-2]; target profiles as
GD_1D.py:190-248 with amplitude A_T, tracking path = linear ramp phi_0 -> phi_T.  Members: default_rng(7).choice(1024, 16).

usage: python oracle/make_golden_ensemble.py
"""
import io, os, sys, contextlib, time
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
sys.path.insert(0, os.path.join(HERE, "_mpl_shim"))
sys.path.insert(0, "/root/reference/src/1D/Vch_control_1D")
os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache_golden")
os.chdir("/tmp")
import Forward_solver as F, backward_solver as B, cost_and_function as C, GD_1D as G
from config import ForwardSolverConfig


def quiet(f, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return f(*a, **k)


def profile(x, Lx, ct, A):
    if ct == 1:
        return A * np.sin(2.0 * np.pi * x / Lx)
    if ct == 2:
        return A * np.cos(2.0 * np.pi * x / Lx)
    raw = np.tan(2.0 * np.pi * 0.45 * (x / Lx - 0.5))
    sc = np.max(np.abs(raw))
    return A * raw / (sc if sc > 1e-12 else 1.0)


def main():
    Btot, alpha = 1024, 100.0
    rng = np.random.default_rng(1234)
    ct = rng.integers(1, 4, Btot); A = rng.uniform(0.3, 0.8, Btot)
    ksp = np.exp(rng.uniform(np.log(1e-5), np.log(1e-3), Btot)); b1 = rng.uniform(0.1, 1.0, Btot); b2 = rng.uniform(5, 20, Btot)
    b3 = np.exp(rng.uniform(np.log(1e-4), np.log(1e-2), Btot))
    members = np.sort(np.random.default_rng(7).choice(Btot, 16, replace=False))
    cfg = ForwardSolverConfig()
    t0 = time.time()
    phi0, x, t = quiet(F.run_main_simulation, cfg, store_history=True, control_input=None, verbose=False)
    out = dict(members=members, x=x, t=t, phi0=phi0, alpha=np.float64(alpha),
               choice_t=ct[members], A_T=A[members], ksp=ksp[members], b1=b1[members], b2=b2[members], b3=b3[members])
    R, U1, P1, J1 = [], [], [], []
    for m in members:
        phiT = profile(x, cfg.Lx, int(ct[m]), float(A[m]))
        s = (t / t[-1])[:, None]
        phiQ = (1.0 - s) * phi0[0] + s * phiT
        u0 = np.zeros_like(phi0)
        p, q, r = B.run_backward(phi0, x, t, float(b1[m]), float(b2[m]), phiQ, phiT)
        g = C.calculate_gradient(r, u0, float(b3[m]))
        u1 = G.perform_proximal_and_projection(C.perform_gradient_step(u0, g, alpha), alpha, float(ksp[m]), -1.0, 1.0)
        phi1, _, _ = quiet(F.run_main_simulation, cfg, store_history=True, control_input=u1, verbose=False)
        J = quiet(C.calculate_cost, phi1, u1, phiQ, phiT, x, t, float(b1[m]), float(b2[m]), float(b3[m]), float(ksp[m]))
        R.append(r); U1.append(u1); P1.append(phi1); J1.append(J)
        print(f"member {m}: choice_t {ct[m]}, J1 {J!r}  ({time.time()-t0:.0f}s)", flush=True)
    out.update(r0=np.array(R), u1=np.array(U1), phi1=np.array(P1), J1=np.array(J1))
    np.savez_compressed(os.path.join(OUT, "g1d_ensemble16.npz"), **out)
    print("done")


if __name__ == "__main__":
    main()
