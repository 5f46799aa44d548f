#!/usr/bin/env python
"""Reference-quality adjoint on large grids (test infrastructure only; slow, run in the background, results committed).

The reference solves every adjoint level with a sparse direct solve in fp64 (backward2_solver.py:185, :226-231).  The operator
A = I - tau L + dt/2 L^2 - dt/2 diag(f'') L has condition ~ dt/2 (8/h^2)^2: 1.4e9 at 256^2, 2.2e10 at 512^2, 3.5e11 at 1024^2, so
the reference's OWN p, q, r carry a forward error of up to eps*cond ~ 3e-7 / 5e-6 / 8e-5 there — above BASELINE's 1e-7 gradient
tolerance from 512^2 on (SURVEY.md §7 hard part 2).  Parity of the gradient on those grids is therefore pinned against the
EXACT solution of the same discrete recurrence: the oracle's factorisation (SuperLU) used as a preconditioner for iterative
refinement with residuals, right-hand sides and the q, r recursions evaluated in extended precision (numpy longdouble,
eps 1e-19).  The file also records how far the plain fp64 direct solve (= what the reference computes) is from it.

  g2d_512_adjoint    3 CN steps at 512^2  (forward: oracle, verbatim reference rule)
  g2d_1024_adjoint   2 CN steps at 1024^2 (forward: oracle with the library's floor-aware Newton stop)

usage: python oracle/make_golden_adjoint_refined.py 512 | 1024
"""
import os, sys, time
import numpy as np
import scipy.sparse as sp
from scipy.sparse.linalg import splu

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
sys.path.insert(0, HERE)
import vch_oracle as O

LD = np.longdouble


def lap_ld(v, hx, hy):
    """Mirror-ghost Neumann Laplacian (Forward2_solver.py:105-137) in extended precision."""
    v = v.astype(LD, copy=False)
    e = np.pad(v, 1, mode="reflect")
    return (e[2:, 1:-1] - 2 * v + e[:-2, 1:-1]) / LD(hx) ** 2 + (e[1:-1, 2:] - 2 * v + e[1:-1, :-2]) / LD(hy) ** 2


def fpp_ld(phi, c1, c2, eps=1e-8):
    s = np.clip(phi, -1.0 + eps, 1.0 - eps).astype(LD)
    return 2 * LD(c1) / (1 - s * s) - 2 * LD(c2)


def refined_adjoint(P, phi, x, y, t, b1, b2, phiQ, phiT, log):
    M1, nx1, ny1 = phi.shape
    n = nx1 * ny1
    hx, hy = float(x[1] - x[0]), float(y[1] - y[0])
    L = O.neumann_2d(nx1 - 1, ny1 - 1, hx, hy)
    L2 = (L @ L).tocsr()
    I = sp.eye(n, format="csr")
    tau, g = LD(P.tau), LD(P.gamma)

    def solve(Am, apply_ld, rhs_ld):
        lu = splu(Am.tocsc())
        p0 = lu.solve(np.asarray(rhs_ld, dtype=np.float64).ravel()).reshape(nx1, ny1)     # what the reference computes
        p = p0.astype(LD)
        for it in range(6):
            res = rhs_ld - apply_ld(p)
            rn = float(np.linalg.norm(res.ravel()) / np.linalg.norm(rhs_ld.ravel()))
            if rn < 1e-17:
                break
            p = p + lu.solve(np.asarray(res, dtype=np.float64).ravel()).reshape(nx1, ny1).astype(LD)
        return p0, p, rn

    p = np.zeros((M1, nx1, ny1), dtype=LD); q = np.zeros_like(p); r = np.zeros_like(p)
    p_plain = np.zeros((M1, nx1, ny1)); r_plain = np.zeros((M1, nx1, ny1)); q_plain = np.zeros((M1, nx1, ny1))
    A_T = lambda v: v - tau * lap_ld(v, hx, hy)
    rhsT = LD(b2) * (phi[-1].astype(LD) - phiT.astype(LD))
    p_plain[-1], p[-1], rn = solve(I - P.tau * L, A_T, rhsT)
    q[-1] = -lap_ld(p[-1], hx, hy)
    q_plain[-1] = -(L @ p_plain[-1].ravel()).reshape(nx1, ny1)
    log(f"terminal level solved, refined residual {rn:.1e}")
    for k in range(M1 - 2, -1, -1):
        dt = LD(t[k + 1] - t[k]); hdt = dt / 2
        f1, f0 = fpp_ld(phi[k + 1], P.c1, P.c2), fpp_ld(phi[k], P.c1, P.c2)
        src = hdt * LD(b1) * ((phi[k].astype(LD) - phiQ[k].astype(LD)) + (phi[k + 1].astype(LD) - phiQ[k + 1].astype(LD)))
        Lp1 = lap_ld(p[k + 1], hx, hy)
        rhs = p[k + 1] - tau * Lp1 - hdt * lap_ld(Lp1, hx, hy) + hdt * f1 * Lp1 + src               # B p_{k+1} + src
        f0d = np.asarray(f0, dtype=np.float64).ravel()
        Am = I - P.tau * L + 0.5 * float(dt) * L2 - 0.5 * float(dt) * (sp.diags(f0d) @ L)
        def A_ld(v, f0=f0, hdt=hdt):
            Lv = lap_ld(v, hx, hy)
            return v - tau * Lv + hdt * lap_ld(Lv, hx, hy) - hdt * f0 * Lv
        # the plain fp64 level exactly as the reference computes it (its own p_{k+1}, fp64 matrices)
        F = phi.reshape(M1, n); Q = phiQ.reshape(M1, n)
        src64 = 0.5 * float(dt) * b1 * ((F[k] - Q[k]) + (F[k + 1] - Q[k + 1]))
        Bm = I - P.tau * L - 0.5 * float(dt) * L2 + 0.5 * float(dt) * (sp.diags(O.fpp(F[k + 1], P.c1, P.c2)) @ L)
        rhs64 = Bm @ p_plain[k + 1].ravel() + src64
        lu = splu(Am.tocsc())
        p_plain[k] = lu.solve(rhs64).reshape(nx1, ny1)
        q_plain[k] = -(L @ p_plain[k].ravel()).reshape(nx1, ny1)
        den = P.gamma + 0.5 * float(dt)
        r_plain[k] = (P.gamma - 0.5 * float(dt)) / den * r_plain[k + 1] + 0.5 * float(dt) / den * (q_plain[k] + q_plain[k + 1])
        # refined level
        pk = lu.solve(np.asarray(rhs, dtype=np.float64).ravel()).reshape(nx1, ny1).astype(LD)
        for it in range(8):
            res = rhs - A_ld(pk)
            rn = float(np.linalg.norm(res.ravel()) / np.linalg.norm(rhs.ravel()))
            if rn < 1e-17:
                break
            pk = pk + lu.solve(np.asarray(res, dtype=np.float64).ravel()).reshape(nx1, ny1).astype(LD)
        p[k] = pk
        q[k] = -lap_ld(pk, hx, hy)
        r[k] = (g - hdt) / (g + hdt) * r[k + 1] + hdt / (g + hdt) * (q[k] + q[k + 1])
        log(f"level {k}: refinement steps {it}, residual {rn:.1e}")
    rel = lambda a, b: float(np.linalg.norm((a.astype(LD) - b).ravel()) / np.linalg.norm(b.ravel()))
    err = {"p": rel(p_plain, p), "q": rel(q_plain, q), "r": rel(r_plain, r)}
    return p.astype(np.float64), q.astype(np.float64), r.astype(np.float64), err


def main(N):
    t0 = time.time()
    log = lambda m: print(f"[{N}] {m}  ({time.time()-t0:.0f}s)", flush=True)
    steps, floor, stride = (3, False, 2) if N == 512 else (2, True, 4)
    P, opt = O.Phys2D(Nx=N, Ny=N, T=steps * 1e-2, dt_initial=1e-2), O.Opt2D()
    fw = O.forward_2d(P, floor_aware=floor, progress=lambda s, h: log(f"forward step {s}: {len(h)} evals"))
    phi, x, y, t = fw["phi"], fw["x"], fw["y"], fw["t"]
    phiT, phiQ = O.targets_2d(x, y, t, phi[0], P.Lx, P.Ly, P.T)
    p, q, r, err = refined_adjoint(P, phi, x, y, t, opt.b1, opt.b2, phiQ, phiT, log)
    log(f"plain fp64 direct solve vs exact recurrence: {err}")
    out = dict(stride=np.int64(stride), t=t, x=x, y=y, err_plain_p=err["p"], err_plain_q=err["q"], err_plain_r=err["r"])
    for key, a in (("phi", phi), ("p", p), ("r", r)):
        out[key] = np.ascontiguousarray(a[..., ::stride, ::stride])
        out[key + "_norm"] = np.array([np.linalg.norm(v.ravel()) for v in a])
    u1 = O.soft_prox(np.zeros_like(r), r, opt.alpha_max, opt.kappa_sparsity, opt.u_min, opt.u_max)
    out["u1"] = np.ascontiguousarray(u1[..., ::stride, ::stride]); out["u1_support"] = np.array([(u1 != 0).sum()])
    # nodes whose |alpha r| lies within the reference's own error of the threshold: their support bit is not defined by the reference
    out["u1_near_threshold"] = np.array([int((np.abs(np.abs(opt.alpha_max * r) - opt.alpha_max * opt.kappa_sparsity) < opt.alpha_max * err["r"] * np.abs(r).max()).sum())])
    np.savez_compressed(os.path.join(OUT, f"g2d_{N}_adjoint.npz"), **out)
    log("done")


if __name__ == "__main__":
    main(int(sys.argv[1]))
