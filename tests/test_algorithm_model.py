"""CPU checks of the ALGORITHM the CUDA path implements (not of the CUDA code): `scripts/krylov_proto.py` is a NumPy model of
the Schur reduction, the DCT-I preconditioner, the stencil-free form of the preconditioned operator and the BiCGStab / Newton
loops exactly as csrc/vch2d.cu runs them.  Against the oracle (SciPy SuperLU on the assembled 2n x 2n Jacobian, i.e. the
reference's path) this pins the design decisions that no GPU is needed to verify:
  * the reduced, preconditioned, matrix-free solve reproduces the direct solve to round-off;
  * the forcing term (first linear solve of each Newton solve to 1e-6) leaves the trajectory where it was;
  * the three-dot-product formula behind the BiCGStab half-step exit is accurate where the kernel trusts it."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "scripts"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)
import krylov_proto as K      # noqa: E402
import vch_oracle as O        # noqa: E402

rel = lambda a, b: float(np.linalg.norm(a - b) / np.linalg.norm(b))


def _model_run(P, steps, first_tol=None):
    N = P.Nx
    G = K.Grid(N, kappa=P.kappa, tau=P.tau, gamma=P.gamma, c1=P.c1, c2=P.c2, dt=P.dt_initial)
    W = np.outer(K.trapz_w(N), K.trapz_w(N))
    phi = O.init_phi_2d(N, N)
    mu = -G.kappa * G.lap(phi) + G.c1 * G.flog(phi) - 2 * G.c2 * phi      # initialize_mu(phi0, w = 0)
    w = np.zeros_like(phi); z = np.zeros_like(phi)
    m0 = (phi * W).sum() * G.h ** 2
    hist, its, newton = [phi.copy()], 0.0, 0
    orig = K.bicgstab
    state = {"k": 0}

    def solver(G_, a, b, abar, tol=1e-11, maxit=200):
        t = first_tol if (first_tol and state["k"] == 0) else tol
        state["k"] += 1
        return orig(G_, a, b, abar, t, maxit)

    K.bicgstab = solver
    try:
        for _ in range(steps):
            st = {"its": [], "newton": []}
            state["k"] = 0
            phi, mu, w = K.step(G, phi, mu, w, z, z, "geo", st, floor=False)
            me = (phi * W).sum() * G.h ** 2 - m0
            inter = np.abs(phi) < 0.985
            wi = (W * inter).sum() * G.h ** 2
            if wi > 0:
                phi = np.where(inter, phi - me / wi, phi)
            hist.append(phi.copy()); its += sum(st["its"]); newton += len(st["its"])
    finally:
        K.bicgstab = orig
    return np.array(hist), its, newton


def test_matrix_free_dct_krylov_solve_reproduces_the_direct_solve():
    P = O.Phys2D(Nx=32, Ny=32, T=0.08)
    fw = O.forward_2d(P)                                   # SuperLU on the full Jacobian (reference path)
    hist, its, newton = _model_run(P, 8)
    assert rel(hist, fw["phi"]) < 1e-12
    assert newton == int(np.sum(fw["nres"] - 1))            # residual history length - 1 = linear solves: same as the reference


def test_forcing_term_keeps_the_trajectory():
    P = O.Phys2D(Nx=48, Ny=48, T=0.3)
    strict, its0, n0 = _model_run(P, 30)
    forced, its1, n1 = _model_run(P, 30, first_tol=1e-6)
    assert n1 == n0                                        # Newton takes the same number of iterations
    assert its1 < 0.9 * its0                               # ... with fewer BiCGStab iterations
    assert rel(forced, strict) < 1e-11                     # far inside BASELINE's 1e-8


def test_half_step_norm_formula():
    """||s||^2 = (r,r) - 2 alpha (r,v) + alpha^2 (v,v) for s = r - alpha v: rounding error ~1e-13 (r,r), so it resolves
    ||s||^2 >= 1e-6 (r,r) (the kernel's trust region: (r,r) <= 1e6 thr2 and ||s||^2 <= thr2) to better than 1e-6 relative."""
    rng = np.random.default_rng(3)
    n = 65 * 65
    r = rng.standard_normal(n)
    for ratio in (1e-1, 1e-2, 1e-3):                       # ||s|| / ||r||
        e = rng.standard_normal(n); e *= ratio * np.linalg.norm(r) / np.linalg.norm(e)
        alpha = 0.7
        v = (r - e) / alpha                                # so that s = r - alpha v = e
        s = r - alpha * v
        formula = r @ r - 2 * alpha * (r @ v) + alpha ** 2 * (v @ v)
        assert abs(formula - s @ s) <= 1e-6 * (s @ s)
