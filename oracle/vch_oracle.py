"""CPU oracle for the vCH forward/adjoint/PGD hot path.  TEST INFRASTRUCTURE ONLY.

A NumPy/SciPy restatement of the reference algorithm, written from the formulas (not a copy of
the sources), each function citing the reference file:line it follows.  Only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of `bench.py` may
import this module; the product path (the CUDA library behind include/vch_b200.h) never does.

Parity pin: `tests/test_oracle_golden.py` checks every function here against the vectors in
`tests/golden/*.npz`, which `oracle/make_golden.py` produced by running the unmodified
reference from /root/reference (the reference stores no golden vectors of its own).

Paths below are relative to /root/reference/src/.
"""
from __future__ import annotations

import json
from dataclasses import dataclass, asdict

import numpy as np
import scipy.sparse as sp
from scipy.sparse.linalg import spsolve, splu

DELTA_SEP = 1e-2            # 2D/Vch_control_2D/Forward2_solver.py:510, 1D/Vch_control_1D/Forward_solver.py:42


# --------------------------------------------------------------------------- parameters
@dataclass
class Phys2D:               # defaults: 2D/Vch_control_2D/config.py:103-113
    Nx: int = 128
    Ny: int = 128
    Lx: float = 1.0
    Ly: float = 1.0
    T: float = 1.0
    dt_initial: float = 1e-2
    tau: float = 0.05
    gamma: float = 10.0
    c1: float = 0.75
    c2: float = 1.0
    kappa: float = 0.01 ** 2


@dataclass
class Opt2D:                # 2D/Vch_control_2D/config.py:137-144
    b1: float = 5.0
    b2: float = 10.0
    b3: float = 1e-4
    kappa_sparsity: float = 1e-4
    alpha_max: float = 50.0
    max_iter: int = 500
    u_min: float = -1.0
    u_max: float = 1.0


@dataclass
class Phys1D:               # 1D/Vch_control_1D/config.py:93-102
    N: int = 128
    Lx: float = 1.0
    T: float = 1.0
    dt_initial: float = 1e-2
    tau: float = 0.05
    gamma: float = 10.0
    c1: float = 0.75
    c2: float = 1.0
    kappa: float = 0.03 ** 2


@dataclass
class Opt1D:                # 1D/Vch_control_1D/config.py:115-123
    b1: float = 0.3
    b2: float = 13.0
    b3: float = 0.0019
    kappa_sparsity: float = 0.00009
    alpha_max: float = 100.0
    max_iter: int = 1000
    u_min: float = -1.0
    u_max: float = 1.0


def from_json(cls, s):
    d = json.loads(str(s))
    return cls(**{k: d[k] for k in cls.__dataclass_fields__ if k in d})


# --------------------------------------------------------------------------- shared pieces
def trapz_w(n):
    """[1/2,1,...,1,1/2] — Forward2_solver.py:430-441, Forward_solver.py:237-241."""
    w = np.ones(n)
    w[0] = w[-1] = 0.5
    return w


def neumann_1d(N, h):
    """Mirror-ghost second difference, (N+1)x(N+1) — Forward2_solver.py:105-122 / Forward_solver.py:64-76."""
    a = 1.0 / (h * h)
    L = sp.lil_matrix((N + 1, N + 1))
    L.setdiag(-2.0 * a)
    L.setdiag(a, 1)
    L.setdiag(a, -1)
    L[0, 1] = 2.0 * a
    L[N, N - 1] = 2.0 * a
    return L.tocsr()


def neumann_2d(Nx, Ny, hx, hy):
    """kron(I_{Ny+1}, L1d(Nx,hx)) + kron(L1d(Ny,hy), I_{Nx+1}) — Forward2_solver.py:125-137.

    NB the factor order: on a C-order flattened (Nx+1, Ny+1) field with Nx == Ny this applies the
    hx stencil along the contiguous axis; for Nx != Ny it acts on the flat vector viewed as
    (Ny+1, Nx+1).  The CUDA stencils reproduce exactly this operator.
    """
    return (sp.kron(sp.eye(Ny + 1), neumann_1d(Nx, hx)) + sp.kron(neumann_1d(Ny, hy), sp.eye(Nx + 1))).tocsr()


def flory_log(phi, eps):
    """log((1+s)/(1-s)), s = clip(phi, ±(1-eps)) — Forward2_solver.py:86-102, Forward_solver.py:57-62."""
    s = np.clip(phi, -1.0 + eps, 1.0 - eps)
    return np.log((1.0 + s) / (1.0 - s))


def log_eps(delta_sep=DELTA_SEP):
    return max(1e-8, 0.5 * delta_sep)


def w_update(w, dt, gamma, un, unp1):
    """CN step of gamma w' + w = u — Forward2_solver.py:170-181, Forward_solver.py:88-91."""
    g = gamma / dt
    return ((g - 0.5) * w + 0.5 * (unp1 + un)) / (g + 0.5)


def fpp(phi, c1, c2, eps=1e-8):
    """f''(phi) with clip — backward2_solver.py:40-72, backward_solver.py:36-46."""
    s = np.clip(phi, -1.0 + eps, 1.0 - eps)
    return 2.0 * c1 / (1.0 - s * s) - 2.0 * c2


def soft_prox(u, g, alpha, kappa_sp, umin, umax):
    """u - alpha g → soft threshold alpha*kappa → box — cost2_and_function.py:191-200; GD_1D.py:56-71 + cost_and_function.py:111."""
    v = u - alpha * g
    v = np.sign(v) * np.maximum(np.abs(v) - alpha * kappa_sp, 0.0)
    return np.clip(v, umin, umax)


def kkt_counts(u, r, kappa_sp, tol=1e-6):
    """(#|u|<tol, #|r|<=kappa, #agree) — second_order_conditions_2d.py:238-297, GD_1D.py:115-147."""
    a = np.abs(u) < tol
    b = np.abs(r) <= kappa_sp
    return int(a.sum()), int(b.sum()), int((a == b).sum())


# --------------------------------------------------------------------------- 2D forward
class Grid2D:
    def __init__(self, P: Phys2D):
        self.P = P
        self.hx, self.hy = P.Lx / P.Nx, P.Ly / P.Ny
        self.x = np.linspace(0.0, P.Lx, P.Nx + 1)
        self.y = np.linspace(0.0, P.Ly, P.Ny + 1)
        self.L = neumann_2d(P.Nx, P.Ny, self.hx, self.hy)
        self.shape = (P.Nx + 1, P.Ny + 1)
        self.wts_h = self.hx * self.hy * np.outer(trapz_w(P.Nx + 1), trapz_w(P.Ny + 1))   # Forward2_solver.py:528-531

    def lap(self, v):
        return (self.L @ v.ravel()).reshape(self.shape)


def init_phi_2d(Nx, Ny, delta_sep=DELTA_SEP, amp=0.1, seed=42):
    """Forward2_solver.py:444-486 (zero weighted mean, clip, interior-only re-centering)."""
    phi = amp * np.random.default_rng(seed).standard_normal((Nx + 1, Ny + 1))
    w = np.outer(trapz_w(Nx + 1), trapz_w(Ny + 1))
    W = w.sum()
    phi -= (w * phi).sum() / W
    lo, hi = -1.0 + delta_sep, 1.0 - delta_sep
    phi = np.clip(phi, lo, hi)
    for _ in range(8):
        m = (w * phi).sum()
        if abs(m) <= 1e-14 * W:
            break
        inner = np.abs(phi) < hi - 5e-3
        Wi = float(w[inner].sum())
        if Wi <= 0:
            phi = np.clip(phi - m / W, lo, hi)
            break
        phi[inner] -= m / Wi
    return phi


def mu_init_2d(G: Grid2D, phi, w):
    """mu = -kappa L phi + c1 log(..) - 2 c2 phi - w — Forward2_solver.py:155-167."""
    P = G.P
    return -P.kappa * G.lap(phi) + P.c1 * flory_log(phi, log_eps()) - 2.0 * P.c2 * phi - w


def residual_2d(G: Grid2D, phi, mu, phi0, mu0, w1, w0, dt):
    """[R_phi, R_mu] — Forward2_solver.py:184-221."""
    P = G.P
    Rphi = (P.tau * (phi - phi0) / dt - 0.5 * P.kappa * (G.lap(phi) + G.lap(phi0))
            + (P.c1 * flory_log(phi, log_eps()) - 2.0 * P.c2 * phi0) - 0.5 * (mu + mu0) - 0.5 * (w1 + w0))
    Rmu = (phi - phi0) / dt - 0.5 * (G.lap(mu) + G.lap(mu0))
    return Rphi, Rmu


def jacobian_2d(G: Grid2D, phi, dt):
    """2x2 block Jacobian — Forward2_solver.py:224-253 (phi^2 clipped to 1-delta_sep^2 in the diagonal)."""
    P = G.P
    n = phi.size
    d = P.tau / dt + 2.0 * P.c1 / (1.0 - np.clip(phi.ravel() ** 2, 0.0, 1.0 - DELTA_SEP ** 2))
    I = sp.eye(n, format="csr")
    return sp.bmat([[sp.diags(d) - 0.5 * P.kappa * G.L, -0.5 * I], [I / dt, -0.5 * G.L]], format="csc")


def newton_2d(G: Grid2D, phi0, mu0, w0, w1, dt, tol=1e-6, max_iter=500, floor_aware=False):
    """Forward2_solver.py:323-427.  Returns (phi, mu, residual history).

    floor_aware=True adds the stop rule of the CUDA library (DESIGN.md "fp64-floor-aware Newton stop", csrc/vch2d.cu
    newton_step): R_mu contains L mu, mu is stored to eps|mu| and L amplifies that by 1/hx^2+1/hy^2, so the MEASURED ||R|| cannot
    go below floor = eps (1/hx^2+1/hy^2) ||mu||_2.  Where 1.5 floor >= tol (>= 1024^2; on the reference's grids the rule never
    fires) the reference's criterion ||R|| < tol is applied to the residual with its rounding noise removed: after a full Newton
    step the true residual is the nonlinear remainder c1 [l(phi+dphi) - l(phi) - l'(phi) dphi].  Safety net: stop when an
    iteration fails to halve ||R|| inside 50x of the floor.  Off by default: the pinned goldens use the verbatim reference rule."""
    n = phi0.size
    phi, mu = phi0.copy(), mu_init_2d(G, phi0, w1)                 # :350-351
    hist = []
    lim = 1.0 - DELTA_SEP
    floor = lambda m: 2.220446049250313e-16 * (1.0 / G.hx ** 2 + 1.0 / G.hy ** 2) * np.linalg.norm(m.ravel())
    nR_prev, true_est = None, np.inf
    for _ in range(max_iter):
        Rp, Rm = residual_2d(G, phi, mu, phi0, mu0, w1, w0, dt)
        R = np.concatenate([Rp.ravel(), Rm.ravel()])
        nR = np.linalg.norm(R)
        hist.append(nR)
        if nR < tol:                                               # :364
            break
        if floor_aware:
            fl = floor(mu)
            if (1.5 * fl >= tol and nR < 50.0 * fl and true_est < tol) or (nR_prev is not None and nR > 0.5 * nR_prev and nR < 50.0 * fl):
                break
            nR_prev = nR
        d = spsolve(jacobian_2d(G, phi, dt), -R)                   # :370
        dphi, dmu = d[:n], d[n:]
        pf = phi.ravel()
        amax = 2.0                                                 # :377-391
        with np.errstate(divide="ignore", invalid="ignore"):
            pos, neg = dphi > 0, dphi < 0
            if pos.any():
                amax = min(amax, 0.9 * np.min((lim - pf[pos]) / dphi[pos]))
            if neg.any():
                amax = min(amax, 0.9 * np.min((-lim - pf[neg]) / dphi[neg]))
        if not np.isfinite(amax) or amax <= 0:
            amax = 1
        a = min(1.0, amax)
        best, bphi, bmu, ok = np.inf, phi, mu, False              # :394-425
        for _ls in range(12):
            pt = phi + a * dphi.reshape(phi.shape)
            mt = mu + a * dmu.reshape(mu.shape)
            Rp, Rm = residual_2d(G, pt, mt, phi0, mu0, w1, w0, dt)
            nt = np.linalg.norm(np.concatenate([Rp.ravel(), Rm.ravel()]))
            if nt < best:
                best, bphi, bmu = nt, pt, mt
            if nt <= (1.0 - 1e-4 * a) * nR:
                if floor_aware:      # nonlinear remainder of the full step = the true residual of the accepted iterate
                    lp = 2.0 / (1.0 - np.clip(phi ** 2, 0.0, 1.0 - DELTA_SEP ** 2))
                    dp = dphi.reshape(phi.shape)
                    rem = G.P.c1 * (flory_log(phi + dp, log_eps()) - flory_log(phi, log_eps()) - lp * dp)
                    true_est = float(np.linalg.norm(rem.ravel())) if a == 1.0 else np.inf
                phi, mu, ok = pt, mt, True
                break
            a *= 0.5
        if not ok and best < nR:
            phi, mu = bphi, bmu
    return phi, mu, hist


def forward_2d(P: Phys2D, u=None, phi_init=None, floor_aware=False, progress=None):
    """Time loop — Forward2_solver.py:489-596.  Returns dict(phi (M+1,..), mu (M,..), w (M,..), t, x, y, nres)."""
    G = Grid2D(P)
    phi = init_phi_2d(P.Nx, P.Ny) if phi_init is None else phi_init.copy()
    if u is not None and (u.ndim != 3 or u.shape[1:] != phi.shape):
        raise ValueError(f"control_input must have shape (M, {P.Nx+1}, {P.Ny+1})")
    w = np.zeros_like(phi)
    mu = mu_init_2d(G, phi, w)
    m0 = (G.wts_h * phi).sum()
    lim = 1.0 - DELTA_SEP
    H, MU, W, ts, nres, lastres = [phi.copy()], [], [], [0.0], [], []
    t, step = 0.0, 0
    while t < P.T - 1e-10:
        dt = min(P.dt_initial, P.T - t)
        if u is not None and step < u.shape[0] - 1:
            un, un1 = u[step], u[step + 1]
        else:
            un = un1 = np.zeros_like(phi)
        w1 = w_update(w, dt, P.gamma, un, un1)
        pn, mn, hist = newton_2d(G, phi, mu, w, w1, dt, floor_aware=floor_aware)
        if progress:
            progress(step, hist)
        phi = np.clip(pn, -lim, lim)                               # :562
        err = (G.wts_h * phi).sum() - m0                           # :565-577
        if abs(err) > 1e-16:
            inner = np.abs(phi) < lim - 5e-3
            Wi = float(G.wts_h[inner].sum())
            if Wi > 0.0:
                phi[inner] -= err / Wi
            else:
                phi = np.clip(phi - err / (P.Lx * P.Ly), -lim, lim)
        mu, w = mn, w1
        t += dt
        step += 1
        H.append(phi.copy()); MU.append(mu.copy()); W.append(w.copy()); ts.append(min(t, P.T))
        nres.append(len(hist)); lastres.append(hist[-1])
    return dict(phi=np.array(H), mu=np.array(MU), w=np.array(W), t=np.array(ts), x=G.x, y=G.y,
                nres=np.array(nres), lastres=np.array(lastres))


def targets_2d(x, y, t, phi_init, Lx, Ly, T):
    """choice 1/1 — GD2_configured.py:184, 199, 221-222."""
    xx, yy = np.meshgrid(x, y, indexing="ij")
    phiT = 0.7 * np.sin(2 * np.pi * xx / Lx) * np.cos(np.pi * yy / Ly)
    s = (t / T)[:, None, None]
    return phiT, (1 - s) * phi_init + s * phiT


# --------------------------------------------------------------------------- 2D adjoint
def adjoint_2d(P: Phys2D, phi_hist, x, y, t, b1, b2, phiQ=None, phiT=None):
    """backward2_solver.py:75-246.  Returns (p, q, r), each (M+1, Nx+1, Ny+1)."""
    assert phi_hist.ndim == 3 and t.ndim == 1 and t.shape[0] == phi_hist.shape[0]
    M1, nx1, ny1 = phi_hist.shape
    n = nx1 * ny1
    L = neumann_2d(nx1 - 1, ny1 - 1, float(x[1] - x[0]), float(y[1] - y[0]))
    L2 = (L @ L).tocsr()
    I = sp.eye(n, format="csr")
    F = phi_hist.reshape(M1, n)
    Q = np.zeros_like(F) if phiQ is None else phiQ.reshape(M1, n)
    Tt = np.zeros(n) if phiT is None else phiT.reshape(n)
    p, q, r = np.zeros((M1, n)), np.zeros((M1, n)), np.zeros((M1, n))
    p[-1] = spsolve((I - P.tau * L).tocsc(), b2 * (F[-1] - Tt))                    # :183-187
    q[-1] = -(L @ p[-1])
    for k in range(M1 - 2, -1, -1):
        dt = float(t[k + 1] - t[k])
        if dt <= 1e-14:                                                            # :214-216
            p[k], q[k], r[k] = p[k + 1], q[k + 1], r[k + 1]
            continue
        src = 0.5 * dt * b1 * ((F[k] - Q[k]) + (F[k + 1] - Q[k + 1]))              # :222-224
        Bm = I - P.tau * L - 0.5 * dt * L2 + 0.5 * dt * (sp.diags(fpp(F[k + 1], P.c1, P.c2)) @ L)
        Am = I - P.tau * L + 0.5 * dt * L2 - 0.5 * dt * (sp.diags(fpp(F[k], P.c1, P.c2)) @ L)
        p[k] = spsolve(Am.tocsc(), Bm @ p[k + 1] + src)                            # :226-229
        q[k] = -(L @ p[k])
        den = P.gamma + 0.5 * dt                                                   # :239-242
        r[k] = (P.gamma - 0.5 * dt) / den * r[k + 1] + 0.5 * dt / den * (q[k] + q[k + 1])
    sh = (M1, nx1, ny1)
    return p.reshape(sh), q.reshape(sh), r.reshape(sh)


def _trapz(f, x, axis=-1):
    """np.trapz semantics (sum of 0.5*dx*(f_i+f_{i+1})) without depending on the deprecated alias."""
    f = np.moveaxis(np.asarray(f), axis, -1)
    return (0.5 * np.diff(x) * (f[..., 1:] + f[..., :-1])).sum(-1)


def cost_2d(phi, u, phiQ, phiT, x, y, t, O: Opt2D):
    """J1..J4 — cost2_and_function.py:80-108.  Returns (J, [J1,J2,J3,J4])."""
    sp2 = lambda a: _trapz(_trapz(a, y, -1), x, -1)
    J1 = 0.5 * O.b1 * _trapz(sp2((phi - phiQ) ** 2), t)
    J2 = 0.5 * O.b2 * sp2((phi[-1] - phiT) ** 2)
    J3 = 0.5 * O.b3 * _trapz(sp2(u ** 2), t)
    J4 = O.kappa_sparsity * _trapz(sp2(np.abs(u)), t)
    return float(J1 + J2 + J3 + J4), np.array([J1, J2, J3, J4], dtype=float)


def pgd_iter_2d(P: Phys2D, O: Opt2D, u, phi_hist, t, x, y, phiQ, phiT, alpha):
    """One optimistic iteration — GD2_configured.py:299-313.  Returns (u_new, fwd dict, J, r)."""
    _, _, r = adjoint_2d(P, phi_hist, x, y, t, O.b1, O.b2, phiQ, phiT)
    g = r + O.b3 * u                                               # cost2_and_function.py:150
    un = soft_prox(u, g, alpha, O.kappa_sparsity, O.u_min, O.u_max)
    fw = forward_2d(P, un)
    J, _ = cost_2d(fw["phi"], un, phiQ, phiT, x, y, fw["t"], O)
    return un, fw, J, r


# --------------------------------------------------------------------------- 1D
def init_phi_1d(N, delta_sep=DELTA_SEP, amp=0.01, seed=42):
    """Forward_solver.py:264-277."""
    phi = amp * np.random.default_rng(seed).standard_normal(N + 1)
    w = trapz_w(N + 1)
    phi -= np.dot(w, phi) / w.sum()
    return np.clip(phi, -1 + delta_sep, 1 - delta_sep)


def residual_1d(P: Phys1D, L, phi, mu, phi0, mu0, w1, w0, dt):
    """Forward_solver.py:93-109."""
    Rp = (P.tau * (phi - phi0) / dt - 0.5 * P.kappa * (L @ phi + L @ phi0)
          + (P.c1 * flory_log(phi, log_eps()) - 2.0 * P.c2 * phi0) - 0.5 * (mu + mu0) - 0.5 * (w1 + w0))
    Rm = (phi - phi0) / dt - 0.5 * (L @ mu + L @ mu0)
    return Rp, Rm


def newton_1d(P: Phys1D, L, phi0, mu0, w0, w1, dt, tol=1e-6, max_iter=50):
    """Forward_solver.py:139-235 (guess mu=mu_old; ceiling 0.9*alpha_max; eta=1e-3; trial must stay
    strictly inside ±(1-delta_sep); a failed line search returns at once)."""
    n = phi0.size
    phi, mu = phi0.copy(), mu0.copy()
    lim = 1.0 - DELTA_SEP
    hist = []
    I = np.eye(n)
    for _ in range(max_iter):
        Rp, Rm = residual_1d(P, L, phi, mu, phi0, mu0, w1, w0, dt)
        R = np.concatenate([Rp, Rm])
        nR = np.linalg.norm(R)
        hist.append(nR)
        if nR < tol:
            break
        J = np.zeros((2 * n, 2 * n))                               # :111-137 (no clip of phi^2 in 1D)
        Kpp = -0.5 * P.kappa * L
        np.fill_diagonal(Kpp, np.diag(Kpp) + P.tau / dt + 2.0 * P.c1 / (1.0 - phi ** 2))
        J[:n, :n] = Kpp
        J[:n, n:] = -0.5 * I
        J[n:, :n] = I / dt
        J[n:, n:] = -0.5 * L
        d = np.linalg.solve(J, -R)
        dphi, dmu = d[:n], d[n:]
        with np.errstate(divide="ignore", invalid="ignore"):      # :194-212
            pos, neg = dphi > 0, dphi < 0
            ap = np.min((lim - phi[pos]) / dphi[pos]) if pos.any() else np.inf
            an = np.min((-lim - phi[neg]) / dphi[neg]) if neg.any() else np.inf
        amax = min(ap, an)
        if not np.isfinite(amax) or amax <= 0:
            amax = 1.0
        a = min(1.0, 0.9 * amax)
        for _ls in range(12):                                      # :216-229
            pt, mt = phi + a * dphi, mu + a * dmu
            if np.all(np.abs(pt) < lim):
                Rp, Rm = residual_1d(P, L, pt, mt, phi0, mu0, w1, w0, dt)
                if np.linalg.norm(np.concatenate([Rp, Rm])) <= (1 - 1e-3 * a) * nR:
                    phi, mu = pt, mt
                    break
            a *= 0.5
        else:
            return phi, mu, hist
    return phi, mu, hist


def forward_1d(P: Phys1D, u=None, phi_init=None):
    """Forward_solver.py:286-386: history starts with phi0 twice (t = [0, 0, dt, ...]); step s reads control
    rows s and s+1 (last row repeated past the end); uniform mass shift mass_err/Lx with no re-clip."""
    N = P.N
    h = P.Lx / N
    x = np.linspace(0, P.Lx, N + 1)
    phi = phi_init.copy() if (phi_init is not None and phi_init.shape == (N + 1,)) else init_phi_1d(N)
    L = neumann_1d(N, h).toarray()
    wts_h = h * trapz_w(N + 1)
    m0 = np.dot(wts_h, phi)
    w = np.zeros(N + 1)
    mu = -P.kappa * (L @ phi) + P.c1 * flory_log(phi, log_eps()) - 2.0 * P.c2 * phi - w      # :82-86
    H, MU, W, ts, nres = [phi.copy(), phi.copy()], [], [], [0.0, 0.0], []
    t, step = 0.0, 0
    lim = 1.0 - DELTA_SEP
    while t < P.T - 1e-10:
        dt = min(P.dt_initial, P.T - t)
        if u is not None:
            if step < u.shape[0] - 1:
                un, un1 = u[step], u[step + 1]
            else:
                un = un1 = u[step]
        else:
            un = un1 = np.zeros(N + 1)
        w1 = w_update(w, dt, P.gamma, un, un1)
        pn, mn, hist = newton_1d(P, L, phi, mu, w, w1, dt)
        phi = np.clip(pn, -lim, lim)
        mu, w = mn, w1
        phi = phi - (np.dot(wts_h, phi) - m0) / P.Lx               # :364-366
        t += dt
        step += 1
        H.append(phi.copy()); MU.append(mu.copy()); W.append(w.copy()); ts.append(min(t, P.T)); nres.append(len(hist))
    return dict(phi=np.array(H), mu=np.array(MU), w=np.array(W), t=np.array(ts), x=x, nres=np.array(nres))


def adjoint_1d(phi_hist, x, t, b1, b2, phiQ=None, phiT=None, P: Phys1D | None = None):
    """backward_solver.py:48-125.  The reference hard-wires the default physics (module globals, :29-33);
    P=None reproduces that.  Steps with dt<=0 are skipped, leaving zeros (:110)."""
    P = P or Phys1D()
    M1, n = phi_hist.shape
    Q = np.zeros_like(phi_hist) if phiQ is None else phiQ
    Tt = np.zeros(n) if phiT is None else phiT
    h = x[1] - x[0]
    L = neumann_1d(n - 1, h).toarray()
    L2 = L @ L
    I = np.eye(n)
    p, q, r = np.zeros_like(phi_hist), np.zeros_like(phi_hist), np.zeros_like(phi_hist)
    p[-1] = np.linalg.solve(I - P.tau * L, b2 * (phi_hist[-1] - Tt))
    q[-1] = -(L @ p[-1])
    for k in range(M1 - 2, -1, -1):
        dt = t[k + 1] - t[k]
        if dt <= 0:
            continue
        src = 0.5 * dt * b1 * ((phi_hist[k] - Q[k]) + (phi_hist[k + 1] - Q[k + 1]))
        Bm = I - P.tau * L - 0.5 * dt * L2 + 0.5 * dt * (fpp(phi_hist[k + 1], P.c1, P.c2)[:, None] * L)
        Am = I - P.tau * L + 0.5 * dt * L2 - 0.5 * dt * (fpp(phi_hist[k], P.c1, P.c2)[:, None] * L)
        p[k] = np.linalg.solve(Am, Bm @ p[k + 1] + src)
        q[k] = -(L @ p[k])
        r[k] = (P.gamma - 0.5 * dt) / (P.gamma + 0.5 * dt) * r[k + 1] + 0.5 * dt / (P.gamma + 0.5 * dt) * (q[k] + q[k + 1])
    return p, q, r


def cost_1d(phi, u, phiQ, phiT, x, t, b1, b2, b3, kappa_sp):
    """cost_and_function.py:55-75."""
    J1 = 0.5 * b1 * _trapz(_trapz((phi - phiQ) ** 2, x, 1), t)
    J2 = 0.5 * b2 * _trapz((phi[-1] - phiT) ** 2, x)
    J3 = 0.5 * b3 * _trapz(_trapz(u ** 2, x, 1), t)
    J4 = kappa_sp * _trapz(_trapz(np.abs(u), x, 1), t)
    return float(J1 + J2 + J3 + J4), np.array([J1, J2, J3, J4], dtype=float)


def targets_1d(x, t, phi_init, Lx, A_T=0.7, choice_t=1):
    """GD_1D.py:190-248 (choice_q = 1: ramp over t/t[-1])."""
    if choice_t == 1:
        phiT = A_T * np.sin(2.0 * np.pi * x / Lx)
    elif choice_t == 2:
        phiT = A_T * np.cos(2.0 * np.pi * x / Lx)
    else:
        raw = np.tan(2.0 * np.pi * 0.45 * (x / Lx - 0.5))
        sc = np.max(np.abs(raw))
        phiT = A_T * raw / (sc if sc > 1e-12 else 1.0)
    s = (t / (t[-1] if t[-1] > 0 else 1.0))[:, None]
    return phiT, (1.0 - s) * phi_init + s * phiT


def pgd_iter_1d(P: Phys1D, O: Opt1D, u, phi_hist, t, x, phiQ, phiT, alpha):
    """GD_1D.py:359-376.  Adjoint uses default physics exactly like the reference."""
    _, _, r = adjoint_1d(phi_hist, x, t, O.b1, O.b2, phiQ, phiT)
    un = soft_prox(u, r + O.b3 * u, alpha, O.kappa_sparsity, O.u_min, O.u_max)
    fw = forward_1d(P, un)
    J, _ = cost_1d(fw["phi"], un, phiQ, phiT, x, t, O.b1, O.b2, O.b3, O.kappa_sparsity)
    return un, fw, J, r


# --------------------------------------------------------------------------- large-grid helpers for bench cpu_baseline
def time_step_sample_2d(P: Phys2D, n_steps=1, u=None):
    """Run n_steps forward steps + n_steps adjoint steps of the reference algorithm (SuperLU solves) and
    return wall seconds per (forward step, adjoint step, newton solves per step).  Used only by bench.py."""
    import time
    G = Grid2D(P)
    phi = init_phi_2d(P.Nx, P.Ny)
    w = np.zeros_like(phi)
    mu = mu_init_2d(G, phi, w)
    dt = P.dt_initial
    t0 = time.perf_counter()
    nsolve = 0
    hist_phi = [phi.copy()]
    for s in range(n_steps):
        w1 = w_update(w, dt, P.gamma, np.zeros_like(phi) if u is None else u[s], np.zeros_like(phi) if u is None else u[s + 1])
        pn, mn, hist = newton_2d(G, phi, mu, w, w1, dt)
        nsolve += len(hist) - 1
        phi, mu, w = np.clip(pn, -0.99, 0.99), mn, w1
        hist_phi.append(phi.copy())
    t_fwd = (time.perf_counter() - t0) / n_steps
    F = np.array(hist_phi)
    tt = dt * np.arange(n_steps + 1)
    t0 = time.perf_counter()
    adjoint_2d(P, F, G.x, G.y, tt, 5.0, 10.0)
    t_adj = (time.perf_counter() - t0) / n_steps          # includes the terminal solve (amortised over n_steps)
    return t_fwd, t_adj, nsolve / n_steps
