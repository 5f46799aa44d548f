"""Forward solve of the 1D viscous Cahn–Hilliard state system — B200 drop-in for 1D/Vch_control_1D/Forward_solver.py.
Same function names, argument orders and return structures; the time loop, Newton iteration, line search and the
(banded) linear solves run inside ONE CUDA kernel per call (one CTA per problem, state resident in shared memory).
No CPU path.  Host-side by design: the NumPy RNG of the initial condition, the dense Laplacian object the reference's
callers/tests multiply with, and scalar diagnostics."""
import os
import sys

import numpy as np

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)
import vch_b200_native as _nat                # noqa: E402
from config import ForwardSolverConfig        # noqa: E402

delta_sep = 1e-2          # module constant of the reference (:42)
DEBUG = True
COMPUTE_ENERGY = True

_f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)


def instability_report(c1, c2, kappa, tau, Lx, Nmodes=12):
    q = (np.pi * np.arange(1, Nmodes + 1) / Lx) ** 2
    a = 2 * (c1 - c2)
    lam = (-kappa * q ** 2 - a * q) / (1 + tau * q)
    print(f"a={a:.3g},  max λ={lam.max():.3g} at mode n={lam.argmax()+1},  unstable modes={(lam>0).sum()}")
    return lam


def regularized_log(phi, eps=None):
    eps = max(1e-8, 0.5 * delta_sep) if eps is None else eps
    s = np.clip(phi, -1 + eps, 1 - eps)
    return np.log((1 + s) / (1 - s))


def laplacian_matrix_neumann(N, h):
    """Dense (N+1)x(N+1) mirror-ghost second-difference matrix (reference :64-76) for callers that multiply with it;
    the device kernels apply the same stencil matrix-free."""
    a = 1.0 / (h * h)
    L = np.zeros((N + 1, N + 1))
    i = np.arange(N + 1)
    L[i, i] = -2 * a
    L[i[1:], i[:-1]] = a
    L[i[:-1], i[1:]] = a
    L[0, 1] = L[N, N - 1] = 2 * a
    return L


def apply_laplacian(L, v):
    return L @ v


def trapz_weights(n_nodes: int) -> np.ndarray:
    w = np.ones(n_nodes)
    w[0] = w[-1] = 0.5
    return w


def _ctx(n_nodes, h, tau=0.05, gamma=10.0, c1=0.75, c2=1.0, kappa=0.03 ** 2, dsep=None):
    N = n_nodes - 1
    return _nat.ctx1d(N, h, N * h, tau, gamma, c1, c2, kappa, delta_sep if dsep is None else dsep)


def _h_of(L):
    return float(np.sqrt(2.0 / L[0, 1]))


def initialize_mu(phi, w, c1, c2, L, kappa):
    """mu = -kappa L phi + c1 log((1+phi)/(1-phi)) - 2 c2 phi - w (reference :82-86), evaluated on the device."""
    return _ctx(len(phi), _h_of(L), c1=c1, c2=c2, kappa=kappa).initialize_mu(_f64(phi), _f64(w))


def solve_w(w_old, dt, gamma, u_n, u_np1):
    return _nat.solve_w(_f64(w_old), float(dt), float(gamma), _f64(u_n), _f64(u_np1))


def solve_mu_residual(phi_new, phi_old, mu_new, mu_old, dt, L):
    z = np.zeros_like(_f64(phi_new))
    return _ctx(len(phi_new), _h_of(L)).residual(_f64(phi_new), _f64(phi_old), _f64(mu_new), _f64(mu_old), z, z, float(dt))[1]


def solve_phi_residual(phi_new, phi_old, mu_new, mu_old, w_new, w_old, dt, tau, c1, c2, L, kappa):
    c = _ctx(len(phi_new), _h_of(L), tau=tau, c1=c1, c2=c2, kappa=kappa)
    return c.residual(_f64(phi_new), _f64(phi_old), _f64(mu_new), _f64(mu_old), _f64(w_new), _f64(w_old), float(dt))[0]


def assemble_jacobian(phi_new, dt, tau, c1, L, kappa):
    """Dense 2(N+1) block Jacobian for inspection (reference :111-137); the device solve uses the pentadiagonal
    Schur complement instead and never forms it."""
    n = len(phi_new)
    J = np.zeros((2 * n, 2 * n))
    J[:n, :n] = -0.5 * kappa * L + np.diag(tau / dt + 2.0 * c1 / (1.0 - np.asarray(phi_new) ** 2))
    J[:n, n:] = -0.5 * np.eye(n)
    J[n:, :n] = np.eye(n) / dt
    J[n:, n:] = -0.5 * L
    return J


def newton_raphson(phi_old, mu_old, w_old, w_new, dt, tau, c1, c2, h, delta_sep, L, kappa, return_residual_history=False):
    """One implicit step (reference :139-235): guess (phi_old, mu_old), <= 50 iterations, ||R||_2 < 1e-6, ceiling
    0.9*alpha_max, Armijo eta = 1e-3 with the strict-interior test, immediate return when the line search fails."""
    c = _ctx(len(phi_old), h, tau=tau, c1=c1, c2=c2, kappa=kappa, dsep=delta_sep)
    phi_new, mu_new, hist = c.newton(_f64(phi_old), _f64(mu_old), _f64(w_old), _f64(w_new), float(dt))
    return (phi_new, mu_new, hist) if return_residual_history else (phi_new, mu_new)


def free_energy(phi, kappa, c1, c2, h, w=None, eps=None):
    wts = trapz_weights(len(phi))
    eps = 1e-8 if eps is None else eps
    s = np.clip(phi, -1 + eps, 1 - eps)
    bulk = c1 * ((1 + s) * np.log(1 + s) + (1 - s) * np.log(1 - s)) - c2 * s ** 2
    E = kappa / (2.0 * h) * np.sum(np.diff(phi) ** 2) + h * np.dot(wts, bulk)
    if w is not None:
        E -= h * np.dot(wts, w * phi)
    return E


def init_phi_random(N, delta_sep, amp=0.1, seed=42, enforce_zero_mean=True):
    phi = amp * np.random.default_rng(seed).standard_normal(N + 1)
    if enforce_zero_mean:
        w = trapz_weights(N + 1)
        phi -= np.dot(w, phi) / w.sum()
    return np.clip(phi, -1 + delta_sep, 1 - delta_sep)


def source_u(t, x):
    return np.zeros_like(x)


def _time_grid(T, dt):
    steps, stamps, t = [], [0.0, 0.0], 0.0              # t = 0 is stored twice (reference :329-336)
    while t < T - 1e-10:
        d = min(dt, T - t)
        steps.append(d)
        t += d
        stamps.append(min(t, T))
    return np.array(steps), np.array(stamps)


def run_main_simulation_batch(fwd_config, controls, initial_phi=None):
    """B forward solves with B different controls in ONE multi-problem launch (one CTA per problem: line-search trials,
    finite-difference directions, ensembles).  controls: (B, rows, N+1).  Returns (phi_hists (B, M+2, N+1), x, t_hist);
    member b equals run_main_simulation(fwd_config, True, controls[b])[0] bit for bit."""
    cfg = ForwardSolverConfig() if fwd_config is None else fwd_config
    N, Lx = int(cfg.N), float(cfg.Lx)
    controls = _f64(controls)
    B = int(controls.shape[0])
    if initial_phi is not None and initial_phi.shape == (N + 1,):
        phi0 = _f64(initial_phi).copy()
    else:
        phi0 = init_phi_random(N, delta_sep, amp=0.01, seed=42, enforce_zero_mean=True)
    dts, t_hist = _time_grid(float(cfg.T), float(cfg.dt_initial))
    ctx = _nat.ctx1d(N, Lx / N, Lx, float(cfg.tau), float(cfg.gamma), float(cfg.c1), float(cfg.c2), float(cfg.kappa), delta_sep)
    phi_hists, _, _ = ctx.forward(np.ascontiguousarray(np.broadcast_to(phi0, (B, N + 1))), controls, dts)
    return phi_hists, np.linspace(0, Lx, N + 1), t_hist


def run_main_simulation(fwd_config=None, store_history=False, control_input=None, verbose=True, initial_phi=None):
    """Reference :286-397.  Returns (phi_hist (M+2, N+1), x, t_hist (M+2,)) when store_history (level 0 twice, t_hist =
    [0, 0, dt, ...]); otherwise (phi_final, x, t_hist) after an optional plot."""
    cfg = ForwardSolverConfig() if fwd_config is None else fwd_config
    N, Lx = int(cfg.N), float(cfg.Lx)
    h = Lx / N
    x = np.linspace(0, Lx, N + 1)
    if initial_phi is not None and initial_phi.shape == (N + 1,):
        phi0 = _f64(initial_phi).copy()
        if verbose:
            print("Using provided initial condition for phi.")
    else:
        if verbose and initial_phi is not None:
            print(f"[Warning] Provided initial_phi has incorrect shape. Expected ({N+1},), got {initial_phi.shape}. Defaulting to random.")
        phi0 = init_phi_random(N, delta_sep, amp=0.01, seed=42, enforce_zero_mean=True)
    dts, t_hist = _time_grid(float(cfg.T), float(cfg.dt_initial))
    ctx = _nat.ctx1d(N, h, Lx, float(cfg.tau), float(cfg.gamma), float(cfg.c1), float(cfg.c2), float(cfg.kappa), delta_sep)
    u = None if control_input is None else _f64(control_input)
    phi_hist, _, _ = ctx.forward(phi0, u, dts)
    if verbose:
        t = 0.0
        for k, d in enumerate(dts, start=1):
            t += d
            if k % 100 == 0 or t >= float(cfg.T):
                print(f"Step {k:5d} | t={t:.4e} | ||phi||_inf={np.max(np.abs(phi_hist[k + 1])):.5f}")
        print("Simulation complete.")
    if store_history:
        return phi_hist, x, t_hist
    try:
        import matplotlib.pyplot as plt
        plt.figure(figsize=(10, 6))
        plt.plot(x, phi_hist[-1], label=f"Final state at t={float(cfg.T)}")
        plt.title("Final Profile of φ")
        plt.legend()
        plt.show()
    except ImportError:
        pass
    return phi_hist[-1].copy(), x, t_hist


if __name__ == "__main__":
    run_main_simulation(ForwardSolverConfig(), store_history=False, verbose=True)
