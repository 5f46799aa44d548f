"""Forward solve of the 2D viscous Cahn–Hilliard state system — B200 drop-in for the reference module of the same
name (2D/Vch_control_2D/Forward2_solver.py).  Every function keeps the reference's name, argument order and return
structure (NumPy fp64 in, NumPy fp64 out); the arithmetic runs in the sm_100a kernels of libvch_b200.so through the
C ABI (include/vch_b200.h).  There is no CPU path: without the library or a CUDA device the compute calls raise.

What still runs on the host, by design: the RNG of the initial condition (the reference draws it from NumPy's PCG64,
:455-456, and parity needs the same stream), assembly of the SciPy operator objects that the reference's tests inspect
(`laplacian_matrix_neumann`, `assemble_jacobian` — the solver itself is matrix-free and never uses them), and the
scalar diagnostic `instability_report` (`free_energy` is a fused device reduction).
"""
import os
import sys

import numpy as np
import scipy.sparse as sps

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)
import vch_b200_native as _nat                                    # noqa: E402
from config import ForwardSolverConfig, load_params, get_user_input_for_config   # noqa: E402,F401

DEBUG = True
COMPUTE_ENERGY = False
ENERGY_EVERY_N_STEPS = 100
_DELTA_SEP = 1e-2            # hard-wired in the reference's time loop (:510)


# ----------------------------------------------------------------------------- host-side helpers kept for API parity
def instability_report(c1, c2, kappa, tau, Lx, Nmodes=12):
    """Growth rates of the linearised problem for the first Nmodes cosine modes (reference :53-83)."""
    q = (np.pi * np.arange(1, Nmodes + 1) / Lx) ** 2
    a = 2 * (c1 - c2)
    lam = (-kappa * q ** 2 - a * q) / (1 + tau * q)
    print(f"a={a:.3g},  max λ={lam.max():.3g} at mode n={lam.argmax()+1},  unstable modes={(lam>0).sum()}")
    return lam


def regularized_log(phi, delta_sep):
    """log((1+s)/(1-s)) with s = clip(phi, ±(1-eps)), eps = max(1e-8, delta_sep/2) (reference :86-102).
    Host helper for callers/tests; the kernels evaluate the same expression on the device."""
    eps = max(1e-8, 0.5 * delta_sep)
    s = np.clip(phi, -1.0 + eps, 1.0 - eps)
    return np.log((1.0 + s) / (1.0 - s))


def trapz_weights(n_nodes):
    w = np.ones(int(n_nodes))
    w[0] = w[-1] = 0.5
    return w


def laplacian_matrix_neumann_1d(N, h):
    """(N+1)x(N+1) second-difference matrix with mirrored ghost nodes, CSR (reference :105-122)."""
    a = 1.0 / (h * h)
    off = np.full(N, a)
    up, lo = off.copy(), off.copy()
    up[0] = 2.0 * a
    lo[-1] = 2.0 * a
    return sps.diags([lo, np.full(N + 1, -2.0 * a), up], [-1, 0, 1], format="csr")


def laplacian_matrix_neumann(Nx, Ny, hx, hy):
    """kron(I_{Ny+1}, L_x) + kron(L_y, I_{Nx+1}) as SciPy CSR (reference :125-137).  The device stencils implement
    exactly this operator (same factor order); this matrix exists for callers that want to inspect or multiply it."""
    L = (sps.kron(sps.eye(Ny + 1, format="csr"), laplacian_matrix_neumann_1d(Nx, hx))
         + sps.kron(laplacian_matrix_neumann_1d(Ny, hy), sps.eye(Nx + 1, format="csr"))).tocsr()
    L._vch_spacing = (float(hx), float(hy))
    return L


def _spacing(L, Nx, Ny):
    """Recover (hx, hy) from a Laplacian built by laplacian_matrix_neumann (entry (0,1) = 2/hx^2, (0,Nx+1) = 2/hy^2).

    The kernels implement exactly that mirror-ghost 5-point operator, so any OTHER matrix handed in through the reference's
    `L` argument (scaled, regularised, periodic, built for another grid) cannot be honoured: it is rejected with ValueError
    instead of being silently replaced (the reference would compute `L @ v` with whatever it is given)."""
    sp = getattr(L, "_vch_spacing", None)
    n = (Nx + 1) * (Ny + 1)
    if sp is not None and L.shape == (n, n):
        return sp
    L = L.tocsr() if sps.issparse(L) else sps.csr_matrix(L)
    if L.shape != (n, n):
        raise ValueError(f"L must be the ({n}, {n}) Neumann Laplacian of the ({Nx+1}, {Ny+1}) grid, got {L.shape}")
    a, b = float(L[0, 1]), float(L[0, Nx + 1])
    if not (a > 0.0 and b > 0.0 and np.isfinite(a) and np.isfinite(b)):
        raise ValueError("L is not the mirror-ghost Neumann Laplacian this library implements (corner row)")
    ax, ay = 0.5 * a, 0.5 * b                                  # 1/hx^2, 1/hy^2
    i = (Nx + 1) + 1                                            # an interior node next to the corner (Nx, Ny >= 2)
    ok = abs(float(L[0, 0]) + 2.0 * (ax + ay)) <= 1e-9 * (ax + ay)
    if Nx >= 2 and Ny >= 2:
        ok = ok and abs(float(L[i, i]) + 2.0 * (ax + ay)) <= 1e-9 * (ax + ay) and abs(float(L[i, i + 1]) - ax) <= 1e-9 * ax \
            and abs(float(L[i, i - 1]) - ax) <= 1e-9 * ax and abs(float(L[i, i + Nx + 1]) - ay) <= 1e-9 * ay
        ok = ok and L.nnz <= 5 * n
    if not ok:
        raise ValueError("L is not the mirror-ghost Neumann 5-point Laplacian this library implements; the CUDA kernels "
                         "cannot apply an arbitrary operator")
    return float(np.sqrt(1.0 / ax)), float(np.sqrt(1.0 / ay))


def _ctx(Nx, Ny, hx, hy, tau=0.05, gamma=10.0, c1=0.75, c2=1.0, kappa=1e-4, delta_sep=_DELTA_SEP):
    return _nat.ctx2d(Nx, Ny, hx, hy, Nx * hx, Ny * hy, tau, gamma, c1, c2, kappa, delta_sep)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


# ----------------------------------------------------------------------------- device-backed building blocks
def apply_laplacian(L, v, Nx, Ny):
    if v.ndim != 2 or v.shape != (Nx + 1, Ny + 1):
        raise ValueError(f"Input field must have shape ({Nx+1}, {Ny+1})")
    hx, hy = _spacing(L, Nx, Ny)
    return _ctx(Nx, Ny, hx, hy).apply_laplacian(_f64(v))


def initialize_mu(phi, w, c1, c2, kappa, L, Nx, Ny, delta_sep):
    hx, hy = _spacing(L, Nx, Ny)
    return _ctx(Nx, Ny, hx, hy, c1=c1, c2=c2, kappa=kappa, delta_sep=delta_sep).initialize_mu(_f64(phi), _f64(w))


def solve_w(w_old, dt, gamma, u_n, u_np1):
    return _nat.solve_w(_f64(w_old), float(dt), float(gamma), _f64(u_n), _f64(u_np1))


def _residuals(phi_new, phi_old, mu_new, mu_old, w_new, w_old, dt, tau, c1, c2, kappa, L, Nx, Ny, delta_sep):
    hx, hy = _spacing(L, Nx, Ny)
    c = _ctx(Nx, Ny, hx, hy, tau=tau, c1=c1, c2=c2, kappa=kappa, delta_sep=delta_sep)
    return c.residual(_f64(phi_new), _f64(phi_old), _f64(mu_new), _f64(mu_old), _f64(w_new), _f64(w_old), float(dt))


def solve_mu_residual(phi_new, phi_old, mu_new, mu_old, dt, L, Nx, Ny):
    z = np.zeros_like(phi_new, dtype=np.float64)
    return _residuals(phi_new, phi_old, mu_new, mu_old, z, z, dt, 0.05, 0.75, 1.0, 1e-4, L, Nx, Ny, _DELTA_SEP)[1]


def solve_phi_residual(phi_new, phi_old, mu_new, mu_old, w_new, w_old, dt, tau, c1, c2, kappa, L, Nx, Ny, delta_sep):
    return _residuals(phi_new, phi_old, mu_new, mu_old, w_new, w_old, dt, tau, c1, c2, kappa, L, Nx, Ny, delta_sep)[0]


def assemble_jacobian(phi_new, dt, tau, c1, kappa, L, delta_sep):
    """Block Jacobian as SciPy CSR for inspection (reference :224-253).  The device Newton solve is matrix-free
    (Schur reduction + DCT-preconditioned BiCGStab) and never forms this matrix."""
    f = np.asarray(phi_new, dtype=np.float64).ravel()
    n = f.size
    diag = tau / dt + 2.0 * c1 / (1.0 - np.clip(f * f, 0.0, 1.0 - delta_sep ** 2))
    Lc = L.tocsr()
    eye = sps.eye(n, format="csr")
    return sps.bmat([[sps.diags(diag) - 0.5 * kappa * Lc, -0.5 * eye], [eye / dt, -0.5 * Lc]], format="csr")


def free_energy(phi, kappa, c1, c2, hx, hy, w=None, eps=None):
    """Discrete Ginzburg–Landau/Flory–Huggins energy (reference :256-319): one fused reduction kernel (vch_free_energy)."""
    f = _f64(phi)
    if f.ndim != 2:
        raise ValueError("phi must be a 2D array of shape (Ny+1, Nx+1)")
    return _nat.free_energy(f, float(kappa), float(c1), float(c2), float(hx), float(hy),
                            None if w is None else _f64(w), 1e-8 if eps is None else float(eps))


def newton_raphson(phi_old, mu_old, w_old, w_new, dt, tau, c1, c2, kappa, delta_sep, L, Nx, Ny, hx, hy,
                   return_residual_history=False):
    """One implicit step (phi_{n+1}, mu_{n+1}) — reference :323-427: same initial guess, ||R||_2 < 1e-6 stop, step
    ceiling, Armijo rule and best-trial fallback; the sparse LU is replaced by the matrix-free Krylov solve."""
    c = _ctx(Nx, Ny, hx, hy, tau=tau, c1=c1, c2=c2, kappa=kappa, delta_sep=delta_sep)
    phi_new, mu_new, hist = c.newton(_f64(phi_old), _f64(mu_old), _f64(w_old), _f64(w_new), float(dt))
    return (phi_new, mu_new, hist) if return_residual_history else (phi_new, mu_new)


def init_phi_random(Nx, Ny, delta_sep, amp=0.5, seed=42, enforce_zero_mean=True):
    """Random initial phase field with zero trapezoid-weighted mean (reference :444-486).  Host RNG on purpose."""
    phi = amp * np.random.default_rng(seed).standard_normal((Nx + 1, Ny + 1))
    wts = np.outer(trapz_weights(Nx + 1), trapz_weights(Ny + 1))
    total = np.sum(wts)
    lo, hi = -1.0 + delta_sep, 1.0 - delta_sep
    if enforce_zero_mean:
        phi -= np.sum(wts * phi) / total
    phi = np.clip(phi, lo, hi)
    if enforce_zero_mean:
        for _ in range(8):
            defect = np.sum(wts * phi)
            if abs(defect) <= 1e-14 * total:
                break
            free = np.abs(phi) < (hi - 5e-3)
            wfree = float(np.sum(wts[free]))
            if wfree <= 0:
                phi = np.clip(phi - defect / total, lo, hi)
                break
            phi[free] -= defect / wfree
    return phi


def _time_grid(T, dt):
    """The reference's while-loop (:539-585) in isolation: per-step dt and the stored time stamps."""
    steps, stamps, t = [], [0.0], 0.0
    while t < T - 1e-10:
        d = min(dt, T - t)
        steps.append(d)
        t += d
        stamps.append(min(t, T))
    return np.array(steps), np.array(stamps)


def run_main_simulation(config, store_history=False, control_input=None, verbose=True):
    """Time-march phi from t = 0 to T under an optional control (reference :489-608).

    Returns (phi_hist (M+1, Nx+1, Ny+1), (x, y), t_hist (M+1,)) when store_history, otherwise shows the final field
    (if matplotlib is installed) and returns None — as the reference does."""
    Nx, Ny = int(config.Nx), int(config.Ny)
    Lx, Ly = float(config.Lx), float(config.Ly)
    hx, hy = Lx / Nx, Ly / Ny
    x, y = np.linspace(0.0, Lx, Nx + 1), np.linspace(0.0, Ly, Ny + 1)
    phi0 = init_phi_random(Nx, Ny, _DELTA_SEP, amp=0.1, seed=42)        # module attribute, resolved at call time
    if control_input is not None and (control_input.ndim != 3 or control_input.shape[1:] != phi0.shape):
        raise ValueError(f"control_input must have shape (M, {Nx+1}, {Ny+1})")
    dts, t_hist = _time_grid(float(config.T), float(config.dt_initial))
    ctx = _nat.ctx2d(Nx, Ny, hx, hy, Lx, Ly, float(config.tau), float(config.gamma), float(config.c1), float(config.c2),
                     float(config.kappa), _DELTA_SEP)
    phi_hist, _, _ = ctx.forward(_f64(phi0), None if control_input is None else _f64(control_input), dts)
    if verbose:
        t = 0.0
        for k, d in enumerate(dts, start=1):
            t += d
            if k % 10 == 0 or t >= float(config.T):
                print(f"Step {k:5d} | t={t:.4e} | ||phi||_inf={np.max(np.abs(phi_hist[k])):.5f}")
        print("Simulation complete.")
    if store_history:
        return phi_hist, (x, y), t_hist
    try:
        import matplotlib.pyplot as plt
        plt.figure(figsize=(6, 5))
        plt.imshow(phi_hist[-1].T, origin="lower", extent=[x[0], x[-1], y[0], y[-1]], vmin=-1.0, vmax=1.0, cmap="RdBu_r")
        plt.title(f"Final Profile of φ at t={float(config.T)}")
        plt.colorbar(label="φ")
        plt.show()
    except ImportError:
        pass
    return None


if __name__ == "__main__":
    run_main_simulation(config=get_user_input_for_config(ForwardSolverConfig, "Forward Solver Parameters",
                                                         load_params().forward_solver), store_history=False, verbose=True)
