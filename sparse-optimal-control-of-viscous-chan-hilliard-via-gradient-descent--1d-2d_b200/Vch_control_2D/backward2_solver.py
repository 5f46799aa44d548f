"""Adjoint (backward-in-time) sweep of the 2D problem — B200 drop-in for 2D/Vch_control_2D/backward2_solver.py.

Same recurrence as the reference (:183-242): terminal solve (I - tau L) p_M = b2 (phi_M - phi_Omega), then for each
interval  A(phi_n) p_n = B(phi_{n+1}) p_{n+1} + src,  q_n = -L p_n,  and the Crank–Nicolson filter for r_n.  The sparse
LU solves become DCT-preconditioned BiCGStab solves on the device; the whole sweep is one C-ABI call.
"""
import os
import sys
from typing import Optional, Tuple

import numpy as np

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)
import vch_b200_native as _nat                     # noqa: E402
from config import ForwardSolverConfig             # noqa: E402
from Forward2_solver import laplacian_matrix_neumann   # noqa: E402,F401  (re-exported like the reference)


def fpp_log(phi, c1, c2, eps=1e-8):
    """f''(phi) = 2 c1 / (1 - s^2) - 2 c2 with s = clip(phi, ±(1-eps)) (reference :40-72); host helper for callers."""
    s = np.clip(phi, -1.0 + eps, 1.0 - eps)
    return 2.0 * c1 / (1.0 - s * s) - 2.0 * c2


def run_backward(phi_hist: np.ndarray, x: np.ndarray, y: np.ndarray, t_hist: np.ndarray, config: ForwardSolverConfig,
                 b1: float, b2: float, phi_Q: Optional[np.ndarray] = None,
                 phi_T_target: Optional[np.ndarray] = None) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Returns (p, q, r), each (M+1, Nx+1, Ny+1).  Shape errors are AssertionErrors like the reference (:141-145)."""
    assert phi_hist.ndim == 3, "phi_hist must be (M+1, Nx+1, Ny+1)"
    levels, nx1, ny1 = phi_hist.shape
    assert x.ndim == 1 and y.ndim == 1, "x and y must be 1D arrays"
    assert x.size >= 2 and y.size >= 2, "x and y must have at least 2 points"
    assert t_hist.ndim == 1 and t_hist.shape[0] == levels, "t_hist must align with phi_hist"
    Nx, Ny = nx1 - 1, ny1 - 1
    hx, hy = float(x[1] - x[0]), float(y[1] - y[0])
    ctx = _nat.ctx2d(Nx, Ny, hx, hy, Nx * hx, Ny * hy, float(config.tau), float(config.gamma), float(config.c1),
                     float(config.c2), float(config.kappa), 1e-2)
    f64 = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float64)
    Q = None if phi_Q is None else f64(phi_Q).reshape(phi_hist.shape)
    T = None if phi_T_target is None else f64(phi_T_target).reshape(nx1, ny1)
    return ctx.adjoint(f64(phi_hist), f64(t_hist), float(b1), float(b2), Q, T)
