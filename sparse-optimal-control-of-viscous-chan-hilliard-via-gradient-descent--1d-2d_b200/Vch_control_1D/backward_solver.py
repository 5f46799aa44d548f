"""Adjoint sweep of the 1D problem — B200 drop-in for 1D/Vch_control_1D/backward_solver.py.
Keeps the reference's quirks on purpose: the physics are the DEFAULT ForwardSolverConfig() values captured at import
(:29-33; the signature has no config), and time levels with dt <= 0 are skipped leaving zero rows (:110)."""
import os
import sys
from typing import Optional, Tuple

import numpy as np

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)
import vch_b200_native as _nat                        # noqa: E402
from Forward_solver import laplacian_matrix_neumann   # noqa: E402,F401
from config import ForwardSolverConfig                # noqa: E402

_cfg = ForwardSolverConfig()
c1, c2, tau, gamma = _cfg.c1, _cfg.c2, _cfg.tau, _cfg.gamma
kappa = _cfg.kappa


def fpp_log(phi: np.ndarray, eps: float = 1e-8) -> np.ndarray:
    s = np.clip(phi, -1 + eps, 1 - eps)
    return 2.0 * c1 / (1.0 - s ** 2) - 2.0 * c2


def run_backward(phi_hist: np.ndarray, x: np.ndarray, t_hist: np.ndarray, b1: float, b2: float,
                 phi_Q: Optional[np.ndarray] = None, phi_T_target: Optional[np.ndarray] = None
                 ) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """(p, q, r), each shaped like phi_hist.  One kernel launch: pentadiagonal solve per level in shared memory."""
    f64 = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float64)
    n = phi_hist.shape[1]
    h = float(x[1] - x[0])
    ctx = _nat.ctx1d(n - 1, h, (n - 1) * h, tau, gamma, c1, c2, kappa, 1e-2)
    return ctx.adjoint(f64(phi_hist), f64(t_hist), float(b1), float(b2), f64(phi_Q), f64(phi_T_target))
