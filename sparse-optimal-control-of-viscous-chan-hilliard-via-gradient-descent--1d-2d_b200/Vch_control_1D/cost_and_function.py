"""Cost functional and smooth gradient of the 1D problem — B200 drop-in for 1D/Vch_control_1D/cost_and_function.py."""
import os
import sys

import numpy as np

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)
import vch_b200_native as _nat                      # noqa: E402
from Forward_solver import run_main_simulation      # noqa: E402,F401  (the reference imports it here too)

_f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)


def calculate_cost(phi_hist, u, phi_Q_target, phi_T_target, x, t_hist, b1, b2, b3, kappa, verbose: bool = True) -> float:
    """J1..J4 by nested np.trapezoid in x then t (reference :55-75); prints the components on every call — the
    reference ignores `verbose` (:77-82) and so does this."""
    n = phi_hist.shape[1]
    h = float(x[1] - x[0])
    ctx = _nat.ctx1d(n - 1, h, (n - 1) * h, 0.05, 10.0, 0.75, 1.0, 0.03 ** 2, 1e-2)
    J = ctx.cost(_f64(phi_hist), _f64(u), _f64(phi_Q_target), _f64(phi_T_target), _f64(x), _f64(t_hist), b1, b2, b3, kappa)
    total, j1, j2, j3, j4 = (float(v) for v in J)
    print(f"  Tracking Cost (J1): {j1:.6g}")
    print(f"  Terminal Cost (J2): {j2:.6g}")
    print(f"  Control Energy (J3): {j3:.6g}")
    print(f"  Sparsity Cost (J4): {j4:.6g}")
    print("-----------------------------")
    print(f"  Total Cost: {total:.6g}")
    return total


def calculate_cost_batch(phi_hists, us, phi_Q_target, phi_T_target, x, t_hist, b1, b2, b3, kappa) -> np.ndarray:
    """Total cost of B trajectories (B, levels, N+1) against the same targets, one multi-problem launch; quiet.
    Returns J (B, 5) = (total, J1, J2, J3, J4) per member, each row equal to calculate_cost's values."""
    phi_hists = _f64(phi_hists)
    B, lv, n = phi_hists.shape
    h = float(x[1] - x[0])
    ctx = _nat.ctx1d(n - 1, h, (n - 1) * h, 0.05, 10.0, 0.75, 1.0, 0.03 ** 2, 1e-2)
    Q = np.ascontiguousarray(np.broadcast_to(_f64(phi_Q_target), (B, lv, n)))
    T = np.ascontiguousarray(np.broadcast_to(_f64(phi_T_target), (B, n)))
    return ctx.cost(phi_hists, _f64(us), Q, T, _f64(x), _f64(t_hist), b1, b2, b3, kappa)


def calculate_gradient(r, u, b3):
    """r + b3 u (reference :99)."""
    _, g, _ = _nat.grad_prox(_f64(u), _f64(r), float(b3), 0.0, 0.0, -np.inf, np.inf, want_grad=True)
    return g


def perform_gradient_step(u_current, grad_smooth, alpha):
    """u - alpha grad (reference :111): the prox kernel with zero threshold and an open box."""
    un, _, _ = _nat.grad_prox(_f64(u_current), _f64(grad_smooth), 0.0, float(alpha), 0.0, -np.inf, np.inf)
    return un
