"""Cost functional, smooth gradient and proximal step of the 2D problem — B200 drop-in for
2D/Vch_control_2D/cost2_and_function.py.  Each call is one fused reduction / elementwise kernel behind the C ABI."""
import os
import sys

import numpy as np

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)
import vch_b200_native as _nat          # noqa: E402
from config import OptimizationConfig   # noqa: E402

_f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)


def calculate_cost(phi_hist, u, phi_Q_target, phi_T_target, x, y, t_hist, opt_config: OptimizationConfig) -> float:
    """J = b1/2 |phi - phi_Q|^2 + b2/2 |phi(T) - phi_T|^2 + b3/2 |u|^2 + kappa |u|_1 with nested trapezoid weights in
    y, x, t (reference :80-108).  Prints the four components on every call, as the reference does (:113-118)."""
    nx1, ny1 = phi_hist.shape[1:]
    hx, hy = float(x[1] - x[0]), float(y[1] - y[0])
    ctx = _nat.ctx2d(nx1 - 1, ny1 - 1, hx, hy, (nx1 - 1) * hx, (ny1 - 1) * hy, 0.05, 10.0, 0.75, 1.0, 1e-4, 1e-2)
    J = ctx.cost(_f64(phi_hist), _f64(u), _f64(phi_Q_target), _f64(phi_T_target), _f64(x), _f64(y), _f64(t_hist),
                 float(opt_config.b1), float(opt_config.b2), float(opt_config.b3), float(opt_config.kappa_sparsity))
    total, c1, c2, c3, c4 = (float(v) for v in J[:5])
    print(f"  Tracking Cost (J1): {c1:.6g}")
    print(f"  Terminal Cost (J2): {c2:.6g}")
    print(f"  Control Energy (J3): {c3:.6g}")
    print(f"  Sparsity Cost (J4): {c4:.6g}")
    print("  -----------------------------")
    print(f"  Total Cost:         {total:.6g}")
    return total


def calculate_gradient(r, u, opt_config: OptimizationConfig):
    """grad = r + b3 u (reference :150)."""
    _, g, _ = _nat.grad_prox(_f64(u), _f64(r), float(opt_config.b3), 0.0, 0.0, -np.inf, np.inf, want_grad=True)
    return g


def proximal_step(u_current, grad_smooth, alpha, opt_config: OptimizationConfig):
    """clip(soft_threshold(u - alpha grad, alpha kappa), u_min, u_max) (reference :191-200).  The kernel evaluates
    r + b3 u with r := grad and b3 := 0, i.e. exactly u - alpha*grad in NumPy's operation order."""
    un, _, _ = _nat.grad_prox(_f64(u_current), _f64(grad_smooth), 0.0, float(alpha), float(opt_config.kappa_sparsity),
                              float(opt_config.u_min), float(opt_config.u_max))
    return un
