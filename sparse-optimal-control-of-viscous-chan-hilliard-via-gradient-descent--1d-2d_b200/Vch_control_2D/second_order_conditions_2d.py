"""KKT / second-order diagnostics of the 2D problem — drop-in for 2D/Vch_control_2D/second_order_conditions_2d.py.

`verify_sparsity_condition` (reference :238-297) is the KKT reduction on the hot path: its three counts come from one
fused device kernel (vch_kkt_counts).  `approximate_second_order_condition_2d` (reference :120-235) re-runs the forward solve and the
cost for perturbed controls; the directions are independent problems and can run concurrently (batch / VCH_FD_BATCH).
"""
import os
import sys
from typing import List, Optional

import numpy as np

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)
import vch_b200_native as _nat                                   # noqa: E402
from config import ForwardSolverConfig, OptimizationConfig       # noqa: E402
from Forward2_solver import run_main_simulation                   # noqa: E402
from cost2_and_function import calculate_cost                     # noqa: E402


def _cone_direction(u_star, u_min, u_max, rng, tol=1e-8):
    """Unit random direction that points inward where u* sits on a bound (critical cone of the box)."""
    v = rng.standard_normal(size=u_star.shape)
    at_lo, at_hi = u_star <= u_min + tol, u_star >= u_max - tol
    v[at_lo] = np.abs(v[at_lo])
    v[at_hi] = -np.abs(v[at_hi])
    nv = np.linalg.norm(v)
    if nv < 1e-12:
        v = np.zeros_like(v)
        v.ravel()[0] = nv = 1.0
    return v / nv


def approximate_second_order_condition_2d(u_star, r_star, phi_star, x, y, t_hist,
                                          opt_config: Optional[OptimizationConfig] = None, b1=None, b2=None, b3=None,
                                          kappa=None, phi_Q_target=None, phi_T_target=None, u_min=-np.inf, u_max=np.inf,
                                          num_directions: int = 10, epsilon: float = 1e-4, seed: Optional[int] = None,
                                          fwd_config: Optional[ForwardSolverConfig] = None, batch: Optional[int] = None) -> List[float]:
    """(J(u*+eps h) - J(u*) - eps <r*+b3 u*, h>) / (eps^2/2) along random critical-cone directions h.
    batch > 1 (argument, or VCH_FD_BATCH): that many perturbed forward solves + costs run concurrently, one worker thread and
    one library context each (vch_b200_native.run_concurrent); directions and results are those of the sequential loop."""
    if opt_config is None:
        if any(v is None for v in (b1, b2, b3, kappa)):
            raise ValueError("Either provide opt_config or all of (b1, b2, b3, kappa_sparsity).")
        opt_config = OptimizationConfig(b1=float(b1), b2=float(b2), b3=float(b3), kappa_sparsity=float(kappa))
    rng = np.random.default_rng(seed)
    Q = np.zeros_like(phi_star) if phi_Q_target is None else phi_Q_target
    T = np.zeros_like(phi_star[-1]) if phi_T_target is None else phi_T_target
    J0 = calculate_cost(phi_star, u_star, Q, T, x, y, t_hist, opt_config)
    g = r_star + opt_config.b3 * u_star
    out: List[float] = []
    print(f"Testing {num_directions} random directions in the critical cone...")
    if batch is None:
        batch = int(os.environ.get("VCH_FD_BATCH", "1"))
    hs = [_cone_direction(u_star, u_min, u_max, rng) for _ in range(num_directions)]   # drawn in the reference's order

    def one(i):
        u_eps = u_star + epsilon * hs[i]
        phi_eps, _, _ = run_main_simulation(config=fwd_config, store_history=True, control_input=u_eps, verbose=False)
        J1 = calculate_cost(phi_eps, u_eps, Q, T, x, y, t_hist, opt_config)
        return float((J1 - J0 - epsilon * float(np.sum(g * hs[i]))) / (0.5 * epsilon ** 2))
    out = _nat.run_concurrent(one, num_directions, max(1, batch))
    for i, d2 in enumerate(out):
        print(f"  Direction {i+1}/{num_directions}: estimated d²J/dh² ≈ {d2:.6e}")
    return out


def verify_sparsity_condition(u_optimal, r_optimal, kappa, tol=1e-6) -> None:
    """Prints how well  u* = 0  <=>  |r*| <= kappa  holds; "satisfied" above 99 % agreement."""
    f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    n_zero, n_small, n_match = _nat.kkt_counts(f64(u_optimal), f64(r_optimal), float(kappa), float(tol))
    total = int(np.size(u_optimal))
    print("\n" + "=" * 60 + "\nVERIFYING SPARSITY CONDITION\nCondition: u*(x,t) = 0  <=>  |r*(x,t)| <= kappa\n" + "=" * 60)
    print(f"Sparsity of final control (u* ≈ 0): {100.0 * n_zero / total:.2f}% ({n_zero}/{total} points)")
    print(f"Region where |r*| <= kappa:          {100.0 * n_small / total:.2f}% ({n_small}/{total} points)")
    print(f"Percentage of points where the conditions match: {100.0 * n_match / total:.2f}%")
    print("\n✓ The sparsity condition is satisfied." if 100.0 * n_match / total > 99.0
          else "\n⚠ The sparsity condition is not fully satisfied.")
    print("=" * 60)
