"""Finite-difference second-order check of the 1D problem — drop-in for 1D/Vch_control_1D/second_order_conditions.py.
The perturbed forward solves and costs of all directions run as one multi-problem launch (batch axis of the 1D kernels);
the direction sampler follows the reference's critical-cone rules incl. the L1 kink (:33-55)."""
from __future__ import annotations

import os
import sys
from typing import List

import numpy as np

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)
from config import ForwardSolverConfig          # noqa: E402
from Forward_solver import run_main_simulation, run_main_simulation_batch   # noqa: E402
from cost_and_function import calculate_cost, calculate_cost_batch     # noqa: E402


def _generate_direction(u_star, r_star, u_min, u_max, kappa, b3, rng, tol=1e-8, tol_s=1e-9):
    """Unit direction in the critical cone: inward at active bounds; at u* = 0 frozen where |r*+b3 u*| < kappa and
    one-signed where the subgradient is at +-kappa."""
    v = rng.standard_normal(size=u_star.shape)
    s = r_star + b3 * u_star
    zero = np.abs(u_star) <= tol
    rules = ((u_star <= u_min + tol, +1), (u_star >= u_max - tol, -1), (zero & (np.abs(s) < kappa - tol_s), 0),
             (zero & (s >= kappa - tol_s), -1), (zero & (s <= -kappa + tol_s), +1))
    for mask, sign in rules:
        if np.any(mask):
            v[mask] = sign * np.abs(v[mask])
    nv = np.linalg.norm(v)
    if nv == 0:
        v[np.unravel_index(np.argmax(np.abs(s)), s.shape)] = nv = 1.0
    return v / nv


def _coerce_rng(seed_or_rng=None):
    if isinstance(seed_or_rng, np.random.Generator):
        return seed_or_rng
    try:
        return np.random.default_rng(None if seed_or_rng is None else int(seed_or_rng))
    except Exception:
        return np.random.default_rng()


def approximate_second_order_condition(fwd_config: ForwardSolverConfig, u_star, r_star, phi_star, x, t_hist, b1, b2, b3,
                                       kappa, phi_Q_target, phi_T_target, u_min, u_max, num_directions: int = 10,
                                       epsilon: float = 1e-4, seed: int | None = None, rng=None, batch=None) -> List[float]:
    """batch (argument or VCH_FD_BATCH; default: all directions): perturbed forward solves + costs evaluated that many at a
    time as one multi-problem launch; batch = 1 is the reference's loop.  The directions (and so the results) are the same."""
    rng = _coerce_rng(rng if rng is not None else seed)
    J0 = calculate_cost(phi_star, u_star, phi_Q_target, phi_T_target, x, t_hist, b1, b2, b3, kappa, verbose=False)
    g = r_star + b3 * u_star
    out: List[float] = []
    if batch is None:
        batch = int(os.environ.get("VCH_FD_BATCH", str(max(1, num_directions))))
    if batch > 1:
        hs = [_generate_direction(u_star, r_star, u_min, u_max, kappa, b3, rng) for _ in range(num_directions)]
        for lo in range(0, num_directions, batch):
            us = np.stack([u_star + epsilon * h for h in hs[lo:lo + batch]])
            phis, _, _ = run_main_simulation_batch(fwd_config, us)
            J = calculate_cost_batch(phis, us, phi_Q_target, phi_T_target, x, t_hist, b1, b2, b3, kappa)
            for j, h in enumerate(hs[lo:lo + batch]):
                out.append(float((float(J[j, 0]) - J0 - epsilon * np.sum(g * h)) / (0.5 * epsilon ** 2)))
        return out
    for _ in range(num_directions):
        h = _generate_direction(u_star, r_star, u_min, u_max, kappa, b3, rng)
        u_eps = u_star + epsilon * h
        phi_eps, _, _ = run_main_simulation(fwd_config=fwd_config, store_history=True, control_input=u_eps, verbose=False)
        J1 = calculate_cost(phi_eps, u_eps, phi_Q_target, phi_T_target, x, t_hist, b1, b2, b3, kappa, verbose=False)
        out.append(float((J1 - J0 - epsilon * np.sum(g * h)) / (0.5 * epsilon ** 2)))
    return out
