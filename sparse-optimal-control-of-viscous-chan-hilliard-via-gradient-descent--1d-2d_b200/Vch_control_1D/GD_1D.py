"""Proximal-gradient driver of the 1D sparse-control problem — B200 drop-in for 1D/Vch_control_1D/GD_1D.py.

Loop semantics follow the reference (:353-480): optimistic step with alpha_prev, fallback backtracking that restarts
from alpha_prev itself (beta = 0.8, <= 5 forwards), growth x1.2 (x2.0 after a 10-iteration plateau < 1e-7), stop when
the relative control change < 1e-5 after 10 iterations.  `optimize()` is that loop as a function; `optimize_ensemble()`
runs B independent problems at once (BASELINE config 4): every device call carries a leading batch axis and the batch
can be split across GPUs with no communication (see `shard_range`).
"""
import os
import sys
import time

import numpy as np

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)
import vch_b200_native as _nat                                                        # noqa: E402
from Forward_solver import run_main_simulation, run_main_simulation_batch, init_phi_random, _time_grid, delta_sep  # noqa: E402
from backward_solver import run_backward                                              # noqa: E402
from cost_and_function import calculate_cost, calculate_cost_batch, calculate_gradient, perform_gradient_step   # noqa: E402
from second_order_conditions import approximate_second_order_condition               # noqa: E402
from config import (ForwardSolverConfig, OptimizationConfig, get_user_input_for_config, get_yes_no_input,   # noqa: E402
                    save_params, load_params)

INTERACTIVE = True
DEFAULT_TARGET_CHOICE = 1      # 1: sine, 2: cosine, 3: normalised tangent
DEFAULT_TRACKING_CHOICE = 1    # 1: ramp initial -> target, 2: zeros

_f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)


def perform_proximal_and_projection(u_temp, alpha, kappa, u_min, u_max):
    """clip(sign(v) max(|v| - alpha kappa, 0), u_min, u_max) for v = u_temp (reference :56-71): the prox kernel with a
    zero gradient, so v passes through unchanged."""
    u_temp = _f64(u_temp)
    un, _, _ = _nat.grad_prox(u_temp, np.zeros_like(u_temp), 0.0, float(alpha), float(kappa), float(u_min), float(u_max))
    return un


def perform_backtracking_line_search(u_k, cost_k, grad_smooth, phi_Q_target, phi_T_target, x, t_hist, b1, b2, b3, kappa,
                                     u_min, u_max, fwd_config, alpha_init=10.0, beta=0.8, max_ls_iter=5, batch=None):
    """Returns (alpha, u_next, cost_next, phi_next, rejected_seconds, accepted_seconds, trials) — reference :73-113.

    batch > 1 (argument, or VCH_LS_BATCH): the trial step sizes alpha_init * beta^j are evaluated `batch` at a time as ONE
    multi-problem launch (forward solves and costs carry a batch axis; one CTA per trial), and the first trial in the
    reference's order that lowers the cost is returned — the same (alpha, u, cost, phi, trials) as the sequential search."""
    if batch is None:
        batch = int(os.environ.get("VCH_LS_BATCH", "1"))
    if batch > 1:
        alphas = [float(alpha_init)]
        for _ in range(1, max_ls_iter):
            alphas.append(alphas[-1] * beta)            # the reference's running product, not a power
        rejected, done = 0.0, 0
        u_next = phi_next = cost_next = None
        while done < max_ls_iter:
            chunk = alphas[done:done + batch]
            t0 = time.perf_counter()
            us = np.stack([perform_proximal_and_projection(perform_gradient_step(u_k, grad_smooth, a), a, kappa, u_min, u_max)
                           for a in chunk])
            phis, _, _ = run_main_simulation_batch(fwd_config, us)
            J = calculate_cost_batch(phis, us, phi_Q_target, phi_T_target, x, t_hist, b1, b2, b3, kappa)
            dt = (time.perf_counter() - t0) / len(chunk)
            for j, a in enumerate(chunk):
                u_next, phi_next, cost_next = us[j], phis[j], float(J[j, 0])
                if cost_next < cost_k:
                    return a, u_next, cost_next, phi_next, rejected, dt, done + j + 1
                rejected += dt
            done += len(chunk)
        print("[Warning] Line search could not find a step that reduces cost.")
        return alphas[-1] * beta, u_next, cost_next, phi_next, rejected, 0.0, max_ls_iter
    alpha, rejected, trials = alpha_init, 0.0, 0
    u_next = phi_next = cost_next = None
    for _ in range(max_ls_iter):
        trials += 1
        u_next = perform_proximal_and_projection(perform_gradient_step(u_k, grad_smooth, alpha), alpha, kappa, u_min, u_max)
        t0 = time.perf_counter()
        phi_next, _, _ = run_main_simulation(fwd_config, store_history=True, control_input=u_next, verbose=False)
        cost_next = calculate_cost(phi_next, u_next, phi_Q_target, phi_T_target, x, t_hist, b1, b2, b3, kappa, verbose=False)
        dt = time.perf_counter() - t0
        if cost_next < cost_k:
            return alpha, u_next, cost_next, phi_next, rejected, dt, trials
        alpha *= beta
        rejected += dt
    print("[Warning] Line search could not find a step that reduces cost.")
    return alpha, u_next, cost_next, phi_next, rejected, 0.0, trials


def verify_sparsity_condition(u_optimal, r_optimal, kappa, tol=1e-6):
    """u* = 0  <=>  |r*| <= kappa, counted by one fused device kernel (reference :115-147)."""
    n_zero, n_small, n_match = _nat.kkt_counts(_f64(u_optimal), _f64(r_optimal), float(kappa), float(tol))
    total = int(np.size(u_optimal))
    print("\n" + "=" * 60 + "\nVERIFYING SPARSITY CONDITION (Theorem 4.7)\nCondition: u*(x,t) = 0  <=>  |r*(x,t)| <= kappa\n" + "=" * 60)
    print(f"Sparsity of final control (u* ≈ 0): {100.0 * n_zero / total:.2f}% ({n_zero}/{total} points)")
    print(f"Region where |r*| <= kappa:          {100.0 * n_small / total:.2f}% ({n_small}/{total} points)")
    print(f"Percentage of points where the conditions match: {100.0 * n_match / total:.2f}%")
    print("\n✓ The sparsity condition is satisfied." if 100.0 * n_match / total > 99.0
          else "\n⚠ The sparsity condition is not fully satisfied.")
    print("=" * 60)


def _ask_int(prompt, allowed):
    while True:
        try:
            v = int(input(prompt).strip())
            if v in allowed:
                return v
            print(f"Invalid choice. Please enter one of {list(allowed)}.")
        except ValueError:
            print("Invalid input. Please enter a number.")


def target_profile(x, Lx, choice_t=1, A_T=0.7, k_tan=0.45):
    """phi_T for the three reference shapes (:224-240)."""
    if choice_t == 1:
        return A_T * np.sin(2.0 * np.pi * x / Lx)
    if choice_t == 2:
        return A_T * np.cos(2.0 * np.pi * x / Lx)
    raw = np.tan(2.0 * np.pi * k_tan * (x / Lx - 0.5))
    peak = np.max(np.abs(raw))
    return A_T * raw / (peak if peak > 1e-12 else 1.0)


def build_targets_1d(x, t_hist, phi_initial, Lx, T, interactive=False, choice_t=DEFAULT_TARGET_CHOICE,
                     choice_q=DEFAULT_TRACKING_CHOICE):
    """(phi_T_target, phi_Q_target) — reference :151-254.  phi_Q ramps over t_hist / t_hist[-1] (choice_q = 1) or is 0."""
    A_T, k_tan = 0.7, 0.45
    if interactive:
        print("\n" + "=" * 50 + "\n🎯 CHOOSE YOUR TARGET STATE (phi_T)\n" + "=" * 50)
        print("  1: Sinusoidal  (A_T * sin(2πx/Lx))\n  2: Cosine      (A_T * cos(2πx/Lx))\n  3: Tan (safe)  (normalized, no poles; max|φ_T| = A_T)")
        choice_t = _ask_int("Enter your choice for the final target (1/2/3): ", (1, 2, 3))
        try:
            txt = input("Enter amplitude A_T for φ_T (default 0.7): ").strip()
            A_T = float(txt) if txt else 0.7
        except Exception:
            print("[warn] Invalid amplitude; using A_T = 0.7")
        if choice_t == 3:
            try:
                txt = input("Enter k_tan ∈ (0, 0.5) for tan (default 0.45): ").strip()
                k_tan = max(1e-3, min(0.49, float(txt))) if txt else 0.45
            except Exception:
                print("[warn] Invalid k_tan; using 0.45")
    phi_T = target_profile(x, Lx, choice_t, A_T, k_tan)
    print(("  -> φ_T: Sinusoidal", "  -> φ_T: Cosine", f"  -> φ_T: Tan (safe), k_tan={k_tan:g}, normalized to amplitude A_T")[choice_t - 1])
    if interactive:
        print("\n" + "=" * 50 + "\n🚀 CHOOSE YOUR TRACKING TRAJECTORY (phi_Q)\n" + "=" * 50)
        print("  1: Linear path from initial state to φ_T\n  2: Zero path (φ_Q ≡ 0)")
        choice_q = _ask_int("Enter your choice for the tracking path (1/2): ", (1, 2))
    if choice_q == 1:
        s = (t_hist / (t_hist[-1] if t_hist[-1] > 0 else 1.0))[:, np.newaxis]
        phi_Q = (1.0 - s) * phi_initial + s * phi_T
        print("  -> φ_Q mode: time-ramp (initial → φ_T)")
    else:
        phi_Q = np.zeros((len(t_hist), len(x)))
        print("  -> φ_Q mode: zeros")
    return phi_T, phi_Q


def optimize(fwd_config, opt_config, choice_t=DEFAULT_TARGET_CHOICE, choice_q=DEFAULT_TRACKING_CHOICE, max_iter=None,
             verbose=True):
    """The reference's PGD loop (:333-480) as a function.  Returns a dict of results and histories."""
    o = opt_config
    phi_k, x, t_hist = run_main_simulation(fwd_config, store_history=True, verbose=False)
    phi_T, phi_Q = build_targets_1d(x, t_hist, phi_k[0].copy(), float(fwd_config.Lx), float(fwd_config.T), False, choice_t, choice_q)
    u_k = np.zeros_like(phi_k)
    cost_k = calculate_cost(phi_k, u_k, phi_Q, phi_T, x, t_hist, o.b1, o.b2, o.b3, o.kappa_sparsity)
    costs, alphas, trials = [cost_k], [], []
    alpha_prev, plateau, r_k = o.alpha_max, 0, np.zeros_like(u_k)
    timers = dict(backward=0.0, line_search=0.0, accepted=0.0)
    n_iter = int(o.max_iter if max_iter is None else max_iter)
    done = n_iter
    for k in range(n_iter):
        if verbose:
            print(f"\n--- Iteration {k+1}/{n_iter} ---\nCurrent Cost = {cost_k:.6f}")
        t0 = time.perf_counter()
        _, _, r_k = run_backward(phi_k, x, t_hist, o.b1, o.b2, phi_Q, phi_T)
        timers["backward"] += time.perf_counter() - t0
        grad = calculate_gradient(r_k, u_k, o.b3)
        u_try = perform_proximal_and_projection(perform_gradient_step(u_k, grad, alpha_prev), alpha_prev, o.kappa_sparsity, o.u_min, o.u_max)
        t0 = time.perf_counter()
        phi_try, _, _ = run_main_simulation(fwd_config, store_history=True, control_input=u_try, verbose=False)
        cost_try = calculate_cost(phi_try, u_try, phi_Q, phi_T, x, t_hist, o.b1, o.b2, o.b3, o.kappa_sparsity, verbose=False)
        dt_try = time.perf_counter() - t0
        if cost_try < cost_k:
            if verbose:
                print(f"Optimistic step successful with alpha = {alpha_prev:.4f}")
            alpha_k, u_next, cost_next, phi_next = alpha_prev, u_try, cost_try, phi_try
            timers["accepted"] += dt_try
            trials.append(1)
        else:
            if verbose:
                print("[Notice] Optimistic step failed. Engaging full backtracking search...")
            alpha_k, u_next, cost_next, phi_next, rej, acc, ntr = perform_backtracking_line_search(
                u_k, cost_k, grad, phi_Q, phi_T, x, t_hist, o.b1, o.b2, o.b3, o.kappa_sparsity, o.u_min, o.u_max, fwd_config,
                alpha_init=alpha_prev)
            timers["line_search"] += rej; timers["accepted"] += acc
            trials.append(ntr)
        costs.append(cost_next); alphas.append(alpha_k)
        plateau = plateau + 1 if (k > 0 and abs(costs[-1] - costs[-2]) < 1e-7) else 0
        if plateau >= 10:
            alpha_prev, plateau = min(o.alpha_max, alpha_k * 2.0), 0
        else:
            alpha_prev = min(o.alpha_max, alpha_k * 1.2)
        change = np.linalg.norm(u_next - u_k) / (np.linalg.norm(u_k) + 1e-9)
        if verbose:
            print(f"Relative control change: {change:.6e}")
        u_k = u_next.copy()
        if change < 1e-5 and k > 10:
            if verbose:
                print(f"\nConvergence reached at iteration {k+1}.")
            done = k + 1
            break
        cost_k, phi_k = cost_next, phi_next
    return dict(u=u_k, r=r_k, phi_hist=phi_k, x=x, t_hist=t_hist, phi_T=phi_T, phi_Q=phi_Q, cost_history=costs,
                alpha_history=alphas, ls_trials=trials, timers=timers, iterations=done)


def shard_range(n_items, rank, world):
    """Contiguous block partition of an ensemble over ranks (remainder to the first ranks); no communication needed."""
    base, extra = divmod(int(n_items), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def optimistic_iteration_ensemble(ctx_fwd, ctx_adj, u, phi_hist, phi_Q, phi_T, x, t_hist, dts, phi_init, b1, b2, b3, ksp,
                                  alpha, u_min=-1.0, u_max=1.0):
    """One optimistic PGD iteration for B problems at once (arrays carry a leading batch axis; weights are length-B
    vectors).  Four launches: adjoint, prox, forward, cost.  Returns (u_new, phi_hist_new, J (B,5), red (B,4), r)."""
    _, _, r = ctx_adj.adjoint(phi_hist, t_hist, b1, b2, phi_Q, phi_T)
    u_new, red = ctx_fwd.grad_prox(u, r, b3, alpha, ksp, u_min, u_max)
    hist_new, _, _ = ctx_fwd.forward(phi_init, u_new, dts)
    J = ctx_fwd.cost(hist_new, u_new, phi_Q, phi_T, x, t_hist, b1, b2, b3, ksp)
    return u_new, hist_new, J, red, r


def make_ensemble(B, fwd_config=None, seed=1234):
    """BASELINE config 4: B problems on one grid, varied targets and weights: choice_t ~ U{1,2,3},
    A_T ~ U[0.3,0.8], kappa_sp ~ logU[1e-5,1e-3], b1 ~ U[0.1,1], b2 ~ U[5,20], b3 ~ logU[1e-4,1e-2]."""
    cfg = ForwardSolverConfig() if fwd_config is None else fwd_config
    rng = np.random.default_rng(seed)
    x = np.linspace(0, cfg.Lx, cfg.N + 1)
    dts, t_hist = _time_grid(float(cfg.T), float(cfg.dt_initial))
    phi_init = np.tile(init_phi_random(cfg.N, delta_sep, amp=0.01, seed=42), (B, 1))
    ct = rng.integers(1, 4, B); A = rng.uniform(0.3, 0.8, B)
    ksp = np.exp(rng.uniform(np.log(1e-5), np.log(1e-3), B)); b1 = rng.uniform(0.1, 1.0, B); b2 = rng.uniform(5, 20, B)
    b3 = np.exp(rng.uniform(np.log(1e-4), np.log(1e-2), B))
    phi_T = np.stack([target_profile(x, cfg.Lx, int(ct[b]), float(A[b])) for b in range(B)])
    s = (t_hist / t_hist[-1])[None, :, None]
    phi_Q = (1.0 - s) * phi_init[:, None, :] + s * phi_T[:, None, :]
    return dict(x=x, dts=dts, t_hist=t_hist, phi_init=phi_init, phi_T=phi_T, phi_Q=phi_Q, b1=b1, b2=b2, b3=b3, ksp=ksp,
                choice_t=ct, A_T=A)


if __name__ == "__main__":
    print("Welcome to the Cahn-Hilliard Optimal Control Simulator.")
    last = load_params()
    fwd_cfg = get_user_input_for_config(ForwardSolverConfig, "STEP 1: Configure the Forward Solver", previous_instance=last.forward_solver)
    print("\nRunning a baseline simulation with these parameters...")
    phi_hist0, x0, _ = run_main_simulation(fwd_cfg, store_history=True, verbose=False)
    print(f"Baseline simulation complete: ||phi(T)||_inf = {np.max(np.abs(phi_hist0[-1])):.4f}")
    if not get_yes_no_input("Do you want to proceed to optimization with these parameters?"):
        print("Exiting simulator.")
        sys.exit(0)
    opt_cfg = get_user_input_for_config(OptimizationConfig, "STEP 2: Configure the Optimization Algorithm", previous_instance=last.optimization)
    ct = _ask_int("Final target (1: sine, 2: cosine, 3: tan): ", (1, 2, 3)) if INTERACTIVE else DEFAULT_TARGET_CHOICE
    cq = _ask_int("Tracking path (1: ramp, 2: zeros): ", (1, 2)) if INTERACTIVE else DEFAULT_TRACKING_CHOICE
    t_start = time.perf_counter()
    res = optimize(fwd_cfg, opt_cfg, ct, cq)
    print("\nOptimization finished.")
    np.save("optimal_control.npy", res["u"])
    print("Optimal control saved as 'optimal_control.npy'")
    vals = approximate_second_order_condition(fwd_config=fwd_cfg, u_star=res["u"], r_star=res["r"], phi_star=res["phi_hist"],
                                              x=res["x"], t_hist=res["t_hist"], b1=opt_cfg.b1, b2=opt_cfg.b2, b3=opt_cfg.b3,
                                              kappa=opt_cfg.kappa_sparsity, phi_Q_target=res["phi_Q"], phi_T_target=res["phi_T"],
                                              u_min=opt_cfg.u_min, u_max=opt_cfg.u_max, num_directions=3, epsilon=1e-4, seed=42)
    for i, d2 in enumerate(vals, start=1):
        print(f"  Direction {i}: estimated second derivative = {d2:.6e}")
    verify_sparsity_condition(res["u"], res["r"], opt_cfg.kappa_sparsity)
    tm = res["timers"]
    print("\n" + "=" * 60 + "\nCOMPUTATIONAL TIME STUDY (wall-clock)\n" + "=" * 60)
    print(f"Total optimization (loop) time       : {time.perf_counter() - t_start:.3f} s")
    print(f"  ├─ Backward solves (adjoint r)     : {tm['backward']:.3f} s")
    print(f"  ├─ Line-search backtracking (rejs.): {tm['line_search']:.3f} s")
    print(f"  └─ Successful step eval (accepted) : {tm['accepted']:.3f} s")
    save_params(fwd_cfg, opt_cfg, res["iterations"])
