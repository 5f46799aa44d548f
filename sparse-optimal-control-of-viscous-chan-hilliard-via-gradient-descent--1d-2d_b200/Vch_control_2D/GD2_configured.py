"""Proximal-gradient driver of the 2D sparse-control problem — B200 drop-in for 2D/Vch_control_2D/GD2_configured.py.

Same loop as the reference (:295-382): adjoint -> gradient -> optimistic prox step with the previous step size ->
forward -> cost; on failure a backtracking search (alpha0 = 0.8 alpha_prev, beta = 0.8, <= 10 forwards); step growth
x1.2 (x1.5 after a 5-iteration plateau < 1e-5); stop when the relative control change < 1e-5 after 20 iterations.

`optimize()` is the loop as a function.  With `device_resident=True` (default when torch sees a GPU) u, phi_hist, phi_Q
and r stay in HBM across iterations and each optimistic iteration is ONE C-ABI call (vch2d_pgd_iteration); NumPy arrays
are materialised only for what is returned.  Running this file as a script reproduces the reference's interactive flow;
plots are produced only if the reference's optional `visualization_3d` module and matplotlib are importable.
"""
import os
import sys
import time
import warnings
from typing import Tuple

import numpy as np

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)
import vch_b200_native as _nat                                           # noqa: E402
from Forward2_solver import run_main_simulation, _time_grid               # noqa: E402
from backward2_solver import run_backward                                 # noqa: E402
from cost2_and_function import calculate_cost, calculate_gradient, proximal_step   # noqa: E402
from second_order_conditions_2d import approximate_second_order_condition_2d, verify_sparsity_condition  # noqa: E402
from config import (ForwardSolverConfig, OptimizationConfig, load_params, save_params, get_yes_no_input,   # noqa: E402
                    get_user_input_for_config)

INTERACTIVE = True
DEFAULT_TARGET_CHOICE = 1      # 1: sinusoid, 2: centred circle
DEFAULT_TRACKING_CHOICE = 1    # 1: ramp initial -> target, 2: zeros


def perform_backtracking_line_search_2D(u_k, cost_k, grad_smooth, phi_Q_target, phi_T_target, x, y, fwd_config,
                                        opt_config, alpha_init: float = 1.0, beta: float = 0.8, max_ls_iter: int = 10,
                                        batch: int = None
                                        ) -> Tuple[float, np.ndarray, float, np.ndarray, np.ndarray, float, int]:
    """Shrink alpha by beta until J(prox(u_k - alpha g)) < J(u_k) (reference :71-146).
    Returns (alpha, u_next, cost_next, phi_next, t_hist_next, seconds, attempts); the last trial if none succeeds.

    batch > 1 (argument, or VCH_LS_BATCH): `batch` trial step sizes are evaluated concurrently — one worker thread per trial,
    each on its own library context (own stream and work vectors, vch_b200_native.run_concurrent), so the independent forward
    solves overlap on the GPU; the first trial in the reference's order that lowers the cost is returned, i.e. the result
    equals the sequential search's.  Every concurrent trial holds its own trajectory: size `batch` for the host memory."""
    t0 = time.perf_counter()
    if batch is None:
        batch = int(os.environ.get("VCH_LS_BATCH", "1"))
    if batch > 1:
        alphas = [float(alpha_init)]
        for _ in range(1, max_ls_iter):
            alphas.append(alphas[-1] * beta)            # the reference's running product

        def trial(a):
            u = proximal_step(u_k, grad_smooth, a, opt_config)
            phi, _, t = run_main_simulation(config=fwd_config, store_history=True, control_input=u, verbose=False)
            return u, phi, t, calculate_cost(phi, u, phi_Q_target, phi_T_target, x, y, t, opt_config)
        done, last = 0, None
        while done < max_ls_iter:
            chunk = alphas[done:done + batch]
            res = _nat.run_concurrent(lambda j: trial(chunk[j]), len(chunk), batch)
            for j, (u, phi, t, c) in enumerate(res):
                last = (chunk[j], u, c, phi, t)
                if c < cost_k:
                    print(f"   ✓ Backtracking found a good step (α = {chunk[j]:.4f}) after {done + j + 1} attempts.")
                    return chunk[j], u, c, phi, t, time.perf_counter() - t0, done + j + 1
            done += len(chunk)
        print("[Warning] Line search could not find a step that reduces cost. Returning last try.")
        return last[0] * beta, last[1], last[2], last[3], last[4], time.perf_counter() - t0, max_ls_iter
    alpha, u_next, phi_next, t_next, cost_next = alpha_init, u_k, None, None, cost_k
    for attempt in range(1, max_ls_iter + 1):
        u_next = proximal_step(u_k, grad_smooth, alpha, opt_config)
        phi_next, _, t_next = run_main_simulation(config=fwd_config, store_history=True, control_input=u_next, verbose=False)
        cost_next = calculate_cost(phi_next, u_next, phi_Q_target, phi_T_target, x, y, t_next, opt_config)
        if cost_next < cost_k:
            print(f"   ✓ Backtracking found a good step (α = {alpha:.4f}) after {attempt} attempts.")
            return alpha, u_next, cost_next, phi_next, t_next, time.perf_counter() - t0, attempt
        alpha *= beta
    print("[Warning] Line search could not find a step that reduces cost. Returning last try.")
    return alpha, u_next, cost_next, phi_next, t_next, time.perf_counter() - t0, max_ls_iter


def _ask_choice(prompt, allowed):
    while True:
        try:
            v = int(input(prompt).strip())
            if v in allowed:
                return v
            print(f"Invalid choice. Please enter one of {allowed}.")
        except ValueError:
            print("Invalid input. Please enter a number.")


def build_targets(x, y, t_hist, phi_initial, Lx, Ly, T, interactive: bool = False, choice_t: int = 1, choice_q: int = 1
                  ) -> Tuple[np.ndarray, np.ndarray]:
    """Terminal target phi_T (1: 0.7 sin(2 pi x/Lx) cos(pi y/Ly), 2: +-1 disc of radius Lx/3.5) and tracking path phi_Q
    (1: linear ramp phi_initial -> phi_T over t/T, 2: zeros) — reference :149-228."""
    xx, yy = np.meshgrid(x, y, indexing="ij")
    if interactive:
        print("\n" + "=" * 50 + "\n🎯 CHOOSE YOUR TARGET STATE (phi_T)\n" + "=" * 50 + "\n  1: Sinusoidal Pattern\n  2: Centered Circle")
        choice_t = _ask_choice("Enter your choice for the final target (1 or 2): ", (1, 2))
    if choice_t == 1:
        print("  -> φ_T: Sinusoidal Pattern.")
        phi_T = 0.7 * np.sin(2 * np.pi * xx / Lx) * np.cos(np.pi * yy / Ly)
    else:
        print("  -> φ_T: Centered Circle.")
        phi_T = -np.ones_like(xx)
        phi_T[(xx - Lx / 2) ** 2 + (yy - Ly / 2) ** 2 < (Lx / 3.5) ** 2] = 1.0
    if interactive:
        print("\n" + "=" * 50 + "\n🚀 CHOOSE YOUR TRACKING TRAJECTORY (phi_Q)\n" + "=" * 50 +
              "\n  1: Linear path from initial state to final target\n  2: Zero target (force φ→0)")
        choice_q = _ask_choice("Enter your choice for the tracking path (1 or 2): ", (1, 2))
    if choice_q == 1:
        s = (t_hist / T)[:, np.newaxis, np.newaxis]
        phi_Q = (1 - s) * phi_initial + s * phi_T
        print("  -> φ_Q mode: time-ramp (initial → φ_T)")
    else:
        phi_Q = np.zeros((len(t_hist), len(x), len(y)))
        print("  -> φ_Q mode: zeros")
    return phi_T, phi_Q


def _trapz_l2_sq(A, x, y):
    """∫∫ A^2 dy dx with np.trapz semantics, per leading index (monitoring norms, reference :336-346)."""
    wx = np.zeros_like(x); wy = np.zeros_like(y)
    wx[:-1] += 0.5 * np.diff(x); wx[1:] += 0.5 * np.diff(x)
    wy[:-1] += 0.5 * np.diff(y); wy[1:] += 0.5 * np.diff(y)
    return np.einsum("...ij,i,j->...", A * A, wx, wy)


def optimize(fwd_config: ForwardSolverConfig, opt_config: OptimizationConfig, choice_t: int = DEFAULT_TARGET_CHOICE,
             choice_q: int = DEFAULT_TRACKING_CHOICE, max_iter=None, device_resident=None, verbose: bool = True):
    """Run the PGD loop.  Returns a dict with u, phi_hist, r (last adjoint), x, y, t_hist, targets, cost_history,
    alpha_history, tracking/terminal error histories and timers."""
    if device_resident is None:
        try:
            import torch
            device_resident = torch.cuda.is_available()
        except ImportError:
            device_resident = False
    b1, b2, b3, ksp = opt_config.b1, opt_config.b2, opt_config.b3, opt_config.kappa_sparsity
    n_iter = int(opt_config.max_iter if max_iter is None else max_iter)
    phi_k, (x, y), t_hist = run_main_simulation(config=fwd_config, store_history=True, control_input=None, verbose=False)
    phi_init = phi_k[0].copy()
    phi_T, phi_Q = build_targets(x, y, t_hist, phi_init, fwd_config.Lx, fwd_config.Ly, fwd_config.T, False, choice_t, choice_q)
    u_k = np.zeros_like(phi_k)
    cost_k = calculate_cost(phi_k, u_k, phi_Q, phi_T, x, y, t_hist, opt_config)
    hist = dict(cost=[cost_k], alpha=[], track=[], term=[], t_backward=0.0, t_gradprox=0.0, t_forward=0.0, t_cost=0.0,
                t_iteration=0.0, t_linesearch=0.0, ls_calls=0, ls_attempts=0, good_alphas=[])
    alpha_prev, plateau = opt_config.alpha_max, 0
    dts, _ = _time_grid(float(fwd_config.T), float(fwd_config.dt_initial))
    Nx, Ny = int(fwd_config.Nx), int(fwd_config.Ny)
    ctx = _nat.ctx2d(Nx, Ny, fwd_config.Lx / Nx, fwd_config.Ly / Ny, fwd_config.Lx, fwd_config.Ly, fwd_config.tau,
                     fwd_config.gamma, fwd_config.c1, fwd_config.c2, fwd_config.kappa, 1e-2)
    if device_resident:
        import torch
        dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
        d_u, d_phi, d_Q, d_T = dev(u_k), dev(phi_k), dev(phi_Q), dev(phi_T)
        d_un, d_phin, d_r = torch.empty_like(d_u), torch.empty_like(d_u), torch.empty_like(d_u)
    r_k = None
    area = float((x[-1] - x[0]) * (y[-1] - y[0])); span = float(t_hist[-1] - t_hist[0])
    rms = float(np.sqrt(max(area, 1e-30) * max(span, 1e-30)))
    wt = np.zeros_like(t_hist); wt[:-1] += 0.5 * np.diff(t_hist); wt[1:] += 0.5 * np.diff(t_hist)
    denQ = float(np.sqrt(max(np.dot(wt, _trapz_l2_sq(phi_Q, x, y)), 0.0)))
    denQ = rms if denQ < 1e-9 * rms else denQ
    denT = float(np.sqrt(max(_trapz_l2_sq(phi_T, x, y), 0.0))) + 1e-12
    raw_track = raw_term = None          # raw integrals of the accepted iterate (device-resident path)
    for k in range(n_iter):
        it0 = time.perf_counter()
        if verbose:
            print(f"\n📍 Iteration {k+1}/{n_iter} | Current Cost = {cost_k:.6f}")
        if device_resident:
            t0 = time.perf_counter()
            _, _, J, red, _ = ctx.pgd_iteration(d_u, d_phi, d_Q, d_T, t_hist, dts, x, y, b1, b2, b3, ksp, opt_config.u_min,
                                                opt_config.u_max, alpha_prev, u_out=d_un, phi_out=d_phin, r_out=d_r)
            hist["t_iteration"] += time.perf_counter() - t0
            cost_try, change_sq, unorm_sq = float(J[0]), float(red[0]), float(red[1])
        else:
            t0 = time.perf_counter()
            _, _, r_k = run_backward(phi_k, x, y, t_hist, fwd_config, b1, b2, phi_Q, phi_T)
            hist["t_backward"] += time.perf_counter() - t0
            t0 = time.perf_counter()
            grad = calculate_gradient(r_k, u_k, opt_config)
            u_try = proximal_step(u_k, grad, alpha_prev, opt_config)
            hist["t_gradprox"] += time.perf_counter() - t0
            t0 = time.perf_counter()
            phi_try, _, t_try = run_main_simulation(config=fwd_config, store_history=True, control_input=u_try, verbose=False)
            hist["t_forward"] += time.perf_counter() - t0
            t0 = time.perf_counter()
            cost_try = calculate_cost(phi_try, u_try, phi_Q, phi_T, x, y, t_try, opt_config)
            hist["t_cost"] += time.perf_counter() - t0
        if cost_try < cost_k:
            if verbose:
                print(f"   ✓ Optimistic step successful (α = {alpha_prev:.4f})")
            alpha_k, cost_next = alpha_prev, cost_try
            hist["good_alphas"].append(alpha_k)
            if device_resident:
                d_u, d_un = d_un, d_u
                d_phi, d_phin = d_phin, d_phi
                raw_track, raw_term = float(J[5]), float(J[6])
            else:
                change_sq, unorm_sq = float(np.sum((u_try - u_k) ** 2)), float(np.sum(u_k ** 2))
                u_k, phi_k = u_try, phi_try
        elif device_resident:
            # backtracking on device-resident state (reference :71-146): prox kernel -> forward -> cost, nothing leaves HBM
            if verbose:
                print("   ⚠ Optimistic step failed. Backtracking...")
            ls0 = time.perf_counter()
            alpha_k, cost_next = alpha_prev * 0.8, cost_k
            hist["ls_calls"] += 1
            for attempt in range(1, 11):
                hist["ls_attempts"] += 1
                d_un, _, red = _nat.grad_prox(d_u, d_r, b3, alpha_k, ksp, opt_config.u_min, opt_config.u_max)
                d_phin, _, _ = ctx.forward(d_phi[0].contiguous(), d_un, dts)
                J = ctx.cost(d_phin, d_un, d_Q, d_T, x, y, t_hist, b1, b2, b3, ksp)
                cost_next = float(J[0])
                if cost_next < cost_k:
                    if verbose:
                        print(f"   ✓ Backtracking found a good step (α = {alpha_k:.4f}) after {attempt} attempts.")
                    break
                if attempt < 10:
                    alpha_k *= 0.8
            else:
                alpha_k *= 0.8           # the reference returns the shrunk alpha with the last trial
                print("[Warning] Line search could not find a step that reduces cost. Returning last try.")
            hist["t_linesearch"] += time.perf_counter() - ls0
            change_sq, unorm_sq = float(red[0]), float(red[1])
            d_u, d_phi = d_un, d_phin
            d_un, d_phin = torch.empty_like(d_u), torch.empty_like(d_u)
            raw_track, raw_term = float(J[5]), float(J[6])
        else:
            if verbose:
                print("   ⚠ Optimistic step failed. Backtracking...")
            grad = calculate_gradient(r_k, u_k, opt_config)
            alpha_k, u_next, cost_next, phi_next, _, secs, attempts = perform_backtracking_line_search_2D(
                u_k, cost_k, grad, phi_Q, phi_T, x, y, fwd_config, opt_config, alpha_init=alpha_prev * 0.8)
            hist["t_linesearch"] += secs; hist["ls_calls"] += 1; hist["ls_attempts"] += attempts
            change_sq, unorm_sq = float(np.sum((u_next - u_k) ** 2)), float(np.sum(u_k ** 2))
            u_k, phi_k = u_next, phi_next
        hist["cost"].append(cost_next); hist["alpha"].append(alpha_k)
        if device_resident:              # monitoring norms = square roots of the cost kernel's raw integrals (no extra pass)
            hist["track"].append(float(np.sqrt(max(raw_track, 0.0))) / (denQ + 1e-12))
            hist["term"].append(float(np.sqrt(max(raw_term, 0.0))) / denT)
        else:
            hist["track"].append(float(np.sqrt(max(np.dot(wt, _trapz_l2_sq(phi_k - phi_Q, x, y)), 0.0))) / (denQ + 1e-12))
            hist["term"].append(float(np.sqrt(max(_trapz_l2_sq(phi_k[-1] - phi_T, x, y), 0.0))) / denT)
        plateau = plateau + 1 if (k > 0 and abs(hist["cost"][-1] - hist["cost"][-2]) < 1e-5) else 0
        if plateau >= 5:
            if verbose:
                print(f"   [Notice] Cost has plateaued for {plateau} iterations. Boosting step size.")
            alpha_prev, plateau = min(opt_config.alpha_max, alpha_k * 1.5), 0
        else:
            alpha_prev = min(opt_config.alpha_max, alpha_k * 1.2)
        change = np.sqrt(change_sq) / (np.sqrt(unorm_sq) + 1e-9)
        if verbose:
            print(f"   Relative control change: {change:.6e}\n   Iteration time: {time.perf_counter()-it0:.2f}s")
        cost_k = cost_next
        if change < 1e-5 and k > 20:
            if verbose:
                print(f"\n🎉 Convergence reached at iteration {k+1}!")
            break
    if device_resident:
        u_k, phi_k = d_u.cpu().numpy(), d_phi.cpu().numpy()
    _, _, r_k = run_backward(phi_k, x, y, t_hist, fwd_config, b1, b2, phi_Q, phi_T)
    return dict(u=u_k, phi_hist=phi_k, r=r_k, x=x, y=y, t_hist=t_hist, phi_T=phi_T, phi_Q=phi_Q, phi_initial=phi_init,
                cost_history=hist["cost"], alpha_history=hist["alpha"], tracking_error_history=hist["track"],
                terminal_error_history=hist["term"], timers=hist)


if __name__ == "__main__":
    warnings.filterwarnings("ignore")
    print("=" * 60 + "\n    2D GRADIENT DESCENT OPTIMIZATION \n" + "=" * 60)
    params = load_params()
    if INTERACTIVE and get_yes_no_input("Do you want to modify the simulation parameters?"):
        fwd_cfg = get_user_input_for_config(ForwardSolverConfig, "Forward Solver Parameters", params.forward_solver)
        opt_cfg = get_user_input_for_config(OptimizationConfig, "Optimization Parameters", params.optimization)
    else:
        fwd_cfg, opt_cfg = params.forward_solver, params.optimization
    print("\n--- Using the following parameters ---\nForward Solver Config:\n" + fwd_cfg.model_dump_json(indent=2))
    print("Optimization Config:\n" + opt_cfg.model_dump_json(indent=2))
    ct, cq = DEFAULT_TARGET_CHOICE, DEFAULT_TRACKING_CHOICE
    if INTERACTIVE:
        if not get_yes_no_input("Proceed to optimization with these parameters?"):
            print("\n🛑 Optimization cancelled by user.")
            sys.exit(0)
        ct = _ask_choice("Final target — 1: Sinusoidal Pattern, 2: Centered Circle: ", (1, 2))
        cq = _ask_choice("Tracking path — 1: linear ramp, 2: zeros: ", (1, 2))
    start = time.time()
    res = optimize(fwd_cfg, opt_cfg, ct, cq)
    tm = res["timers"]
    print("\n" + "=" * 50 + "\n⏱️  TIME STUDY SUMMARY\n" + "=" * 50)
    print(f"Total fused-iteration time (device): {tm['t_iteration']:.3f}s")
    print(f"Total backward-solve time:        {tm['t_backward']:.3f}s")
    print(f"Total optimistic forward time:    {tm['t_forward']:.3f}s")
    print(f"Total optimistic cost time:       {tm['t_cost']:.3f}s")
    print(f"Total backtracking time:          {tm['t_linesearch']:.3f}s  (calls={tm['ls_calls']}, attempts={tm['ls_attempts']})")
    print(f"Completed Iterations: {len(res['cost_history'])-1} | Final Cost: {res['cost_history'][-1]:.5f} | "
          f"Cost Reduction: {100*(1-res['cost_history'][-1]/res['cost_history'][0]):.2f}% | wall {time.time()-start:.1f}s")
    if tm["good_alphas"]:
        print(f"\n💡 ALPHA ADVISOR: mean accepted optimistic step size = {np.mean(tm['good_alphas']):.4f}")
    try:
        print("\n--- Checking Second-Order Sufficient Condition (Coercivity) ---")
        vals = approximate_second_order_condition_2d(
            u_star=res["u"], r_star=res["r"], phi_star=res["phi_hist"], x=res["x"], y=res["y"], t_hist=res["t_hist"],
            b1=opt_cfg.b1, b2=opt_cfg.b2, b3=opt_cfg.b3, kappa=opt_cfg.kappa_sparsity, phi_Q_target=res["phi_Q"],
            phi_T_target=res["phi_T"], u_min=opt_cfg.u_min, u_max=opt_cfg.u_max, num_directions=5, epsilon=1e-4, seed=42,
            fwd_config=fwd_cfg)
        print("\n✓ Coercivity condition appears to hold in tested directions." if all(v > 0 for v in vals)
              else "\n⚠ Coercivity condition may fail; non-positive second derivatives found.")
        verify_sparsity_condition(res["u"], res["r"], opt_cfg.kappa_sparsity)
    except Exception as exc:
        print(f"\n[Warning] Could not perform final analysis (second-order/sparsity): {exc}")
    try:                                                   # plots only when the optional plotting module is around
        from visualization_3d import plot_convergence_history
        plot_convergence_history(res["cost_history"], res["terminal_error_history"], res["tracking_error_history"])
    except Exception:
        pass
    np.save("optimal_control_2d.npy", res["u"])
    save_params(fwd_cfg, opt_cfg, len(res["cost_history"]) - 1)
    print("\n" + "=" * 50 + "\n✅ OPTIMIZATION COMPLETE\n" + "=" * 50)
