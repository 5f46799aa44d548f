"""Scratch probe: wall time of the 2D / 1D backtracking line search with sequential vs batched trials (every trial is made to
fail: cost_k = -inf, so each search evaluates max_ls_iter forward solves + costs).  python scripts/ls_batch_probe.py [N]"""
import contextlib, io, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "sparse-optimal-control-of-viscous-chan-hilliard-via-gradient-descent--1d-2d_b200")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 128
quiet = lambda f, *a, **k: (lambda b: (contextlib.redirect_stdout(b).__enter__(), f(*a, **k))[1])(io.StringIO())
def q(f, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return f(*a, **k)
sys.path.insert(0, os.path.join(PKG, "Vch_control_2D"))
import GD2_configured as G, config as C, cost2_and_function as Cst
cfg, opt = C.ForwardSolverConfig(Nx=N, Ny=N), C.OptimizationConfig()
phi, (x, y), t = q(G.run_main_simulation, config=cfg, store_history=True, control_input=None, verbose=False)
phiT, phiQ = q(G.build_targets, x, y, t, phi[0].copy(), cfg.Lx, cfg.Ly, cfg.T, False, 1, 1)
rng = np.random.default_rng(0)
u0 = np.zeros_like(phi); grad = 1e-3 * rng.standard_normal(phi.shape)
for batch in (1, 2, 4):
    for rep in range(2):
        t0 = time.perf_counter()
        r = q(G.perform_backtracking_line_search_2D, u0, -np.inf, grad, phiQ, phiT, x, y, cfg, opt, alpha_init=50.0, beta=0.8, max_ls_iter=4, batch=batch)
        dt = time.perf_counter() - t0
    print(f"2D {N}^2 x {len(t)-1} steps, 4 trials, batch={batch}: {dt:.3f} s  (cost of the last trial {r[2]:.12g})", flush=True)
for m in [k for k in list(sys.modules) if k in ("config", "GD2_configured", "cost2_and_function", "Forward2_solver", "backward2_solver", "second_order_conditions_2d")]:
    del sys.modules[m]
sys.path.remove(os.path.join(PKG, "Vch_control_2D")); sys.path.insert(0, os.path.join(PKG, "Vch_control_1D"))
import GD_1D as G1, config as C1, cost_and_function as Cst1, Forward_solver as F1
cfg1, opt1 = C1.ForwardSolverConfig(), C1.OptimizationConfig()
phi1, x1, t1 = q(F1.run_main_simulation, cfg1, True, None, False)
pT, pQ = q(G1.build_targets_1d, x1, t1, phi1[0].copy(), cfg1.Lx, cfg1.T, False, 1, 1)
g1 = 1e-3 * rng.standard_normal(phi1.shape); u1 = np.zeros_like(phi1)
for batch in (1, 5):
    for rep in range(2):
        t0 = time.perf_counter()
        r = q(G1.perform_backtracking_line_search, u1, -np.inf, g1, pQ, pT, x1, t1, opt1.b1, opt1.b2, opt1.b3, opt1.kappa_sparsity, opt1.u_min, opt1.u_max,
              cfg1, alpha_init=10.0, beta=0.8, max_ls_iter=5, batch=batch)
        dt = time.perf_counter() - t0
    print(f"1D default ({cfg1.N} intervals x {len(t1)-2} steps), 5 trials, batch={batch}: {dt:.3f} s  (cost of the last trial {r[2]:.12g})", flush=True)
