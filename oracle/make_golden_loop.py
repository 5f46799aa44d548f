#!/usr/bin/env python
"""Golden for the PGD DRIVER LOOP: runs the unmodified reference's 2D modules through the same statements as the loop body of
src/2D/Vch_control_2D/GD2_configured.py:291-382 (non-interactive) for K iterations on the default 128^2 config and records
cost / step-size / control-change histories.  Slow (~4.5 min per iteration on this container).  Test infrastructure only.

usage: python oracle/make_golden_loop.py [K=8]
"""
import os, sys, io, contextlib, time
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "_mpl_shim"))
sys.path.insert(0, "/root/reference/src/2D/Vch_control_2D")
os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache_golden")
os.chdir("/tmp")
K = int(sys.argv[1]) if len(sys.argv) > 1 else 8
import Forward2_solver as F, backward2_solver as B, cost2_and_function as C, GD2_configured as G
from config import ForwardSolverConfig, OptimizationConfig

def q(f, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return f(*a, **k)

fwd, opt = ForwardSolverConfig(), OptimizationConfig()
phi_k, (x, y), t = q(F.run_main_simulation, config=fwd, store_history=True, control_input=None, verbose=False)
phiT, phiQ = q(G.build_targets, x, y, t, phi_k[0].copy(), fwd.Lx, fwd.Ly, fwd.T, False, 1, 1)
u_k = np.zeros_like(phi_k)
cost_k = q(C.calculate_cost, phi_k, u_k, phiQ, phiT, x, y, t, opt)
costs, alphas, changes, ls = [cost_k], [], [], []
alpha_prev, plateau = opt.alpha_max, 0
t0 = time.time()
for k in range(K):
    _, _, r_k = B.run_backward(phi_k, x, y, t, fwd, opt.b1, opt.b2, phiQ, phiT)
    g = C.calculate_gradient(r_k, u_k, opt)
    u_o = C.proximal_step(u_k, g, alpha_prev, opt)
    phi_o, _, t_o = q(F.run_main_simulation, config=fwd, store_history=True, control_input=u_o, verbose=False)
    c_o = q(C.calculate_cost, phi_o, u_o, phiQ, phiT, x, y, t_o, opt)
    if c_o < cost_k:
        alpha_k, u_n, c_n, phi_n = alpha_prev, u_o, c_o, phi_o
        ls.append(0)
    else:
        alpha_k, u_n, c_n, phi_n, _, _, att = q(G.perform_backtracking_line_search_2D, u_k, cost_k, g, phiQ, phiT, x, y, fwd, opt,
                                                alpha_init=alpha_prev * 0.8)
        ls.append(att)
    costs.append(c_n); alphas.append(alpha_k)
    plateau = plateau + 1 if (k > 0 and abs(costs[-1] - costs[-2]) < 1e-5) else 0
    if plateau >= 5:
        alpha_prev, plateau = min(opt.alpha_max, alpha_k * 1.5), 0
    else:
        alpha_prev = min(opt.alpha_max, alpha_k * 1.2)
    changes.append(np.linalg.norm(u_n - u_k) / (np.linalg.norm(u_k) + 1e-9))
    u_k, cost_k, phi_k = u_n, c_n, phi_n
    print(f"iter {k+1}: J={c_n!r} alpha={alpha_k} ls={ls[-1]} change={changes[-1]:.6e} ({time.time()-t0:.0f}s)", flush=True)
    np.savez_compressed(os.path.join(os.path.dirname(HERE), "tests", "golden", "g2d_128_loop.npz"), J=np.array(costs), alpha=np.array(alphas),
                        change=np.array(changes), ls_attempts=np.array(ls), u_final_sub=u_k[[1, 50, 99]][:, ::4, ::4],
                        phi_final_T=phi_k[-1])
