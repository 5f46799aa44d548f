"""Scratch: the reference's default 2D optimisation (128^2 x 100 steps) for K iterations through the drop-in driver."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "sparse-optimal-control-of-viscous-chan-hilliard-via-gradient-descent--1d-2d_b200")
sys.path.insert(0, PKG); sys.path.insert(0, os.path.join(PKG, "Vch_control_2D"))
import GD2_configured as G
from config import ForwardSolverConfig, OptimizationConfig
K = int(sys.argv[1]) if len(sys.argv) > 1 else 10
t0 = time.perf_counter()
res = G.optimize(ForwardSolverConfig(), OptimizationConfig(), 1, 1, max_iter=K, device_resident=True, verbose=False)
dt = time.perf_counter() - t0
print("cost history:", [float(f"{c:.12g}") for c in res["cost_history"]])
print("alpha:", res["alpha_history"]); print("ls calls/attempts:", res["timers"]["ls_calls"], res["timers"]["ls_attempts"])
print("tracking err:", [round(v, 6) for v in res["tracking_error_history"]]); print("terminal err:", [round(v, 6) for v in res["terminal_error_history"]])
print(f"{K} iterations in {dt:.2f}s ({dt/K:.3f} s/iteration incl. line-search forwards)")
