#!/usr/bin/env python
"""Full-horizon parity of the default solver knobs against the strict solver on ONE GPU.

default: first linear solve of each Newton solve to 1e-6 (forcing term) + forward half-step exit
strict : every linear solve to 1e-11, no half-step exit        (VCH_KRYLOV_FIRST_RTOL=0 VCH_NO_HALF_EXIT=1)
Both keep the fp64-floor Newton stop (without it the reference rule cannot terminate at >= 1024^2).
Runs forward (u = 0) -> adjoint -> prox -> forward(u1) -> cost with both and prints relative differences as one JSON line.

usage: python scripts/parity_vs_strict.py [N=1024] [M=1000]
"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sparse-optimal-control-of-viscous-chan-hilliard-via-gradient-descent--1d-2d_b200"))
sys.path.insert(0, os.path.join(ROOT, "sparse-optimal-control-of-viscous-chan-hilliard-via-gradient-descent--1d-2d_b200", "Vch_control_2D"))
import torch
import vch_b200_native as nat


def run(N, M, strict, env=None):
    for k in ("VCH_KRYLOV_FIRST_RTOL", "VCH_NO_HALF_EXIT", "VCH_KRYLOV_RTOL"):
        os.environ.pop(k, None)
    if strict:
        os.environ["VCH_KRYLOV_FIRST_RTOL"] = "0"; os.environ["VCH_NO_HALF_EXIT"] = "1"
    os.environ.update(env or {})
    import Forward2_solver as F
    from config import ForwardSolverConfig, OptimizationConfig
    cfg, opt = ForwardSolverConfig(Nx=N, Ny=N, T=M * 1e-2), OptimizationConfig()
    ctx = nat.Ctx2D(N, N, 1.0 / N, 1.0 / N, 1.0, 1.0, cfg.tau, cfg.gamma, cfg.c1, cfg.c2, cfg.kappa)
    dev = torch.device("cuda", 0)
    dts = np.full(M, 1e-2)
    t = np.concatenate([[0.0], np.minimum(np.cumsum(dts), cfg.T)])
    x = np.linspace(0, 1, N + 1)
    phi0 = torch.from_numpy(F.init_phi_random(N, N, 1e-2, amp=0.1, seed=42)).to(dev)
    t0 = time.perf_counter()
    h0, _, _ = ctx.forward(phi0, None, dts); torch.cuda.synchronize()
    st_f = dict(ctx.last_stats)
    xx, yy = torch.meshgrid(torch.from_numpy(x).to(dev), torch.from_numpy(x).to(dev), indexing="ij")
    phiT = (0.7 * torch.sin(2 * np.pi * xx) * torch.cos(np.pi * yy)).contiguous()
    s = torch.from_numpy(t / cfg.T).to(dev)[:, None, None]
    phiQ = ((1 - s) * h0[0] + s * phiT).contiguous()
    u0 = torch.zeros_like(h0)
    r = torch.empty_like(h0)
    u1, h1, J, red, st = ctx.pgd_iteration(u0, h0, phiQ, phiT, t, dts, x, x, opt.b1, opt.b2, opt.b3, opt.kappa_sparsity,
                                            opt.u_min, opt.u_max, opt.alpha_max, r_out=r)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    del ctx
    return dict(h0=h0, r=r, u1=u1, h1=h1, J=J, wall=wall, fwd_stats=st_f, it_stats=st)


def rel(a, b):
    # per-level maximum of the relative l2 difference, and the whole-trajectory one
    d = (a - b).flatten(1).norm(dim=1); n = b.flatten(1).norm(dim=1).clamp_min(1e-300)
    return float((a - b).norm() / b.norm().clamp_min(1e-300)), float((d / n).max())


def study(N, M):
    """Several solver settings against the tightest one (rtol 1e-13 everywhere): how well is the trajectory defined at all?"""
    cfgs = [("ref_rtol1e-13", True, {"VCH_KRYLOV_RTOL": "1e-13"}),
            ("strict_rtol1e-12", True, {"VCH_KRYLOV_RTOL": "1e-12"}),
            ("strict_rtol1e-11", True, {}),
            ("half_exit_only", False, {"VCH_KRYLOV_FIRST_RTOL": "0"}),
            ("first1e-8_nohalf", False, {"VCH_KRYLOV_FIRST_RTOL": "1e-8", "VCH_NO_HALF_EXIT": "1"}),
            ("first1e-6_nohalf", False, {"VCH_NO_HALF_EXIT": "1"}),
            ("first1e-8_half", False, {"VCH_KRYLOV_FIRST_RTOL": "1e-8"}),
            ("default_first1e-6_half", False, {}),
            ("rtol1e-12_first1e-7_half", False, {"VCH_KRYLOV_RTOL": "1e-12", "VCH_KRYLOV_FIRST_RTOL": "1e-7"})]
    ref = None
    for name, strict, env in cfgs:
        a = run(N, M, strict, env)
        row = {"config": name, "wall_s": round(a["wall"], 3), "J": float(a["J"][0]),
               "krylov_its": a["fwd_stats"]["krylov_iterations"] + a["it_stats"]["krylov_iterations"],
               "solves": a["fwd_stats"]["newton_linear_solves"] + a["it_stats"]["newton_linear_solves"],
               "stalls": a["fwd_stats"]["krylov_stalls"] + a["it_stats"]["krylov_stalls"]}
        if ref is None:
            ref = a
        else:
            for k in ("h0", "r", "u1", "h1"):
                whole, worst = rel(a[k], ref[k])
                row[k] = [float(f"{whole:.3e}"), float(f"{worst:.3e}")]
            row["support_mismatch"] = int(((a["u1"] != 0) != (ref["u1"] != 0)).sum())
            row["J_rel"] = abs(row["J"] - float(ref["J"][0])) / abs(float(ref["J"][0]))
            del a
        torch.cuda.empty_cache()
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    M = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    if len(sys.argv) > 3 and sys.argv[3] == "study":
        study(N, M)
        sys.exit(0)
    a = run(N, M, strict=False)
    b = run(N, M, strict=True)
    out = {"N": N, "M": M}
    for k in ("h0", "r", "u1", "h1"):
        whole, worst = rel(a[k], b[k])
        out[k] = {"rel_l2": whole, "worst_level_rel_l2": worst}
    out["support_mismatch"] = int(((a["u1"] != 0) != (b["u1"] != 0)).sum())
    out["support_size"] = int((b["u1"] != 0).sum())
    out["J_default"] = float(a["J"][0]); out["J_strict"] = float(b["J"][0])
    out["J_rel"] = abs(out["J_default"] - out["J_strict"]) / abs(out["J_strict"])
    out["wall_default_s"] = a["wall"]; out["wall_strict_s"] = b["wall"]
    out["stats_default"] = {"forward0": a["fwd_stats"], "iteration": a["it_stats"]}
    out["stats_strict"] = {"forward0": b["fwd_stats"], "iteration": b["it_stats"]}
    print(json.dumps(out))
