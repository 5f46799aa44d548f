#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference from /root/reference.

Test infrastructure only.  Runs in the build container (the reference is Python and cannot
travel to the GPU box); the vectors it writes are committed under tests/golden/.

Each case drives the same sequence as one (or two) iterations of the reference PGD loop
(2D: src/2D/Vch_control_2D/GD2_configured.py:291-313, 1D: src/1D/Vch_control_1D/GD_1D.py:333-376)
through the reference's own public functions.  The only instrumentation is a wrapper around
`newton_raphson` that records what the reference computes but does not return
(mu_new, w_new, Newton residual history).

usage: python oracle/make_golden.py [case ...]     (default: all fast cases)
"""
import importlib, os, sys, time, io, contextlib
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
REF = "/root/reference/src"
os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache_golden")
sys.path.insert(0, os.path.join(HERE, "_mpl_shim"))


def _load(dim):
    """Import the reference package for `dim` ('1D'/'2D') by bare module names."""
    for m in ("config", "Forward_solver", "backward_solver", "cost_and_function", "GD_1D",
              "second_order_conditions", "Forward2_solver", "backward2_solver",
              "cost2_and_function", "GD2_configured", "second_order_conditions_2d",
              "visualization_3d"):
        sys.modules.pop(m, None)
    p = os.path.join(REF, dim, f"Vch_control_{dim}")
    sys.path[:] = [q for q in sys.path if not q.startswith(REF)]
    sys.path.insert(0, p)
    os.chdir("/tmp")


def _quiet(f, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return f(*a, **k)


def case_2d(name, cfg_kw, n_iter=2, keep=None, keep_pq=True):
    _load("2D")
    import Forward2_solver as F, backward2_solver as B, cost2_and_function as C, GD2_configured as G
    from config import ForwardSolverConfig, OptimizationConfig
    cfg, opt = ForwardSolverConfig(**cfg_kw), OptimizationConfig()
    rec = {}
    orig = F.newton_raphson

    def wrapped(phi_old, mu_old, w_old, w_new, *a, **k):
        phi_new, mu_new, hist = orig(phi_old, mu_old, w_old, w_new, *a, return_residual_history=True)
        rec["mu"].append(mu_new.copy()); rec["w"].append(w_new.copy()); rec["hist"].append(list(hist))
        return phi_new, mu_new

    def forward(u):
        rec.update(mu=[], w=[], hist=[])
        F.newton_raphson = wrapped
        try:
            phi, (x, y), t = _quiet(F.run_main_simulation, cfg, store_history=True, control_input=u, verbose=False)
        finally:
            F.newton_raphson = orig
        nres = np.array([len(h) for h in rec["hist"]], dtype=np.int64)
        lastres = np.array([h[-1] for h in rec["hist"]])
        return phi, x, y, t, np.array(rec["mu"]), np.array(rec["w"]), nres, lastres, list(rec["hist"])

    t0 = time.time()
    out = dict(cfg_json=np.array(cfg.model_dump_json()), opt_json=np.array(opt.model_dump_json()))
    phi, x, y, t, mu, w, nres, lastres, hists = forward(None)
    phiT, phiQ = _quiet(G.build_targets, x, y, t, phi[0].copy(), cfg.Lx, cfg.Ly, cfg.T, False, 1, 1)
    u = np.zeros_like(phi)
    J = [_quiet(C.calculate_cost, phi, u, phiQ, phiT, x, y, t, opt)]
    sl = slice(None) if keep is None else np.asarray(keep)
    out.update(x=x, y=y, t=t, phiT=phiT, keep=np.arange(len(t))[sl])
    out["newton_hist_step0"] = np.array(hists[0])
    alpha = opt.alpha_max
    for k in range(n_iter):
        # mu/w are recorded per step: entry s is the state at time level s+1
        ks = None if keep is None else np.clip(np.asarray(keep) - 1, 0, None)
        out[f"phi{k}"] = phi[sl]
        out[f"mu{k}"] = mu if ks is None else mu[ks]
        out[f"w{k}"] = w if ks is None else w[ks]
        out[f"nres{k}"] = nres; out[f"lastres{k}"] = lastres
        p, q, r = B.run_backward(phi, x, y, t, cfg, opt.b1, opt.b2, phiQ, phiT)
        if keep_pq:
            out[f"p{k}"] = p[sl]; out[f"q{k}"] = q[sl]
        out[f"r{k}"] = r[sl]
        g = C.calculate_gradient(r, u, opt)
        u = C.proximal_step(u, g, alpha, opt)
        out[f"u{k+1}"] = u[sl]
        out[f"alpha{k}"] = np.float64(alpha)
        out[f"change{k}"] = np.float64(0.0)
        phi, x, y, t, mu, w, nres, lastres, _ = forward(u)
        J.append(_quiet(C.calculate_cost, phi, u, phiQ, phiT, x, y, t, opt))
        print(f"[{name}] iter {k}: J {J[-2]!r} -> {J[-1]!r}  ({time.time()-t0:.1f}s)", flush=True)
        alpha = min(opt.alpha_max, alpha * 1.2)
    out[f"phi{n_iter}"] = phi[sl]; out[f"nres{n_iter}"] = nres; out[f"lastres{n_iter}"] = lastres
    out["J"] = np.array(J)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(f"[{name}] done in {time.time()-t0:.1f}s")


def case_1d(name, cfg_kw, n_iter=2):
    _load("1D")
    import Forward_solver as F, backward_solver as B, cost_and_function as C, GD_1D as G
    from config import ForwardSolverConfig, OptimizationConfig
    cfg, opt = ForwardSolverConfig(**cfg_kw), OptimizationConfig()
    rec = {}
    orig = F.newton_raphson

    def wrapped(phi_old, mu_old, w_old, w_new, *a, **k):
        phi_new, mu_new, hist = orig(phi_old, mu_old, w_old, w_new, *a, return_residual_history=True)
        rec["mu"].append(mu_new.copy()); rec["w"].append(w_new.copy()); rec["hist"].append(list(hist))
        return phi_new, mu_new

    def forward(u):
        rec.update(mu=[], w=[], hist=[])
        F.newton_raphson = wrapped
        try:
            phi, x, t = _quiet(F.run_main_simulation, cfg, store_history=True, control_input=u, verbose=False)
        finally:
            F.newton_raphson = orig
        return phi, x, t, np.array(rec["mu"]), np.array(rec["w"]), np.array([len(h) for h in rec["hist"]])

    out = dict(cfg_json=np.array(cfg.model_dump_json()), opt_json=np.array(opt.model_dump_json()))
    phi, x, t, mu, w, nres = forward(None)
    phiT, phiQ = _quiet(G.build_targets_1d, x=x, t_hist=t, phi_initial=phi[0].copy(), Lx=float(cfg.Lx),
                        T=float(cfg.T), interactive=False, choice_t=1, choice_q=1)
    u = np.zeros_like(phi)
    J = [_quiet(C.calculate_cost, phi, u, phiQ, phiT, x, t, opt.b1, opt.b2, opt.b3, opt.kappa_sparsity)]
    out.update(x=x, t=t, phiT=phiT, phiQ=phiQ)
    alpha = opt.alpha_max
    for k in range(n_iter):
        out[f"phi{k}"] = phi; out[f"mu{k}"] = mu; out[f"w{k}"] = w; out[f"nres{k}"] = nres
        p, q, r = B.run_backward(phi, x, t, opt.b1, opt.b2, phiQ, phiT)
        out[f"p{k}"] = p; out[f"q{k}"] = q; out[f"r{k}"] = r
        g = C.calculate_gradient(r, u, opt.b3)
        u = G.perform_proximal_and_projection(C.perform_gradient_step(u, g, alpha), alpha,
                                              opt.kappa_sparsity, opt.u_min, opt.u_max)
        out[f"u{k+1}"] = u
        phi, x, t, mu, w, nres = forward(u)
        J.append(_quiet(C.calculate_cost, phi, u, phiQ, phiT, x, t, opt.b1, opt.b2, opt.b3, opt.kappa_sparsity))
        print(f"[{name}] iter {k}: J {J[-2]!r} -> {J[-1]!r}", flush=True)
        alpha = min(opt.alpha_max, alpha * 1.2)
    out[f"phi{n_iter}"] = phi; out[f"mu{n_iter}"] = mu; out[f"w{n_iter}"] = w
    out["J"] = np.array(J)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(f"[{name}] done")


CASES = {
    "g1d_default": lambda: case_1d("g1d_default", {}),
    "g1d_n64": lambda: case_1d("g1d_n64", dict(N=64, T=0.3, dt_initial=1e-2)),
    "g2d_32": lambda: case_2d("g2d_32", dict(Nx=32, Ny=32, T=0.2, dt_initial=1e-2)),
    "g2d_rect": lambda: case_2d("g2d_rect", dict(Nx=12, Ny=20, Lx=1.0, Ly=1.5, T=0.05, dt_initial=1e-2)),
    "g2d_64": lambda: case_2d("g2d_64", dict(Nx=64, Ny=64, T=0.5, dt_initial=1e-2), keep=[0, 10, 25, 40, 50], keep_pq=False),
    # 256^2, two CN steps of the uncontrolled forward solve + adjoint over them (≈1.5 min): largest grid pinned to the reference
    "g2d_256": lambda: case_2d("g2d_256", dict(Nx=256, Ny=256, T=0.02, dt_initial=1e-2), n_iter=1, keep=[0, 2], keep_pq=False),
    # slow (≈10 min): the reference's default 2D config, sub-sampled in time
    "g2d_128": lambda: case_2d("g2d_128", {}, n_iter=1, keep=[0, 1, 50, 99, 100], keep_pq=False),
}



def case_soc_2d(name="g2d_32_soc"):
    """approximate_second_order_condition_2d of the UNMODIFIED reference (second_order_conditions_2d.py:120-235) at the second
    PGD iterate of g2d_32 (u1, r1, phi1): 3 critical-cone directions, epsilon 1e-3, seed 42."""
    _load("2D")
    import second_order_conditions_2d as S
    from config import ForwardSolverConfig, OptimizationConfig
    g = np.load(os.path.join(OUT, "g2d_32.npz"))
    cfg = ForwardSolverConfig(**{k: v for k, v in __import__("json").loads(str(g["cfg_json"])).items() if k in ForwardSolverConfig.model_fields})
    opt = OptimizationConfig()
    import GD2_configured as G
    phiT, phiQ = _quiet(G.build_targets, g["x"], g["y"], g["t"], g["phi0"][0].copy(), cfg.Lx, cfg.Ly, cfg.T, False, 1, 1)
    d2 = _quiet(S.approximate_second_order_condition_2d, g["u1"], g["r1"], g["phi1"], g["x"], g["y"], g["t"], opt_config=opt,
                phi_Q_target=phiQ, phi_T_target=phiT, u_min=opt.u_min, u_max=opt.u_max, num_directions=3, epsilon=1e-3, seed=42,
                fwd_config=cfg)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), d2=np.array(d2), epsilon=np.float64(1e-3), seed=np.int64(42))
    print(f"[{name}] d2 = {d2}")


CASES["g2d_32_soc"] = case_soc_2d


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    names = sys.argv[1:] or [c for c in CASES if c not in ("g2d_128", "g2d_256", "g2d_32_soc")]
    for n in names:
        CASES[n]()
