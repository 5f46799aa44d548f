#!/usr/bin/env python
"""Large-grid golden vectors (test infrastructure only; slow — run in the background, results committed).

  g2d_512          the UNMODIFIED reference from /root/reference at 512^2: 3 Crank-Nicolson steps of the uncontrolled
                   forward solve, the adjoint sweep over them, one gradient + prox step (alpha = alpha_max) and the forward
                   solve + cost under the new control — one optimistic PGD iteration, GD2_configured.py:291-313.  The
                   fp64-floor rule of the CUDA library never fires at 512^2 (floor ~ 6e-7 * ... < 1e-6 tolerance margin is
                   checked by the GPU test through the Newton iteration counts), so this is a clean reference pin 4x
                   larger than g2d_256.  ~25 min (SuperLU, 1 thread).
  g2d_1024_oracle  the ORACLE (oracle/vch_oracle.py, SuperLU) at the benchmarked 1024^2 grid with the CUDA library's
                   floor-aware Newton stop written into it (newton_2d(floor_aware=True)): 2 CN steps, adjoint, prox,
                   forward + cost.  The reference's verbatim rule cannot terminate there (its tolerance 1e-6 is below the
                   fp64 resolution of the residual), which is why this case is anchored on the oracle — itself
                   pinned to the reference on every smaller grid — and not on the reference.  ~1.5-2 h, ~20 GB.

Fields are stored sub-sampled in space ([::stride, ::stride]) to keep the fixtures small, together with the l2 norm and the
plain sum of every FULL field, so global errors cannot hide between the sampled nodes.

usage: python oracle/make_golden_large.py g2d_512 | g2d_1024_oracle
"""
import io, os, sys, time, contextlib
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache_golden")


def rec(out, key, a, stride):
    a = np.asarray(a)
    out[key] = np.ascontiguousarray(a[..., ::stride, ::stride])
    out[key + "_norm"] = np.array([np.linalg.norm(x.ravel()) for x in a.reshape((-1,) + a.shape[-2:])])
    out[key + "_sum"] = np.array([x.sum() for x in a.reshape((-1,) + a.shape[-2:])])


def g2d_512():
    sys.path.insert(0, os.path.join(HERE, "_mpl_shim"))
    sys.path.insert(0, "/root/reference/src/2D/Vch_control_2D")
    os.chdir("/tmp")
    import Forward2_solver as F, backward2_solver as B, cost2_and_function as C, GD2_configured as G
    from config import ForwardSolverConfig, OptimizationConfig
    cfg, opt = ForwardSolverConfig(Nx=512, Ny=512, T=0.03, dt_initial=1e-2), OptimizationConfig()
    stride = 2
    hists = []
    orig = F.newton_raphson

    def wrapped(phi_old, mu_old, w_old, w_new, *a, **k):
        phi_new, mu_new, hist = orig(phi_old, mu_old, w_old, w_new, *a, return_residual_history=True)
        hists.append((mu_new.copy(), w_new.copy(), list(hist)))
        print(f"  newton: {len(hist)} evals, last |R| = {hist[-1]:.3e}  ({time.time()-t0:.0f}s)", flush=True)
        return phi_new, mu_new

    def quiet(f, *a, **k):
        with contextlib.redirect_stdout(io.StringIO()):
            return f(*a, **k)

    def forward(u):
        hists.clear()
        F.newton_raphson = wrapped
        try:
            phi, (x, y), t = F.run_main_simulation(cfg, store_history=True, control_input=u, verbose=False)
        finally:
            F.newton_raphson = orig
        return phi, x, y, t, [h[0] for h in hists], [h[1] for h in hists], [h[2] for h in hists]

    t0 = time.time()
    out = dict(cfg_json=np.array(cfg.model_dump_json()), opt_json=np.array(opt.model_dump_json()), stride=np.int64(stride))
    phi, x, y, t, mu, w, nh = forward(None)
    phiT, phiQ = quiet(G.build_targets, x, y, t, phi[0].copy(), cfg.Lx, cfg.Ly, cfg.T, False, 1, 1)
    u = np.zeros_like(phi)
    J = [quiet(C.calculate_cost, phi, u, phiQ, phiT, x, y, t, opt)]
    out.update(x=x, y=y, t=t)
    rec(out, "phi0", phi, stride); rec(out, "mu0", np.array(mu), stride); rec(out, "w0", np.array(w), stride)
    out["nres0"] = np.array([len(h) for h in nh]); out["lastres0"] = np.array([h[-1] for h in nh])
    out["hist0_step0"] = np.array(nh[0])
    p, q, r = B.run_backward(phi, x, y, t, cfg, opt.b1, opt.b2, phiQ, phiT)
    print(f"adjoint done ({time.time()-t0:.0f}s)", flush=True)
    rec(out, "p0", p, stride); rec(out, "r0", r, stride)
    g = C.calculate_gradient(r, u, opt)
    u = C.proximal_step(u, g, opt.alpha_max, opt)
    rec(out, "u1", u, stride)
    out["u1_support"] = np.array([(u != 0).sum()])
    phi, x, y, t, mu, w, nh = forward(u)
    J.append(quiet(C.calculate_cost, phi, u, phiQ, phiT, x, y, t, opt))
    rec(out, "phi1", phi, stride)
    out["nres1"] = np.array([len(h) for h in nh]); out["lastres1"] = np.array([h[-1] for h in nh])
    out["J"] = np.array(J)
    np.savez_compressed(os.path.join(OUT, "g2d_512.npz"), **out)
    print(f"[g2d_512] J {J}  done in {time.time()-t0:.0f}s", flush=True)


def g2d_1024_oracle():
    sys.path.insert(0, HERE)
    import vch_oracle as O
    P, opt = O.Phys2D(Nx=1024, Ny=1024, T=0.02, dt_initial=1e-2), O.Opt2D()
    stride = 4
    t0 = time.time()
    prog = lambda s, h: print(f"  step {s}: newton {len(h)} evals, |R| = {[f'{v:.3e}' for v in h]}  ({time.time()-t0:.0f}s)", flush=True)
    out = dict(cfg_json=np.array(__import__('json').dumps(O.asdict(P))), opt_json=np.array(__import__('json').dumps(O.asdict(opt))),
               stride=np.int64(stride), floor_aware=np.int64(1))
    fw = O.forward_2d(P, floor_aware=True, progress=prog)
    phi, x, y, t = fw["phi"], fw["x"], fw["y"], fw["t"]
    phiT, phiQ = O.targets_2d(x, y, t, phi[0], P.Lx, P.Ly, P.T)
    u = np.zeros_like(phi)
    J = [O.cost_2d(phi, u, phiQ, phiT, x, y, t, opt)[0]]
    out.update(x=x, y=y, t=t)
    rec(out, "phi0", phi, stride); rec(out, "mu0", fw["mu"], stride); rec(out, "w0", fw["w"], stride)
    out["nres0"] = fw["nres"]; out["lastres0"] = fw["lastres"]
    p, q, r = O.adjoint_2d(P, phi, x, y, t, opt.b1, opt.b2, phiQ, phiT)
    print(f"adjoint done ({time.time()-t0:.0f}s)", flush=True)
    rec(out, "p0", p, stride); rec(out, "r0", r, stride)
    u = O.soft_prox(u, r + opt.b3 * u, opt.alpha_max, opt.kappa_sparsity, opt.u_min, opt.u_max)
    rec(out, "u1", u, stride)
    out["u1_support"] = np.array([(u != 0).sum()])
    fw = O.forward_2d(P, u, floor_aware=True, progress=prog)
    J.append(O.cost_2d(fw["phi"], u, phiQ, phiT, x, y, fw["t"], opt)[0])
    rec(out, "phi1", fw["phi"], stride)
    out["nres1"] = fw["nres"]; out["lastres1"] = fw["lastres"]
    out["J"] = np.array(J)
    np.savez_compressed(os.path.join(OUT, "g2d_1024_oracle.npz"), **out)
    print(f"[g2d_1024_oracle] J {J}  done in {time.time()-t0:.0f}s", flush=True)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    {"g2d_512": g2d_512, "g2d_1024_oracle": g2d_1024_oracle}[sys.argv[1]]()
