#!/usr/bin/env python
"""bench.py — PGD iterations/s of the vCH sparse-control hot path on B200 (BASELINE.json metric).

One "step" = one optimistic PGD iteration (GD2_configured.py:299-313): adjoint sweep over the stored trajectory ->
gradient + soft-threshold prox -> forward Newton solve under the new control -> cost functional.
Workload (config 3 of BASELINE.json): 2D 1024^2 grid (1025^2 nodes), horizon M = 1000 CN steps (T = 10, dt = 1e-2),
default physics/weights of the reference's 2D config.py, synthetic targets built as GD2_configured.build_targets(1, 1).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--n 1024] [--horizon 1000]

N > 1 (torchrun, one rank per GPU): every rank owns an independent control problem of the same size (the path shards
across problems with no data-path collective) -> "weak" scaling; the timed region is bracketed by barriers and the max over
ranks is reported.  The same line also carries `slab_4096`: ONE 4096^2 problem row-slab decomposed over all ranks (BASELINE
config 5, strong scaling) next to the single-GPU time of the same problem measured by rank 0 on the same box, and, at N = 1,
`parity_vs_strict` (full-horizon parity of the timed solver settings against the tightest solver) and `cpu_baseline` with a
measured same-config CPU/GPU pair.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "sparse-optimal-control-of-viscous-chan-hilliard-via-gradient-descent--1d-2d_b200")
for p in (PKG, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "PGD iters/s (fwd Newton+adjoint+prox) 2D 1024^2"
UNIT = "it/s"

# Algorithmic bytes per node per launch (fp64; stencil neighbours count once) — SURVEY §8(d) / DESIGN.md §Kernels.
BYTES_PER_NODE = {
    "op_apply_fwd": 24, "op_apply_adj": 24,            # read x, a; write y
    "dct_rows_fft": 16, "dct_cols_fft_solve": 16,      # read + write the field once
    "rows16": 16, "cols16_solve": 16,                  # radix-16 kernels (vch_fft16.cuh): same accounting
    "rows16_pro": 64,                                  # average of mode 2 (read r, v, a; write s, transform = 40) and mode 3
                                                       # (read s, t, p, v, x, a; write x, r, p, transform = 80): one of each per iteration
    "rows16_epi1": 40, "rows16_epi4": 40,              # read transform, addend (= other), r0 / r; write v | t
    "bicg_close_kernel": 32,
    "dct_rows_fft_pro": 40,                            # read r, q|v, a; write p|s and the transform
    "dct_rows_fft_epi1": 40, "dct_rows_fft_epi2": 24,  # read transform input, addend (, r0, r); write v|t
    "residual_kernel": 56,                             # read phi, mu, cphi, cmu; write Rphi, Rmu, a
    "bicg_x_kernel": 72, "bicg_p_kernel": 32, "bicg_s_kernel": 24, "bicg_dot1_kernel": 16, "bicg_dot2_kernel": 16,
    "bicg_init_kernel": 24,                            # read r; write r0, x (adjoint: + write r)
    "schur_rhs_kernel": 24,
    "dmu_ceiling_kernel": 64,                          # read dphi, a, Rphi, phi, mu; write dmu, phi_trial, mu_trial
    "trial_kernel": 48,
    "step_setup_kernel": 56, "solve_w_kernel": 32, "clip_mass_kernel": 16, "mass_shift_kernel": 16,
    "adj_rhs_kernel": 64, "adj_qr_kernel": 40, "adj_terminal_rhs_kernel": 24, "mu_init_kernel": 24,
    # round-2 tile kernels (vch2d_tiles.cuh) and the fused start of a solve
    "residual_tile_kernel": 56,                        # read phi, mu, cphi, cmu; write Rphi, Rmu, a
    "dmu_close_tile_kernel": 88,                       # read x, p, s, a, Rphi, phi, mu; write x, dmu, phi_trial, mu_trial
    "step_setup_tile_kernel": 72,                      # read phi0, mu0, w0, u_n, u_n+1; write w1, cphi, cmu, mu guess
    "adj_rhs_tile_kernel": 80,                         # read p1, q1, phi1, phi0, Q1, Q0; write r, r0, x, a
    "adj_qr_tile_kernel": 48,                          # read p (solution), q1, r1; write p_k, q_k, r_k
    "rows16_schur": 24,                                # read Rphi, Rmu; write the transform
    "rows16_init": 32,                                 # read the transform; write r, r0, x
}


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "", 1).isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "", 1).isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[k] for r in self.rows if len(r) >= 6 for k in range(4) if r[2 + k].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def dt_sequence(T, dt):
    out, t = [], 0.0
    while t < T - 1e-10:
        d = min(dt, T - t)
        out.append(d)
        t += d
    return np.array(out)


# ------------------------------------------------------------------------------------------------ CPU arm
# The reference's own CPU implementation of the path, timed on the host cores of the box.  SuperLU (scipy.sparse.linalg.spsolve,
# Forward2_solver.py:370, backward2_solver.py:229) is sequential, so one core does the work whatever the machine has; the BLAS
# thread count is reported for completeness.  kind = "reference": the UNMODIFIED modules imported from /root/reference (build
# container only — the tree does not exist on the GPU box); kind = "port": oracle/vch_oracle.py, the restatement pinned to the
# reference by tests/test_oracle_golden.py.  A full 1024^2 horizon is ~55 h of SuperLU (one factorisation ~10 min), so the
# number at the metric's config is an extrapolation from measured per-step times at 128^2, 256^2 and 384^2 (power law in the
# node count, least squares over the three points) and is labelled as one; next to it stands a MEASURED complete PGD iteration
# at the reference's default grid (config 2, shortened horizon) that the GPU arm repeats on the same config.
def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        return max([int(d.get("num_threads", 1)) for d in threadpool_info()] or [1])
    except Exception:
        return None


def _fit_power(points):
    """Least-squares power law t = c * nodes^p through (nodes, seconds) points."""
    x = np.log([a for a, _ in points]); y = np.log([b for _, b in points])
    p, lc = np.polyfit(x, y, 1)
    return float(p), float(np.exp(lc))


def cpu_port_iteration(n=128, M=10):
    """One COMPLETE optimistic PGD iteration (adjoint + prox + forward + cost) of the oracle port at n^2 x M, from u0 = 0."""
    import vch_oracle as O
    P, Op = O.Phys2D(Nx=n, Ny=n, T=M * 1e-2), O.Opt2D()
    fw = O.forward_2d(P)                                             # set-up (untimed): the uncontrolled trajectory
    phiT, phiQ = O.targets_2d(fw["x"], fw["y"], fw["t"], fw["phi"][0], P.Lx, P.Ly, P.T)
    t0 = time.perf_counter()
    un, fw1, J, r = O.pgd_iter_2d(P, Op, np.zeros_like(fw["phi"]), fw["phi"], fw["t"], fw["x"], fw["y"], phiQ, phiT, Op.alpha_max)
    return time.perf_counter() - t0, float(J)


def cpu_reference_iteration(n=128, M=20):
    """The same through the UNMODIFIED reference (GD2_configured.py:299-313), when /root/reference is present."""
    import contextlib, io
    ref = "/root/reference/src/2D/Vch_control_2D"
    sys.path.insert(0, os.path.join(ROOT, "oracle", "_mpl_shim")); sys.path.insert(0, ref)
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache_bench")
    cwd = os.getcwd(); os.chdir("/tmp")
    try:
        import Forward2_solver as F, backward2_solver as B, cost2_and_function as C, GD2_configured as G
        from config import ForwardSolverConfig, OptimizationConfig
        cfg, opt = ForwardSolverConfig(Nx=n, Ny=n, T=M * 1e-2), OptimizationConfig()
        with contextlib.redirect_stdout(io.StringIO()):
            phi, (x, y), t = F.run_main_simulation(cfg, store_history=True, control_input=None, verbose=False)   # set-up + Numba JIT
            phiT, phiQ = G.build_targets(x, y, t, phi[0].copy(), cfg.Lx, cfg.Ly, cfg.T, False, 1, 1)
            u = np.zeros_like(phi)
            t0 = time.perf_counter()
            p, q, r = B.run_backward(phi, x, y, t, cfg, opt.b1, opt.b2, phiQ, phiT)
            u1 = C.proximal_step(u, C.calculate_gradient(r, u, opt), opt.alpha_max, opt)
            phi1, _, _ = F.run_main_simulation(cfg, store_history=True, control_input=u1, verbose=False)
            J = C.calculate_cost(phi1, u1, phiQ, phiT, x, y, t, opt)
            dt = time.perf_counter() - t0
        return dt, float(J)
    finally:
        os.chdir(cwd)


def cpu_arm(target_n, horizon, full, prefer_reference):
    """(it/s at target_n^2 x horizon [extrapolated], cpu_baseline dict).  full: the longer sample of --impl reference."""
    import vch_oracle as O
    M0 = 20 if full else 10
    kind = "port"
    if prefer_reference and os.path.isdir("/root/reference/src/2D/Vch_control_2D"):
        try:
            t_it, J = cpu_reference_iteration(128, M0)
            kind = "reference"
        except Exception as exc:   # fall back to the port, say so
            print(f"[bench] reference import failed ({exc!r}); using the oracle port", file=sys.stderr)
            t_it, J = cpu_port_iteration(128, M0)
    else:
        t_it, J = cpu_port_iteration(128, M0)
    pts = [(129 ** 2, t_it / M0)]
    detail = {"config2_iteration": {"grid": "128^2", "steps": M0, "seconds": round(t_it, 3), "J": J,
                                     "what": "one complete optimistic PGD iteration (adjoint sweep + gradient/prox + forward solve + cost), "
                                             f"reference default grid and weights, horizon {M0} of the default 100 steps"}}
    steps = []
    for n in (256, 384):
        tf, ta, ns = O.time_step_sample_2d(O.Phys2D(Nx=n, Ny=n), n_steps=1)
        pts.append(((n + 1) ** 2, tf + ta))
        steps.append({"grid": f"{n}^2", "fwd_step_s": round(tf, 3), "adj_step_s": round(ta, 3), "newton_solves": ns})
    p, c = _fit_power(pts)
    per_step = c * ((target_n + 1) ** 2) ** p
    sec = per_step * horizon
    detail.update({"single_step_samples": steps, "exponent": round(p, 3), "points_s_per_time_step": [[a, round(b, 4)] for a, b in pts],
                   "sec_per_iteration_extrapolated": sec, "extrapolated": True})
    base = {"value": 1.0 / sec, "unit": UNIT, "cores": 1, "kind": kind, "blas_threads": blas_threads(), "host_cpus": os.cpu_count(),
            "sample": (f"{'unmodified reference' if kind == 'reference' else 'oracle port of the reference'} (SciPy SuperLU, sequential): one complete "
                       f"PGD iteration at 128^2 x {M0} steps ({t_it:.1f} s, measured) + one forward and one adjoint time step at 256^2 and 384^2; "
                       f"power law in nodes (exponent {p:.2f}) extrapolated to {target_n}^2 x {horizon} steps — a full horizon there is "
                       f"~{sec / 3600:.0f} h of sparse LU"),
            "detail": detail}
    return 1.0 / sec, base


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.perf_counter()
    value, base = cpu_arm(args.n, args.horizon, full=True, prefer_reference=True)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": 1,
            "warmup": 0, "ms_per_step": 1e3 / value, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"2D {args.n}^2 grid ({args.n+1}^2 nodes), M={args.horizon} CN steps (T={args.horizon*1e-2:g}), reference 2D defaults, "
                                   "targets build_targets(1,1), optimistic PGD iteration from u0=0",
                       "note": "one bounded CPU sample regardless of --steps/--warmup (see cpu_baseline.sample); the value at this config is extrapolated"},
            "cpu_baseline": base,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": time.perf_counter() - t0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ B200 arm: extra legs
def _targets(torch, hist0, x, t, T, dev, row0=0, rows=None):
    """GD2_configured.build_targets(choice_t=1, choice_q=1) on the device (rows: slab of the x axis)."""
    xs = x if rows is None else x[row0:row0 + rows]
    xx, yy = torch.meshgrid(torch.from_numpy(xs).to(dev), torch.from_numpy(x).to(dev), indexing="ij")
    phiT = (0.7 * torch.sin(2 * np.pi * xx) * torch.cos(np.pi * yy)).contiguous()
    s = torch.from_numpy(t / T).to(dev)[:, None, None]
    return phiT, ((1 - s) * hist0[0] + s * phiT).contiguous()


def parity_vs_strict_leg(nat, F2, P, Op, N, M, dev):
    """Full-horizon parity of the solver settings the timed steps use (forcing term 1e-6 on the first linear solve of each Newton
    solve, forward half-step exit) against the tightest solver of the same library (every solve to 1e-13, no forcing term, no
    half-step exit): uncontrolled forward -> adjoint -> prox -> forward under the new control -> cost, same inputs.
    BASELINE tolerances: phi <= 1e-8, gradient and J <= 1e-7, identical support.  Both keep the fp64-floor Newton stop, without
    which the reference's rule cannot terminate at >= 1024^2; profiles/r02_parity_study_* holds the wider study."""
    import torch
    dts = np.full(M, 1e-2)
    t = np.concatenate([[0.0], np.minimum(np.cumsum(dts), P.T)])
    x = np.linspace(0.0, 1.0, N + 1)
    phi0 = torch.from_numpy(F2.init_phi_random(N, N, 1e-2, amp=0.1, seed=42)).to(dev)
    keys = ("VCH_KRYLOV_FIRST_RTOL", "VCH_NO_HALF_EXIT", "VCH_KRYLOV_RTOL")
    saved = {k: os.environ.get(k) for k in keys}

    def run(env):
        for k in keys:
            os.environ.pop(k, None)
        os.environ.update({k: v for k, v in saved.items() if v is not None})
        os.environ.update(env)
        c = nat.Ctx2D(N, N, 1.0 / N, 1.0 / N, 1.0, 1.0, P.tau, P.gamma, P.c1, P.c2, P.kappa, device=dev.index or 0)
        h0, _, _ = c.forward(phi0, None, dts)
        phiT, phiQ = _targets(torch, h0, x, t, P.T, dev)
        r = torch.empty_like(h0)
        u1, h1, J, _, st = c.pgd_iteration(torch.zeros_like(h0), h0, phiQ, phiT, t, dts, x, x, Op.b1, Op.b2, Op.b3, Op.kappa_sparsity,
                                           Op.u_min, Op.u_max, Op.alpha_max, r_out=r)
        torch.cuda.synchronize()
        del c, phiQ
        return {"phi0": h0, "r": r, "u1": u1, "phi1": h1, "J": float(J[0]), "its": int(st["krylov_iterations"])}

    try:
        a = run({})
        b = run({"VCH_KRYLOV_RTOL": "1e-13", "VCH_KRYLOV_FIRST_RTOL": "0", "VCH_NO_HALF_EXIT": "1"})
    finally:
        for k in keys:
            os.environ.pop(k, None)
        os.environ.update({k: v for k, v in saved.items() if v is not None})
    out = {"reference_solver": "same library, every linear solve to 1e-13, no forcing term, no half-step exit",
           "workload": f"{N}^2 x {M} steps: forward(u=0), adjoint, prox(alpha_max), forward(u1), cost"}
    for k in ("phi0", "r", "u1", "phi1"):
        d = (a[k] - b[k]).flatten(1).norm(dim=1); nb = b[k].flatten(1).norm(dim=1).clamp_min(1e-300)
        out[k] = {"rel_l2": float((a[k] - b[k]).norm() / b[k].norm().clamp_min(1e-300)), "worst_level_rel_l2": float((d / nb).max())}
    out["support_mismatch"] = int(((a["u1"] != 0) != (b["u1"] != 0)).sum())
    out["support_size"] = int((b["u1"] != 0).sum())
    out["J_rel"] = abs(a["J"] - b["J"]) / abs(b["J"])
    out["krylov_iterations"] = {"timed_settings": a["its"], "tightest": b["its"]}
    out["pass"] = bool(out["phi0"]["rel_l2"] <= 1e-8 and out["phi1"]["rel_l2"] <= 1e-8 and out["r"]["rel_l2"] <= 1e-7
                       and out["u1"]["rel_l2"] <= 1e-7 and out["J_rel"] <= 1e-7)
    del a, b
    torch.cuda.empty_cache()
    return out


def gpu_config2_iteration(nat, F2, Op, n, M, local):
    """One complete PGD iteration at the reference's default grid (n = 128), M steps, from u0 = 0, through the C ABI with HOST
    (NumPy) buffers — the same work cpu_port_iteration times on the CPU.  Returns (seconds, J)."""
    import torch
    from config import ForwardSolverConfig
    P = ForwardSolverConfig(Nx=n, Ny=n, T=M * 1e-2)
    c = nat.Ctx2D(n, n, 1.0 / n, 1.0 / n, 1.0, 1.0, P.tau, P.gamma, P.c1, P.c2, P.kappa, device=local)
    dts = np.full(M, 1e-2)
    t = np.concatenate([[0.0], np.cumsum(dts)])
    x = np.linspace(0.0, 1.0, n + 1)
    h0, _, _ = c.forward(F2.init_phi_random(n, n, 1e-2, amp=0.1, seed=42), None, dts)
    X, Y = np.meshgrid(x, x, indexing="ij")
    phiT = 0.7 * np.sin(2 * np.pi * X) * np.cos(np.pi * Y)
    s = (t / P.T)[:, None, None]
    phiQ = (1 - s) * h0[0] + s * phiT
    u0 = np.zeros_like(h0)
    best, J = None, None
    for _ in range(3):
        t0 = time.perf_counter()
        _, _, Jv, _, _ = c.pgd_iteration(u0, h0, phiQ, phiT, t, dts, x, x, Op.b1, Op.b2, Op.b3, Op.kappa_sparsity, Op.u_min, Op.u_max, Op.alpha_max)
        torch.cuda.synchronize()
        dtw = time.perf_counter() - t0
        best, J = (dtw if best is None else min(best, dtw)), float(Jv[0])
    return best, J


def slab_leg(nat, F2, P, Op, args, rank, world, local):
    """BASELINE config 5 next to the headline number: ONE 4096^2 control problem, M = args.slab_horizon CN steps.
    world > 1: the grid is row-slab decomposed over all ranks (slab mode: halo rows, DCT transposes and reduction partials travel
    through peer memory over NVLink, no host collective in the step) and rank 0 ALSO runs the same problem alone on its GPU, so
    the record holds the single-GPU time of the same box.  world == 1: the single-GPU time only."""
    import torch
    import torch.distributed as dist
    Ns, Ms, dt = args.slab_n, args.slab_horizon, 1e-2
    h = 1.0 / Ns
    dev = torch.device("cuda", local)
    dts = np.full(Ms, dt)
    T = Ms * dt
    t_hist = np.concatenate([[0.0], np.minimum(np.cumsum(dts), T)])
    x = np.linspace(0.0, 1.0, Ns + 1)
    phi_full = F2.init_phi_random(Ns, Ns, 1e-2, amp=0.1, seed=42)

    def timed(ctx, r0, nr, group_barrier):
        phi0 = torch.from_numpy(np.ascontiguousarray(phi_full[r0:r0 + nr])).to(dev)
        ha, _, _ = ctx.forward(phi0, None, dts)
        phiT, phiQ = _targets(torch, ha, x, t_hist, T, dev, r0, nr)
        st = {"u": torch.zeros_like(ha), "h": ha, "un": torch.empty_like(ha), "hn": torch.empty_like(ha)}
        rbuf = torch.empty_like(ha)
        res = {}

        def step():
            _, _, J, _, s = ctx.pgd_iteration(st["u"], st["h"], phiQ, phiT, t_hist, dts, x, x, Op.b1, Op.b2, Op.b3, Op.kappa_sparsity,
                                              Op.u_min, Op.u_max, Op.alpha_max, u_out=st["un"], phi_out=st["hn"], r_out=rbuf)
            st["u"], st["un"] = st["un"], st["u"]; st["h"], st["hn"] = st["hn"], st["h"]
            res["J"], res["stats"] = float(J[0]), s

        step()                                   # warm-up (graphs, staging)
        group_barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.slab_steps):
            step()
        e1.record()
        group_barrier()
        ms = e0.elapsed_time(e1) / args.slab_steps
        return ms, res

    def bar():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    out = {"workload": f"ONE 2D {Ns}^2 grid ({Ns+1}^2 nodes), M={Ms} CN steps, reference defaults, targets (1,1), chained optimistic PGD iterations",
           "steps": args.slab_steps, "n_gpus": world}
    try:
        ms1 = None
        if rank == 0:                             # single-GPU time of the same problem on this box
            c1 = nat.Ctx2D(Ns, Ns, h, h, 1.0, 1.0, P.tau, P.gamma, P.c1, P.c2, P.kappa, device=local)
            ms1, r1 = timed(c1, 0, Ns + 1, lambda: torch.cuda.synchronize())
            out["one_gpu"] = {"ms_per_pgd_iteration": ms1, "ms_per_time_step": ms1 / Ms, "J": r1["J"],
                              "krylov_iterations": int(r1["stats"]["krylov_iterations"]), "linear_solves": int(r1["stats"]["newton_linear_solves"])}
            del c1
            torch.cuda.empty_cache()
        if world > 1:
            bar()
            cs = nat.SlabCtx2D.create_distributed(Ns, h, 1.0, P.tau, P.gamma, P.c1, P.c2, P.kappa, 1e-2, device=local)
            msN, rN = timed(cs, cs.row0, cs.rows, bar)
            tm = torch.tensor([msN], device=dev, dtype=torch.float64)
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            if rank == 0:
                msN = float(tm.item())
                out["slab"] = {"ms_per_pgd_iteration": msN, "ms_per_time_step": msN / Ms, "it_per_s": 1e3 / msN, "J": rN["J"], "rows_per_gpu": cs.rows,
                               "krylov_iterations": int(rN["stats"]["krylov_iterations"]), "linear_solves": int(rN["stats"]["newton_linear_solves"]),
                               "krylov_stalls": int(rN["stats"]["krylov_stalls"])}
                out["speedup_vs_1gpu"] = ms1 / msN
                out["efficiency"] = ms1 / msN / world
                out["J_rel_diff_vs_1gpu"] = abs(rN["J"] - out["one_gpu"]["J"]) / abs(out["one_gpu"]["J"])
                out["scaling"] = "strong"
            del cs
        else:
            out["note"] = "single-GPU run: slab mode needs >= 2 ranks (its 2- and 4-rank parity tests are skipped on a 1-GPU box)"
    except Exception as exc:      # report, do not hide
        out["error"] = repr(exc)
    torch.cuda.empty_cache()
    return out if rank == 0 else None


# ------------------------------------------------------------------------------------------------ B200 arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    import vch_b200_native as nat
    # problem set-up through the product's own drop-in modules (reference names): config defaults and the host-RNG initial field
    sys.path.insert(0, os.path.join(PKG, "Vch_control_2D"))
    import Forward2_solver as F2
    from config import ForwardSolverConfig, OptimizationConfig

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nat.require_device()
    dev = torch.device("cuda", local)

    N, M, dt = args.n, args.horizon, 1e-2
    P = ForwardSolverConfig(Nx=N, Ny=N, T=M * dt)
    Op = OptimizationConfig()
    parity = parity_vs_strict_leg(nat, F2, P, Op, N, M, dev) if (rank == 0 and not args.no_parity) else None
    ctx = nat.Ctx2D(N, N, 1.0 / N, 1.0 / N, 1.0, 1.0, P.tau, P.gamma, P.c1, P.c2, P.kappa, device=local)
    dts = np.full(M, dt)
    t_hist = np.concatenate([[0.0], np.minimum(np.cumsum(dts), P.T)])
    x = np.linspace(0.0, 1.0, N + 1)
    n = (N + 1) ** 2
    field_bytes = 8 * n

    # synthetic problem: reference IC (host RNG, Forward2_solver.py:444-486; --rank-seeds: the same problem on every rank or one per rank),
    # uncontrolled forward solve, targets as GD2_configured.build_targets(choice_t=1, choice_q=1)
    phi0 = torch.from_numpy(F2.init_phi_random(N, N, 1e-2, amp=0.1, seed=42 + (rank if args.rank_seeds == "distinct" else 0))).to(dev)
    hist_a, _, _ = ctx.forward(phi0, None, dts)
    xx, yy = torch.meshgrid(torch.from_numpy(x).to(dev), torch.from_numpy(x).to(dev), indexing="ij")
    phiT = (0.7 * torch.sin(2 * np.pi * xx) * torch.cos(np.pi * yy)).contiguous()
    s = torch.from_numpy(t_hist / P.T).to(dev)[:, None, None]
    phiQ = ((1 - s) * hist_a[0] + s * phiT).contiguous()
    del s, xx, yy
    u_a = torch.zeros_like(hist_a)
    u_b, hist_b, r_buf = torch.empty_like(hist_a), torch.empty_like(hist_a), torch.empty_like(hist_a)

    alpha = Op.alpha_max
    state = {"u": u_a, "h": hist_a, "un": u_b, "hn": hist_b, "J": None, "stats": None}

    def step():
        _, _, J, red, st = ctx.pgd_iteration(state["u"], state["h"], phiQ, phiT, t_hist, dts, x, x, Op.b1, Op.b2, Op.b3,
                                             Op.kappa_sparsity, Op.u_min, Op.u_max, alpha, u_out=state["un"],
                                             phi_out=state["hn"], r_out=r_buf)
        state["u"], state["un"] = state["un"], state["u"]
        state["h"], state["hn"] = state["hn"], state["h"]
        state["J"], state["stats"] = J, st

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    # The e2e leg replays the SAME iterates as the timed steps (later PGD iterates cost more Krylov iterations, so timing
    # different iterates would not be comparable): snapshot the post-warm-up state into pinned host memory now.
    e2e_host, e2e_err = None, None
    if not args.no_e2e:
        try:
            need = 5 * (M + 1) * field_bytes                 # pinned host buffers of this rank
            with open("/proc/meminfo") as fh:
                avail = next(int(l.split()[1]) * 1024 for l in fh if l.startswith("MemAvailable"))
            fits = torch.tensor([1 if need * world <= 0.8 * avail else 0], device=dev)
            if world > 1:
                dist.all_reduce(fits, op=dist.ReduceOp.MIN)  # one decision for all ranks (collectives below must match)
            if int(fits.item()) == 0:                        # all ranks share one host; never drive the box out of memory
                raise MemoryError(f"e2e leg skipped: needs {need * world / 1e9:.0f} GB pinned host memory, {avail / 1e9:.0f} GB available")
            pin = lambda tns: tns.cpu().pin_memory()
            e2e_host = (pin(state["u"]), pin(state["h"]), pin(phiQ), pin(phiT))
        except Exception as exc:   # report, do not hide
            e2e_err = repr(exc)
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    l0 = ctx.launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    agg = {"newton_linear_solves": 0, "krylov_iterations": 0, "newton_residual_evals": 0}
    e0.record()
    for _ in range(args.steps):
        step()
        for k in agg:
            agg[k] += state["stats"][k]
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev, dtype=torch.float64)
    per_rank = None
    if world > 1:
        # every rank's own time and work next to the maximum that the metric uses
        mine = torch.tensor([float(ms.item()), float(agg["krylov_iterations"]) / args.steps, float(agg["newton_linear_solves"]) / args.steps],
                            device=dev, dtype=torch.float64)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = {"ms_per_step": [round(float(a[0]), 2) for a in allr], "krylov_iterations_per_step": [float(a[1]) for a in allr],
                    "linear_solves_per_step": [float(a[2]) for a in allr]}
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(ms.item())
    launches = ctx.launches() - l0
    value = world * 1e3 / ms_per_step

    J_last = None if state.get("J") is None else float(state["J"][0])

    # ---- roofline leg: per-kernel CUDA-event timing of a shortened iteration (first M_prof levels of the same arrays)
    roof = None
    if rank == 0:
        Mp = min(M, args.profile_steps)
        ctx.profile(True)
        ctx.pgd_iteration(state["u"][:Mp + 1], state["h"][:Mp + 1], phiQ[:Mp + 1], phiT, t_hist[:Mp + 1], dts[:Mp], x, x, Op.b1,
                          Op.b2, Op.b3, Op.kappa_sparsity, Op.u_min, Op.u_max, alpha, u_out=state["un"][:Mp + 1],
                          phi_out=state["hn"][:Mp + 1], r_out=r_buf[:Mp + 1])
        rep = ctx.profile_report()
        ctx.profile(False)
        tot = sum(v[0] for v in rep.values())
        peak, peak_src = load_peaks()
        table = {}
        for k, (tms, cnt) in sorted(rep.items(), key=lambda kv: -kv[1][0]):
            bpn = BYTES_PER_NODE.get(k)
            row = {"share": round(tms / tot, 4), "launches": cnt, "avg_us": round(1e3 * tms / cnt, 3)}
            if bpn and k not in ("grad_prox_kernel", "cost_kernel"):
                row["GBps"] = round(bpn * n / (tms / cnt * 1e-3) / 1e9, 1)
                row["frac"] = round(row["GBps"] / peak, 4)
            table[k] = row
        top = next(iter(table))
        bpn = BYTES_PER_NODE.get(top, 16)
        ach = bpn * n / (rep[top][0] / rep[top][1] * 1e-3) / 1e9
        traffic = None
        try:    # DRAM bytes per launch of that kernel from the committed ncu --set full capture (profiles/)
            with open(os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")) as fh:
                traffic = json.load(fh)["kernels"].get(top) if N == 1024 else None
        except Exception:
            traffic = None
        # whole iteration: sum of algorithmic bytes of every profiled launch over the summed kernel time
        all_bytes = sum(BYTES_PER_NODE.get(k, 0) * n * cnt for k, (tms, cnt) in rep.items())
        all_bytes += sum(24 * (Mp + 1) * n for k in rep if k in ("grad_prox_kernel", "cost_kernel"))
        roof = {"bound": "hbm", "kernel": top, "achieved": round(ach, 1), "peak": peak, "unit": "GB/s", "frac": round(ach / peak, 4),
                "frac_of_nominal_8000": round(ach / 8000.0, 4),
                "whole_iteration": {"algorithmic_GB": round(all_bytes / 1e9, 2), "GBps": round(all_bytes / (tot * 1e-3) / 1e9, 1),
                                    "frac": round(all_bytes / (tot * 1e-3) / 1e9 / peak, 4)},
                "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": bpn * n,
                "share_of_step": table[top]["share"], "profiled_steps": Mp, "kernels": table}

    # ---- end-to-end leg: the same call with HOST buffers (pinned), H2D/D2H inside the timed region
    e2e = None
    if not args.no_e2e and e2e_host is None:
        e2e = {"value": None, "unit": UNIT, "error": e2e_err}
    elif not args.no_e2e:
        lv = M + 1
        try:
            hu, hh, hq, hT = e2e_host
            hun, hhn = torch.empty_like(hu).pin_memory(), torch.empty_like(hu).pin_memory()
            # free the device-resident copies so the staged call has room at any horizon
            del u_a, u_b, hist_a, hist_b, r_buf
            state.clear()
            torch.cuda.empty_cache()
            reps = max(1, min(args.steps, args.e2e_steps))
            bounded = args.stream_budget_gb > 0          # out-of-core mode: chunk rings instead of whole trajectories on the device
            hr = torch.empty_like(hu).pin_memory() if bounded else None
            if bounded:
                ctx.set_stream_budget(int(args.stream_budget_gb * 1e9))

            hbuf = {"u": hu, "h": hh, "un": hun, "hn": hhn}

            def host_step():                                     # chained like the device steps: outputs become the next inputs
                out = ctx.pgd_iteration(hbuf["u"].numpy(), hbuf["h"].numpy(), hq.numpy(), hT.numpy(), t_hist, dts, x, x, Op.b1, Op.b2,
                                        Op.b3, Op.kappa_sparsity, Op.u_min, Op.u_max, alpha, u_out=hbuf["un"].numpy(),
                                        phi_out=hbuf["hn"].numpy(), r_out=hr.numpy() if bounded else None)
                hbuf["u"], hbuf["un"] = hbuf["un"], hbuf["u"]
                hbuf["h"], hbuf["hn"] = hbuf["hn"], hbuf["h"]
                return out
            # staging warm-up: a throw-away call that does not advance the iterates
            ctx.pgd_iteration(hu.numpy(), hh.numpy(), hq.numpy(), hT.numpy(), t_hist, dts, x, x, Op.b1, Op.b2, Op.b3,
                              Op.kappa_sparsity, Op.u_min, Op.u_max, alpha, u_out=hun.numpy(), phi_out=hhn.numpy(),
                              r_out=hr.numpy() if bounded else None)
            # PCIe yardstick: the adjoint sweep cannot start before its inputs arrive, so H2D bandwidth bounds the e2e number
            probe = torch.empty(min(hh.numel(), 1 << 28), dtype=torch.float64, device=dev)     # 2 GiB
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            probe.copy_(hh.view(-1)[:probe.numel()], non_blocking=True); torch.cuda.synchronize()
            ev0.record(); probe.copy_(hh.view(-1)[:probe.numel()], non_blocking=True); ev1.record(); torch.cuda.synchronize()
            h2d_gbps = probe.numel() * 8 / (ev0.elapsed_time(ev1) * 1e-3) / 1e9
            ev0.record(); hhn.view(-1)[:probe.numel()].copy_(probe, non_blocking=True); ev1.record(); torch.cuda.synchronize()
            d2h_gbps = probe.numel() * 8 / (ev0.elapsed_time(ev1) * 1e-3) / 1e9
            del probe
            barrier()
            t0 = time.perf_counter()
            for _ in range(reps):
                host_step()
            torch.cuda.synchronize()
            tt = torch.tensor([(time.perf_counter() - t0) / reps], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            e2e = {"value": world / float(tt.item()), "unit": UNIT, "h2d_bytes_per_step": int(((5 if bounded else 3) * lv + 1) * field_bytes),
                   "d2h_bytes_per_step": int((3 if bounded else 2) * lv * field_bytes + 9 * 8), "steps": reps,
                   "iterates": "the same PGD iterates as the device-timed steps (state snapshot taken after the warm-up)",
                   "pcie_GBps": {"h2d": round(h2d_gbps, 1), "d2h": round(d2h_gbps, 1),
                                 "note": "the adjoint sweep consumes phi_hist and phi_Q (2/3 of the H2D bytes) before anything else can run"},
                   "device_staging": (f"chunk rings, {args.stream_budget_gb:g} GB budget (phi_Q is uploaded twice, r makes a round trip)"
                                      if bounded else "whole trajectories"),
                   "call": "vch2d_pgd_iteration(mem=VCH_MEM_HOST): u, phi_hist, phi_Q, phi_T in; u_new, phi_hist_new, J, norms out"}
        except Exception as exc:   # report, do not hide
            e2e = {"value": None, "unit": UNIT, "error": repr(exc)}

    # ---- BASELINE config 5 on the same box: ONE 4096^2 problem over all ranks (slab mode) next to its single-GPU time
    slab = None
    if not args.no_slab:
        try:
            del hist_a, hist_b, u_a, u_b, r_buf
        except NameError:
            pass
        state.clear()
        torch.cuda.empty_cache()
        slab = slab_leg(nat, F2, P, Op, args, rank, world, local)
    ens1d = None
    if not args.no_ensemble:
        try:
            ens1d = ensemble_leg(nat, args, rank, world, local)
        except Exception as exc:   # report, do not hide
            ens1d = {"error": repr(exc)}
    conc = None
    if rank == 0 and world == 1 and not args.no_concurrent:
        try:
            conc = concurrent_leg(nat, F2, Op, N, dev)
        except Exception as exc:   # report, do not hide
            conc = {"error": repr(exc)}
    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu:
            _, cpu = cpu_arm(N, M, full=False, prefer_reference=False)
            # the SAME config on the GPU, end to end with host buffers: a measured (not extrapolated) ratio
            try:
                it = cpu["detail"]["config2_iteration"]
                gsec, gJ = gpu_config2_iteration(nat, F2, Op, 128, it["steps"], local)
                cpu["same_config"] = {"workload": f"2D 128^2 (reference default grid), M={it['steps']}, one complete PGD iteration from u0=0, host buffers",
                                      "cpu_s": it["seconds"], "gpu_s": round(gsec, 5), "ratio": round(it["seconds"] / gsec, 1),
                                      "J_cpu": it["J"], "J_gpu": gJ, "J_rel_diff": abs(gJ - it["J"]) / abs(it["J"])}
            except Exception as exc:
                cpu["same_config"] = {"error": repr(exc)}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": f"2D {N}^2 grid ({N+1}^2 nodes), M={M} CN steps (T={M*dt:g}), reference 2D defaults, "
                                       "targets build_targets(1,1), optimistic PGD iteration from u0=0",
                           "problems_per_gpu": 1, "rank_problems": ("one GPU" if world == 1 else (
                               "every rank solves its own copy of the same synthetic problem, no communication (per-GPU work fixed)" if args.rank_seeds == "same"
                               else "every rank solves a problem of its own (initial noise seeded by rank), no communication")),
                           "l2": "inputs (8.4 GB trajectories) exceed the 126 MB L2; no flush needed",
                           "newton": "reference rule + fp64-floor stop (DESIGN.md)", "krylov_rel_tol": 1e-11,
                           "krylov_first_solve_rel_tol": float(os.environ.get("VCH_KRYLOV_FIRST_RTOL", 1e-6))},
                "gpu_launches": int(launches), "per_rank": per_rank, "clocks": clocks, "e2e": e2e, "roofline": roof, "cpu_baseline": cpu,
                "parity_vs_strict": parity, "slab_4096": slab, "ensemble1d": ens1d, "concurrent_problems": conc,
                "solver": {"linear_solves_per_iteration": agg["newton_linear_solves"] / args.steps,
                           "newton_residual_evals_per_time_step": agg["newton_residual_evals"] / (args.steps * M),
                           "krylov_its_per_solve": agg["krylov_iterations"] / max(1, agg["newton_linear_solves"]),
                           "J_last": J_last}}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def _load_gd1d():
    """GD_1D beside the already imported 2D drop-ins: both packages have a module called `config`."""
    saved = sys.modules.pop("config", None)
    path = os.path.join(PKG, "Vch_control_1D")
    sys.path.insert(0, path)
    try:
        import GD_1D as G
    finally:
        sys.path.remove(path)
        if saved is not None:
            sys.modules["config"] = saved
    return G


def concurrent_leg(nat, F2, Op, N, dev, M=100, iters=2, kmax=3):
    """K independent problems of the metric's grid on ONE GPU at the same time (one library context, one stream and one host thread
    per problem) against the same problems one after the other: aggregate PGD iterations/s.  The kernels of the Krylov loop are
    single-wave latency chains, so further problems fill what one leaves idle.  NOT the headline (that stays one problem per GPU, the
    latency a user of the reference waits for); shortened horizon (M steps) so that the leg takes seconds."""
    import threading
    import torch
    from config import ForwardSolverConfig
    dt = 1e-2
    P = ForwardSolverConfig(Nx=N, Ny=N, T=M * dt)
    dts = np.full(M, dt)
    t_hist = np.concatenate([[0.0], np.minimum(np.cumsum(dts), P.T)])
    x = np.linspace(0.0, 1.0, N + 1)

    class Problem:
        def __init__(self, seed):
            self.stream = torch.cuda.Stream()
            with torch.cuda.stream(self.stream):
                self.ctx = nat.Ctx2D(N, N, 1.0 / N, 1.0 / N, 1.0, 1.0, P.tau, P.gamma, P.c1, P.c2, P.kappa, device=dev.index or 0)
                phi0 = torch.from_numpy(F2.init_phi_random(N, N, 1e-2, amp=0.1, seed=seed)).to(dev)
                self.h0, _, _ = self.ctx.forward(phi0, None, dts)
                self.phiT, self.phiQ = _targets(torch, self.h0, x, t_hist, P.T, dev)
                self.u0 = torch.zeros_like(self.h0)
                self.buf = [torch.empty_like(self.h0) for _ in range(5)]
            self.stream.synchronize()
            self.J = None

        def run(self, n_it):        # always from (u0, h0): iteration i reads the pair written by iteration i - 1
            u, h = self.u0, self.h0
            pairs = [(self.buf[0], self.buf[1]), (self.buf[2], self.buf[3])]
            with torch.cuda.stream(self.stream):
                for i in range(n_it):
                    un, hn = pairs[i & 1]
                    _, _, J, _, _ = self.ctx.pgd_iteration(u, h, self.phiQ, self.phiT, t_hist, dts, x, x, Op.b1, Op.b2, Op.b3, Op.kappa_sparsity,
                                                           Op.u_min, Op.u_max, Op.alpha_max, u_out=un, phi_out=hn, r_out=self.buf[4])
                    u, h = un, hn
                    self.J = float(J[0])
                self.stream.synchronize()

    probs = [Problem(42 + k) for k in range(kmax)]
    for pr in probs:
        pr.run(1)
    torch.cuda.synchronize()
    out = {"workload": f"{N}^2 x {M} steps, {iters} chained optimistic PGD iterations per problem from u0 = 0, K problems on one GPU", "K": {}}
    for K in range(1, kmax + 1):
        t0 = time.perf_counter()
        for pr in probs[:K]:
            pr.run(iters)
        torch.cuda.synchronize()
        seq = time.perf_counter() - t0
        Js = [pr.J for pr in probs[:K]]
        t0 = time.perf_counter()
        th = [threading.Thread(target=pr.run, args=(iters,)) for pr in probs[:K]]
        [t.start() for t in th]; [t.join() for t in th]
        torch.cuda.synchronize()
        con = time.perf_counter() - t0
        out["K"][str(K)] = {"one_after_the_other_it_per_s": round(K * iters / seq, 3), "concurrent_it_per_s": round(K * iters / con, 3),
                            "gain": round(seq / con, 3), "J_bit_identical": Js == [pr.J for pr in probs[:K]]}
    del probs
    torch.cuda.empty_cache()
    return out


def ensemble_leg(nat, args, rank, world, local, B=1024, warmup=2, steps=5):
    """BASELINE config 4 inside the default run, so that the driver's records carry it: B independent 1D control problems (default
    1D grid, varied targets / weights), one optimistic PGD iteration for the whole ensemble per step (4 launches), the batch split
    across the ranks without communication.  The iterate from u0 = 0 is repeated (`--workload ensemble1d` chains the iterates:
    same time per step)."""
    import torch
    import torch.distributed as dist
    G = _load_gd1d()
    dev = torch.device("cuda", local)
    cfg = G.ForwardSolverConfig()
    ens = G.make_ensemble(B)
    lo, hi = G.shard_range(B, rank, world)
    ctx = nat.Ctx1D(cfg.N, cfg.Lx / cfg.N, cfg.Lx, cfg.tau, cfg.gamma, cfg.c1, cfg.c2, cfg.kappa, device=local)
    up = lambda a: torch.from_numpy(np.ascontiguousarray(a[lo:hi])).to(dev)
    phi_init, phiQ, phiT = up(ens["phi_init"]), up(ens["phi_Q"]), up(ens["phi_T"])
    w = {k: np.ascontiguousarray(ens[k][lo:hi]) for k in ("b1", "b2", "b3", "ksp")}
    hist, _, _ = ctx.forward(phi_init, None, ens["dts"])
    u0 = torch.zeros_like(hist)
    run = lambda: G.optimistic_iteration_ensemble(ctx, ctx, u0, hist, phiQ, phiT, ens["x"], ens["t_hist"], ens["dts"], phi_init,
                                                  w["b1"], w["b2"], w["b3"], w["ksp"], 100.0)
    for _ in range(warmup):
        run()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    l0 = ctx.launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        J = run()[2]
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
    Jsum = torch.tensor([float(J[:, 0].sum())], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(Jsum, op=dist.ReduceOp.SUM)
    n1, lv = cfg.N + 1, len(ens["t_hist"])
    # algorithmic bytes of one iteration per problem: adjoint reads phi, Q, writes r (3 trajectories); prox reads u, r, writes u_new;
    # forward reads u_new, writes phi_new; cost reads phi_new, u_new, Q  ->  11 trajectory passes (the state lives in shared memory)
    gb = 11 * 8.0 * n1 * lv * B / 1e9
    return {"workload": f"{B} independent 1D problems (N = {cfg.N}, {lv - 2} steps), make_ensemble(seed 1234), one optimistic PGD iteration each from u0 = 0",
            "n_gpus": world, "problems_per_gpu": hi - lo, "ms_per_ensemble_iteration": float(ms.item()),
            "problem_it_per_s": B * 1e3 / float(ms.item()), "launches_per_iteration": int((ctx.launches() - l0) // steps),
            "sum_J": float(Jsum.item()), "scaling": "strong",
            "hbm_GBps_algorithmic": round(gb / (float(ms.item()) * 1e-3), 1),
            "bound": "latency of one CTA per problem (22 state vectors in shared memory, sequential time loop); HBM traffic is negligible"}


def run_ensemble1d(args):
    """Secondary workload (BASELINE config 4): B independent 1D control problems on the default 1D grid, varied targets and
    weights, one optimistic PGD iteration per step for the WHOLE ensemble (4 launches).  The batch is split across ranks
    with no communication (strong scaling: B is fixed); only the per-problem costs are gathered."""
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(PKG, "Vch_control_1D"))
    import vch_b200_native as nat
    import GD_1D as G
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    B = args.batch
    cfg = G.ForwardSolverConfig()
    ens = G.make_ensemble(B)
    lo, hi = G.shard_range(B, rank, world)
    ctx = nat.Ctx1D(cfg.N, cfg.Lx / cfg.N, cfg.Lx, cfg.tau, cfg.gamma, cfg.c1, cfg.c2, cfg.kappa, device=local)
    up = lambda a: torch.from_numpy(np.ascontiguousarray(a[lo:hi])).to(dev)
    phi_init, phiQ, phiT = up(ens["phi_init"]), up(ens["phi_Q"]), up(ens["phi_T"])
    w = {k: np.ascontiguousarray(ens[k][lo:hi]) for k in ("b1", "b2", "b3", "ksp")}
    hist, _, _ = ctx.forward(phi_init, None, ens["dts"])
    state = {"u": torch.zeros_like(hist), "h": hist, "J": None}

    def step():
        u1, h1, J, red, _ = G.optimistic_iteration_ensemble(ctx, ctx, state["u"], state["h"], phiQ, phiT, ens["x"], ens["t_hist"],
                                                            ens["dts"], phi_init, w["b1"], w["b2"], w["b3"], w["ksp"], 100.0)
        state.update(u=u1, h=h1, J=J)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    l0 = ctx.launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    per_step = []
    e0.record()
    for _ in range(args.steps):
        t0 = time.perf_counter()
        step()
        per_step.append((time.perf_counter() - t0) * 1e3)     # every library call of a step returns synchronised
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev, dtype=torch.float64)
    Jsum = torch.tensor([float(state["J"][:, 0].sum())], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(Jsum, op=dist.ReduceOp.SUM)
    if rank == 0:
        print(json.dumps({"metric": "1D ensemble problem-iterations/s (config 4)", "value": B * 1e3 / float(ms.item()), "unit": "problem-it/s",
                          "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(ms.item()),
                          "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                          "config": {"workload": f"{B} independent 1D problems (N=128, 100 steps), make_ensemble(seed 1234), one optimistic PGD iteration each",
                                     "problems_per_gpu": hi - lo},
                          "gpu_launches": int(ctx.launches() - l0), "sum_J": float(Jsum.item()),
                          "ms_per_step_median": float(np.median(per_step)), "ms_per_step_min": float(np.min(per_step)),
                          "ms_per_step_max": float(np.max(per_step))}), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_slab2d(args):
    """Secondary workload (BASELINE config 5): ONE 2D control problem decomposed into row slabs over the ranks (strong
    scaling: the problem is fixed).  The kernels exchange halo rows, preconditioner transposes and reduction partials through
    peer memory over NVLink (CUDA IPC); there is no host-side collective in the step.  At 1 rank the ordinary context runs the
    same problem."""
    import torch
    import torch.distributed as dist
    import vch_b200_native as nat
    sys.path.insert(0, os.path.join(PKG, "Vch_control_2D"))
    import Forward2_solver as F2
    from config import ForwardSolverConfig, OptimizationConfig
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    N, M, dt = args.n, args.horizon, 1e-2
    P = ForwardSolverConfig(Nx=N, Ny=N, T=M * dt)
    Op = OptimizationConfig()
    h = 1.0 / N
    if world > 1:
        ctx = nat.SlabCtx2D.create_distributed(N, h, 1.0, P.tau, P.gamma, P.c1, P.c2, P.kappa, 1e-2, device=local)
        r0, nr = ctx.row0, ctx.rows
    else:
        ctx = nat.Ctx2D(N, N, h, h, 1.0, 1.0, P.tau, P.gamma, P.c1, P.c2, P.kappa, device=local)
        r0, nr = 0, N + 1
    dts = np.full(M, dt)
    t_hist = np.concatenate([[0.0], np.minimum(np.cumsum(dts), P.T)])
    x = np.linspace(0.0, 1.0, N + 1)
    phi0 = torch.from_numpy(np.ascontiguousarray(F2.init_phi_random(N, N, 1e-2, amp=0.1, seed=42)[r0:r0 + nr])).to(dev)   # every rank cuts the same global field
    hist_a, _, _ = ctx.forward(phi0, None, dts)
    xx, yy = torch.meshgrid(torch.from_numpy(x[r0:r0 + nr]).to(dev), torch.from_numpy(x).to(dev), indexing="ij")
    phiT = (0.7 * torch.sin(2 * np.pi * xx) * torch.cos(np.pi * yy)).contiguous()
    s = torch.from_numpy(t_hist / P.T).to(dev)[:, None, None]
    phiQ = ((1 - s) * hist_a[0] + s * phiT).contiguous()
    del s, xx, yy
    state = {"u": torch.zeros_like(hist_a), "h": hist_a, "un": torch.empty_like(hist_a), "hn": torch.empty_like(hist_a), "J": None, "stats": None}
    r_buf = torch.empty_like(hist_a)

    def step():
        _, _, J, red, st = ctx.pgd_iteration(state["u"], state["h"], phiQ, phiT, t_hist, dts, x, x, Op.b1, Op.b2, Op.b3,
                                             Op.kappa_sparsity, Op.u_min, Op.u_max, Op.alpha_max, u_out=state["un"],
                                             phi_out=state["hn"], r_out=r_buf)
        state["u"], state["un"] = state["un"], state["u"]
        state["h"], state["hn"] = state["hn"], state["h"]
        state["J"], state["stats"] = J, st

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    l0 = ctx.launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        st = state["stats"]
        print(json.dumps({"metric": "PGD iterations/s on ONE slab-decomposed 2D problem (config 5)", "value": 1e3 / float(ms.item()), "unit": "it/s",
                          "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(ms.item()),
                          "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                          "config": {"workload": f"ONE 2D {N}^2 grid ({N+1}^2 nodes), M={M} CN steps, row slabs over {world} GPU(s), reference 2D "
                                                 "defaults, targets (1,1), chained optimistic PGD iterations from u0=0",
                                     "rows_per_gpu": nr, "trajectory_GB_per_gpu": round(8.0 * (M + 1) * nr * (N + 1) / 1e9, 3),
                                     "cache": "per-GPU trajectories exceed L2 (no flush needed)"},
                          "gpu_launches": int(ctx.launches() - l0), "clocks": clocks, "J": float(state["J"][0]),
                          "ms_per_time_step": float(ms.item()) / M,
                          "solver": {k: st[k] for k in ("newton_residual_evals", "newton_linear_solves", "krylov_iterations", "krylov_stalls")}}),
              flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", "--grid", dest="n", type=int, default=1024, help="grid intervals per side (--grid: spelling that torchrun does not mistake for one of its own options)")
    ap.add_argument("--horizon", type=int, default=1000, help="CN time steps M")
    ap.add_argument("--profile-steps", type=int, default=20)
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--stream-budget-gb", type=float, default=0.0,
                    help="e2e leg: cap the device staging (vch2d_set_stream_budget) -> bounded-memory chunk-ring mode; 0 = whole trajectories")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the full-horizon parity_vs_strict leg")
    ap.add_argument("--no-slab", action="store_true", help="skip the 4096^2 slab_4096 leg")
    ap.add_argument("--no-ensemble", action="store_true", help="skip the 1D ensemble leg (BASELINE config 4)")
    ap.add_argument("--rank-seeds", default="same", choices=["same", "distinct"],
                    help="N > 1: every rank solves its own copy of the SAME synthetic problem (per-GPU work fixed as N grows: weak scaling) or a "
                         "problem of its own (different initial noise: the iteration counts, hence the work, differ from rank to rank)")
    ap.add_argument("--no-concurrent", action="store_true", help="skip the concurrent-problems leg (K problems on one GPU, N = 1 only)")
    ap.add_argument("--slab-n", type=int, default=4096)
    ap.add_argument("--slab-horizon", type=int, default=20)
    ap.add_argument("--slab-steps", type=int, default=2)
    ap.add_argument("--workload", default="pgd2d", choices=["pgd2d", "ensemble1d", "slab2d"],
                    help="pgd2d = BASELINE metric (default); ensemble1d = config 4; slab2d = config 5 (one problem over all ranks)")
    ap.add_argument("--batch", type=int, default=1024, help="ensemble1d: number of problems")
    args = ap.parse_args()
    if args.workload == "ensemble1d":
        run_ensemble1d(args)
    elif args.workload == "slab2d":
        run_slab2d(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
