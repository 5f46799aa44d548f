#!/usr/bin/env python
"""bench.py — PGD iterations/s of the vCH sparse-control hot path on B200 (BASELINE.json metric).

One "step" = one optimistic PGD iteration (GD2_configured.py:299-313): adjoint sweep over the stored trajectory ->
gradient + soft-threshold prox -> forward Newton solve under the new control -> cost functional.
Workload (config 3 of BASELINE.json): 2D 1024^2 grid (1025^2 nodes), horizon M = 1000 CN steps (T = 10, dt = 1e-2),
default physics/weights of the reference's 2D config.py, synthetic targets built as GD2_configured.build_targets(1, 1).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--n 1024] [--horizon 1000]

N > 1 (torchrun, one rank per GPU): every rank owns an independent control problem of the same size (the path shards
across problems with no data-path collective; slab decomposition of one grid is DESIGN.md's next row) -> "weak" scaling;
the timed region is bracketed by barriers and the max over ranks is reported.
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


# ------------------------------------------------------------------------------------------------ CPU arm (oracle port)
def cpu_sample(sizes=(128, 256), target_n=1024, horizon=1000):
    """Time the reference algorithm (oracle port: SciPy SuperLU solves, same formulas) for one forward CN step and one
    adjoint step at each grid in `sizes`, fit t ~ nodes^p, extrapolate one PGD iteration at target_n^2 x horizon."""
    import vch_oracle as O
    pts = []
    for n in sizes:
        P = O.Phys2D(Nx=n, Ny=n)
        tf, ta, nsolve = O.time_step_sample_2d(P, n_steps=1)
        pts.append(((n + 1) ** 2, tf + ta, tf, ta, nsolve))
    if len(pts) > 1:
        p = float(np.log(pts[-1][1] / pts[0][1]) / np.log(pts[-1][0] / pts[0][0]))
    else:
        p = 1.5
    per_step = pts[-1][1] * ((target_n + 1) ** 2 / pts[-1][0]) ** p
    sec_per_iter = per_step * horizon
    return 1.0 / sec_per_iter, {"points": [{"grid": f"{int(round(a ** 0.5)) - 1}^2", "fwd_step_s": round(c, 3), "adj_step_s": round(d, 3),
                                            "newton_solves": e} for a, _, c, d, e in pts],
                                "exponent": round(p, 3), "sec_per_iteration_extrapolated": sec_per_iter}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.perf_counter()
    vals, info = [], None
    steps = max(1, min(args.steps, 2))
    for _ in range(steps):
        v, info = cpu_sample(target_n=args.n, horizon=args.horizon)
        vals.append(v)
    value = float(np.mean(vals))
    sample = ("oracle port of the reference algorithm (SciPy SuperLU spsolve, 1 thread): 1 forward CN step (Newton to 1e-6) + "
              "1 adjoint step at 128^2 and 256^2, power-law fit in nodes, extrapolated to "
              f"{args.n}^2 x {args.horizon} steps; {steps} samples; exponent {info['exponent']}")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": 0, "ms_per_step": 1e3 / value, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"2D {args.n}^2 grid, M={args.horizon} CN steps, reference 2D defaults, targets (1,1)"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample, "detail": info},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": time.perf_counter() - t0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ B200 arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    import vch_b200_native as nat
    import vch_oracle as O

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
    P = O.Phys2D(Nx=N, Ny=N, T=M * dt)
    Op = O.Opt2D()
    ctx = nat.Ctx2D(N, N, 1.0 / N, 1.0 / N, 1.0, 1.0, P.tau, P.gamma, P.c1, P.c2, P.kappa, device=local)
    dts = np.full(M, dt)
    t_hist = np.concatenate([[0.0], np.minimum(np.cumsum(dts), P.T)])
    x = np.linspace(0.0, 1.0, N + 1)
    n = (N + 1) ** 2
    field_bytes = 8 * n

    # synthetic problem: reference IC (host RNG, Forward2_solver.py:444-486; seed differs per rank = independent problems),
    # uncontrolled forward solve, targets as GD2_configured.build_targets(choice_t=1, choice_q=1)
    phi0 = torch.from_numpy(O.init_phi_2d(N, N, seed=42 + rank)).to(dev)
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
    if world > 1:
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
            with open(os.path.join(ROOT, "profiles", "r01_ncu_traffic.json")) as fh:
                traffic = json.load(fh)["kernels"].get("dct_rows_fft" if top.startswith("dct_rows_fft") else top) if N == 1024 else None
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

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu:
            v, info = cpu_sample(target_n=N, horizon=M)
            cpu = {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
                   "sample": "oracle port (SciPy SuperLU, 1 thread): 1 forward CN step + 1 adjoint step at 128^2 and 256^2, "
                             f"power-law fit in nodes (exponent {info['exponent']}), extrapolated to {N}^2 x {M} steps",
                   "detail": info}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": f"2D {N}^2 grid ({N+1}^2 nodes), M={M} CN steps (T={M*dt:g}), reference 2D defaults, "
                                       "targets build_targets(1,1), optimistic PGD iteration from u0=0",
                           "problems_per_gpu": 1, "l2": "inputs (8.4 GB trajectories) exceed the 126 MB L2; no flush needed",
                           "newton": "reference rule + fp64-floor stop (DESIGN.md)", "krylov_rel_tol": 1e-11,
                           "krylov_first_solve_rel_tol": float(os.environ.get("VCH_KRYLOV_FIRST_RTOL", 1e-6))},
                "gpu_launches": int(launches), "clocks": clocks, "e2e": e2e, "roofline": roof, "cpu_baseline": cpu,
                "solver": {"linear_solves_per_iteration": agg["newton_linear_solves"] / args.steps,
                           "newton_residual_evals_per_time_step": agg["newton_residual_evals"] / (args.steps * M),
                           "krylov_its_per_solve": agg["krylov_iterations"] / max(1, agg["newton_linear_solves"]),
                           "J_last": J_last}}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


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
    e0.record()
    for _ in range(args.steps):
        step()
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
                          "gpu_launches": int(ctx.launches() - l0), "sum_J": float(Jsum.item())}), flush=True)
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
    import vch_oracle as O
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    N, M, dt = args.n, args.horizon, 1e-2
    P = O.Phys2D(Nx=N, Ny=N, T=M * dt)
    Op = O.Opt2D()
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
    phi0 = torch.from_numpy(np.ascontiguousarray(O.init_phi_2d(N, N, seed=42)[r0:r0 + nr])).to(dev)   # every rank cuts the same global field
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
