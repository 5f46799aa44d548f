"""Tolerances are BASELINE.json's: 1e-8 on trajectories, 1e-7 on gradient-like quantities and J (reduction order differs
between the decompositions, which moves Newton stop decisions at the rounding floor on grids >~ 600^2).

Slab-mode check (run under torchrun, one rank per GPU): every stage compares the rank's slab of the decomposed
problem with the same rows of the single-GPU solution computed by the same rank.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        scripts/slab_check.py [N] [M] [stage ...]
"""
import os, sys, time
import numpy as np, torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sparse-optimal-control-of-viscous-chan-hilliard-via-gradient-descent--1d-2d_b200"))
import vch_b200_native as nat

N = int(sys.argv[1]) if len(sys.argv) > 1 else 128
M = int(sys.argv[2]) if len(sys.argv) > 2 else 4
stages = sys.argv[3:] or ["selftest", "lap", "jac", "forward", "adjoint", "pgd", "host"]
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
dist.init_process_group("gloo", rank=rank, world_size=world)
dev = torch.cuda.current_device()
phys = dict(tau=0.05, gamma=10.0, c1=0.75, c2=1.0, kappa=1e-4)
h = 1.0 / N
slab = nat.SlabCtx2D.create_distributed(N, h, 1.0, phys["tau"], phys["gamma"], phys["c1"], phys["c2"], phys["kappa"], 1e-2)
full = nat.Ctx2D(N, N, h, h, 1.0, 1.0, phys["tau"], phys["gamma"], phys["c1"], phys["c2"], phys["kappa"], 1e-2, device=dev)
r0, nr = slab.row0, slab.rows
sl = slice(r0, r0 + nr)
def say(*a):
    sys.stdout.write(f"[rank {rank}] " + " ".join(str(v) for v in a) + "\n"); sys.stdout.flush()   # one write per line
def rel(a, b):
    a = a.double().reshape(-1); b = b.double().reshape(-1)
    return float((a - b).norm() / max(float(b.norm()), 1e-300))
worst = 0.0
def check(name, a, b, tol):
    global worst
    e = rel(a, b); worst = max(worst, e / tol)
    say(f"{name}: rel err {e:.3e} (tol {tol:g}) {'ok' if e <= tol else 'FAIL'}")
say(f"N={N} rows [{r0},{r0+nr}) of {N+1}; stages {stages}")
rng = np.random.default_rng(42)
phi0 = 0.1 * rng.standard_normal((N + 1, N + 1)); phi0 -= phi0.mean()
x = np.linspace(0, 1, N + 1)
xx, yy = np.meshgrid(x, x, indexing="ij")
g = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
phi0_d = g(phi0)

if "selftest" in stages:
    out = slab.selftest()
    exp = [world * (world + 1) / 2, 1.0, float(world), float(rank) if rank > 0 else 0.0, float(rank + 2) if rank < world - 1 else 0.0]
    say("selftest", out.tolist(), "expected", exp, "ok" if np.allclose(out, exp) else "FAIL")
    if not np.allclose(out, exp): worst = 1e9
if "lap" in stages:
    ref = full.apply_laplacian(phi0_d)
    mine = slab.apply_laplacian(phi0_d[sl].contiguous())
    check("laplacian", mine, ref[sl], 1e-15)
    w = torch.zeros_like(phi0_d)
    check("initialize_mu", slab.initialize_mu(phi0_d[sl].contiguous(), w[sl].contiguous()), full.initialize_mu(phi0_d, w)[sl], 1e-14)
if "jac" in stages:
    phi = g(0.6 * np.tanh(3 * np.sin(2 * np.pi * xx) * np.cos(np.pi * yy)) + 0.05 * phi0)
    Rp = g(rng.standard_normal((N + 1, N + 1))); Rm = g(rng.standard_normal((N + 1, N + 1)))
    d1, d2, its = full.jacobian_solve(phi, 1e-2, Rp, Rm)
    e1, e2, its2 = slab.jacobian_solve(phi[sl].contiguous(), 1e-2, Rp[sl].contiguous(), Rm[sl].contiguous())
    say(f"jacobian_solve its full {its} slab {its2}")
    check("jac dphi", e1, d1[sl], 1e-9); check("jac dmu", e2, d2[sl], 1e-9)
dts = np.full(M, 1e-2); t = 1e-2 * np.arange(M + 1)
phiT = g(0.7 * np.sin(2 * np.pi * xx) * np.cos(np.pi * yy))
if any(k in stages for k in ("forward", "adjoint", "pgd", "host")):
    u = g(0.5 * np.sin(np.pi * xx)[None] * np.ones((M + 1, 1, 1)))
    torch.cuda.synchronize(); t0 = time.perf_counter()
    hf, _, _ = full.forward(phi0_d, u, dts); torch.cuda.synchronize(); t1 = time.perf_counter()
    hs, _, _ = slab.forward(phi0_d[sl].contiguous(), u[:, sl].contiguous(), dts); torch.cuda.synchronize(); t2 = time.perf_counter()
    say(f"forward: full {t1-t0:.3f}s slab {t2-t1:.3f}s; stats full {full.last_stats.get('krylov_iterations')} slab {slab.last_stats.get('krylov_iterations')}")
    check("forward phi_hist", hs, hf[:, sl], 1e-8)
    s_ = g(t / t[-1])[:, None, None]
    phiQ = (1 - s_) * hf[0] + s_ * phiT
if "adjoint" in stages:
    pf, qf, rf = full.adjoint(hf, t, 5.0, 10.0, phiQ, phiT)
    ps, qs, rs = slab.adjoint(hf[:, sl].contiguous(), t, 5.0, 10.0, phiQ[:, sl].contiguous(), phiT[sl].contiguous())
    check("adjoint p", ps, pf[:, sl], 1e-8); check("adjoint q", qs, qf[:, sl], 1e-7); check("adjoint r", rs, rf[:, sl], 1e-7)
if "pgd" in stages:
    u0 = torch.zeros_like(hf)
    args = (5.0, 10.0, 1e-4, 1e-4, -1.0, 1.0, 50.0)
    u1, h1, J, red, st = full.pgd_iteration(u0, hf, phiQ, phiT, t, dts, x, x, *args)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    u1s, h1s, Js, reds, sts = slab.pgd_iteration(u0[:, sl].contiguous(), hf[:, sl].contiguous(), phiQ[:, sl].contiguous(),
                                                 phiT[sl].contiguous(), t, dts, x, x, *args)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    say(f"pgd slab {t1-t0:.3f}s J full {J[0]:.15g} slab {Js[0]:.15g}; red full {red.tolist()} slab {reds.tolist()}")
    check("pgd u_new", u1s, u1[:, sl], 1e-7); check("pgd phi_hist_new", h1s, h1[:, sl], 1e-8)
    if abs(Js[0] - J[0]) > 1e-9 * abs(J[0]): say("J mismatch FAIL"); worst = max(worst, 1e9)
    say("support identical:", bool(((u1s != 0) == (u1[:, sl] != 0)).all()))
if "host" in stages:    # host-buffer (streamed PCIe) entry point on the slab: NumPy in, NumPy out
    u0 = torch.zeros_like(hf)
    args = (5.0, 10.0, 1e-4, 1e-4, -1.0, 1.0, 50.0)
    u1d, h1d, Jd, redd, _ = slab.pgd_iteration(u0[:, sl].contiguous(), hf[:, sl].contiguous(), phiQ[:, sl].contiguous(),
                                               phiT[sl].contiguous(), t, dts, x, x, *args)
    nph = lambda a: np.ascontiguousarray(a.cpu().numpy())
    u1h, h1h, Jh, redh, _ = slab.pgd_iteration(nph(u0[:, sl]), nph(hf[:, sl]), nph(phiQ[:, sl]), nph(phiT[sl]), t, dts, x, x, *args)
    check("host-path u_new", torch.from_numpy(u1h), u1d.cpu(), 1e-14); check("host-path phi_hist_new", torch.from_numpy(h1h), h1d.cpu(), 1e-14)
    if abs(Jh[0] - Jd[0]) > 1e-14 * abs(Jd[0]): say("host-path J mismatch FAIL"); worst = max(worst, 1e9)
    slab.set_stream_budget(6 * 3 * 3 * 8 * slab.shape[0] * slab.shape[1])      # chunk rings of 3 levels: bounded-memory mode
    u1b, h1b, Jb, redb, _ = slab.pgd_iteration(nph(u0[:, sl]), nph(hf[:, sl]), nph(phiQ[:, sl]), nph(phiT[sl]), t, dts, x, x, *args)
    slab.set_stream_budget(0)
    check("bounded-memory u_new", torch.from_numpy(u1b), u1d.cpu(), 1e-14); check("bounded-memory phi_hist_new", torch.from_numpy(h1b), h1d.cpu(), 1e-14)
    if abs(Jb[0] - Jd[0]) > 1e-12 * abs(Jd[0]): say("bounded-memory J mismatch FAIL"); worst = max(worst, 1e9)
if "desync" in stages:   # a rank that never arrives: the waiting rank must fail with VCH_E_COMM after the wait limit, not hang
    if rank == 0:
        t0 = time.perf_counter()
        try:
            slab.selftest()
            say("desync: no error raised FAIL"); worst = 1e9
        except RuntimeError as exc:
            ok = "error 6" in str(exc) and time.perf_counter() - t0 < 60
            say(f"desync: raised after {time.perf_counter() - t0:.1f}s: {exc} {'ok' if ok else 'FAIL'}")
            if not ok: worst = 1e9
dist.barrier()
say("RESULT", "PASS" if worst <= 1.0 else "FAIL", f"(worst err/tol {worst:.2e})")
dist.destroy_process_group()
sys.exit(0 if worst <= 1.0 else 1)
