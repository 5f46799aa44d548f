"""Design study for round 2 (CPU, NumPy): a BiCGStab iteration with 6 instead of 7 kernel launches.

Today (csrc/vch2d.cu, enqueue_bicg_iteration):  rows[p = r + beta q] cols rows[v, (r0,v),(v,v),(r,v)]  rows[s = r - alpha v] cols
rows[t, (t,s),(t,t)]  update[x += alpha p + omega s, r = s - omega t, q = p - omega v, (r,r), (r0,r)].
The update kernel exists only because beta needs rho_new = (r0, r_new) and the stop test needs (r_new, r_new) — two grid-wide
sums over a vector that the update itself produces.  Both follow from dot products the SECOND epilogue can take while t is in
registers:        rho_new = (r0,s) - omega (r0,t),        (r,r) = (s,s) - 2 omega (t,s) + omega^2 (t,t),
so the x/r/p updates can move into the first row transform of the NEXT iteration (r = s - omega t, x += alpha p + omega s,
p = r + beta (p - omega v), all element-wise on vectors that kernel can read), plus one closing x-update per solve.

This script runs both recurrences on the Newton systems of a phase-separated state and prints iterations, the true relative
residual of the returned x, and how far the formula-based scalars drift from the directly computed ones.
Usage: python scripts/krylov_merged_study.py N state.npy   (state = np.stack([phi, mu, w]) as written by krylov_proto.py)"""
import sys
import numpy as np
import krylov_proto as P

N = int(sys.argv[1]); G = P.Grid(N)
phi, mu, w = np.load(sys.argv[2])
dt = G.dt
a = G.tau / dt + 2 * G.c1 / (1 - np.minimum(phi ** 2, G.dsq))
gdt = G.gamma / dt; w1 = ((gdt - 0.5) * w) / (gdt + 0.5); lf = G.lap(phi)
cphi = -G.tau * phi / dt - 0.5 * G.kappa * lf - 2 * G.c2 * phi - 0.5 * mu - 0.5 * (w1 + w); cmu = -phi / dt - 0.5 * G.lap(mu)
m = -G.kappa * lf + G.c1 * G.flog(phi) - 2 * G.c2 * phi - w1
rp = G.tau / dt * phi - 0.5 * G.kappa * G.lap(phi) + G.c1 * G.flog(phi) - 0.5 * m + cphi
rm = phi / dt - 0.5 * G.lap(m) + cmu
b = G.lap(rp) - rm
c0, c2, lam = 1 / dt, G.kappa / 2, G.lam
abar = np.sqrt(a.min() * a.max()); sym = c0 + lam * (abar + c2 * lam)
op = lambda x: x + G.spec((a - abar) * x, lam / sym)
A = lambda x: c0 * x - G.lap(a * x) + c2 * G.lap(G.lap(x))
pb = G.spec(b, 1 / sym); nb2 = pb.ravel() @ pb.ravel()
dot = lambda u, v: float(u.ravel() @ v.ravel())


def standard(tol):
    x = np.zeros_like(pb); r = pb.copy(); r0 = r.copy(); p = np.zeros_like(pb); v = np.zeros_like(pb)
    rho = alpha = omega = 1.0
    for it in range(1, 100):
        rho_new = dot(r0, r); beta = (rho_new / rho) * (alpha / omega); rho = rho_new
        p = r + beta * (p - omega * v); v = op(p); alpha = rho / dot(r0, v); s = r - alpha * v
        t = op(s); omega = dot(t, s) / dot(t, t); x += alpha * p + omega * s; r = s - omega * t
        if dot(r, r) <= tol ** 2 * nb2: return x, it
    return x, 99


def merged(tol, trust=1e6):
    """6-launch form: scalars of the NEXT iteration from dot products taken while s and t are at hand."""
    x = np.zeros_like(pb); r = pb.copy(); r0 = r.copy()
    p = r.copy(); rho = dot(r0, r); rr = rho; drift = 0.0
    for it in range(1, 100):
        v = op(p); alpha = rho / dot(r0, v); s = r - alpha * v
        t = op(s)
        ts, tt, ss, r0s, r0t = dot(t, s), dot(t, t), dot(s, s), dot(r0, s), dot(r0, t)     # ONE fused reduction (second epilogue)
        omega = ts / tt
        rho_new = r0s - omega * r0t                      # = (r0, s - omega t)
        rr_f = ss - 2 * omega * ts + omega ** 2 * tt     # = ||s - omega t||^2
        r_true = s - omega * t
        drift = max(drift, abs(rho_new - dot(r0, r_true)) / max(abs(dot(r0, r_true)), 1e-300),
                    abs(rr_f - dot(r_true, r_true)) / dot(r_true, r_true))
        stop = rr_f <= tol ** 2 * nb2 and ss <= trust * tol ** 2 * nb2      # formula trusted only near the threshold
        if stop:
            x += alpha * p + omega * s                   # closing update (one kernel per solve)
            return x, it, drift
        # --- what the first row transform of the next iteration does element-wise:
        beta = (rho_new / rho) * (alpha / omega); rho = rho_new
        x += alpha * p + omega * s
        r = s - omega * t
        p = r + beta * (p - omega * v)
    return x, 99, drift


for tol in (1e-6, 1e-11):
    x0, i0 = standard(tol); x1, i1, dr = merged(tol)
    res = lambda x: np.linalg.norm(A(x) - b) / np.linalg.norm(b)
    print(f"tol {tol:g}: standard {i0} its (true rel res {res(x0):.2e}) | merged {i1} its (true rel res {res(x1):.2e}), "
          f"x diff {np.linalg.norm(x1 - x0) / np.linalg.norm(x0):.2e}, max relative drift of the formula scalars {dr:.2e}")
