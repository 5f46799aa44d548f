"""Forcing-term study for the forward Newton/BiCGStab solve (CPU, NumPy prototype of the CUDA algorithm: krylov_proto.py).

Usage: python scripts/krylov_forcing_study.py N M     (e.g. 128 200)
Runs M Crank-Nicolson steps of the N^2 default problem with every linear solve at 1e-11 ("strict"), then with the FIRST
solve of each Newton solve at 1e-4 .. 1e-7 and prints BiCGStab iterations, Newton solves and the largest relative
trajectory difference to the strict run.  Result that motivated vch2d_set_krylov_first (default 1e-6), 128^2 x 200 steps:
  first 1e-6: iterations x0.780, Newton solves unchanged, max rel. trajectory difference 2.1e-12
  first 1e-5: x0.737, 2.4e-11;   first 1e-4: x0.694, 6.0e-10;   first 1e-7: x0.826, 2.0e-13;   all 1e-9: x0.820, 9.4e-13
and on a 1024^2 state after 200 steps: x0.773, 2e-15 per step.  Not part of the product; no GPU needed."""
import sys, numpy as np, time
import krylov_proto as P
N = int(sys.argv[1]); M = int(sys.argv[2])
def run(sched, M, N, verbose=False):
    G = P.Grid(N)
    rng = np.random.default_rng(42)
    W = np.outer(P.trapz_w(N), P.trapz_w(N))
    phi = 0.1 * rng.standard_normal((N + 1, N + 1)); phi -= (phi * W).sum() / W.sum()
    mu = np.zeros_like(phi); w = np.zeros_like(phi); z = np.zeros_like(phi)
    m0 = (phi * W).sum() * G.h**2
    hist = [phi.copy()]; tot_its = 0; tot_newton = 0; rh = []
    for n in range(M):
        dt = G.dt; gdt = G.gamma / dt
        w1 = ((gdt - 0.5) * w + 0.5 * (z + z)) / (gdt + 0.5)
        lf = G.lap(phi)
        cphi = -G.tau * phi / dt - 0.5 * G.kappa * lf - 2 * G.c2 * phi - 0.5 * mu - 0.5 * (w1 + w)
        cmu = -phi / dt - 0.5 * G.lap(mu)
        f = phi.copy(); m = -G.kappa * lf + G.c1 * G.flog(phi) - 2 * G.c2 * phi - w1
        def resid(f, m):
            rp = G.tau / dt * f - 0.5 * G.kappa * G.lap(f) + G.c1 * G.flog(f) - 0.5 * m + cphi
            rm = f / dt - 0.5 * G.lap(m) + cmu
            return rp, rm
        rp, rm = resid(f, m); nr = np.sqrt(np.sum(rp**2) + np.sum(rm**2)); rs = [nr]
        for k in range(50):
            if nr < 1e-6: break
            a = G.tau / dt + 2 * G.c1 / (1 - np.minimum(f**2, G.dsq))
            b = G.lap(rp) - rm
            tol = sched(k, nr)
            d, its = P.bicgstab(G, a, b, np.sqrt(a.min() * a.max()), tol=tol)
            tot_its += its; tot_newton += 1
            dm = 2 * (a * d - 0.5 * G.kappa * G.lap(d) + rp)
            f = f + d; m = m + dm
            rp, rm = resid(f, m); nr = np.sqrt(np.sum(rp**2) + np.sum(rm**2)); rs.append(nr)
        rh.append(rs)
        f = np.clip(f, -0.99, 0.99)
        me = (f * W).sum() * G.h**2 - m0
        inter = np.abs(f) < 0.985; wi = (W * inter).sum() * G.h**2
        if wi > 0: f = np.where(inter, f - me / wi, f)
        phi, mu, w = f, m, w1
        hist.append(phi.copy())
    return np.array(hist), tot_its, tot_newton, rh

t0 = time.time()
H0, i0, n0, rh0 = run(lambda k, nr: 1e-11, M, N)
print('strict: its', i0, 'newton', n0, 'time', time.time() - t0)
for j in (0, 5, 50, M - 1):
    if j < M: print('  step', j, ['%.2e' % r for r in rh0[j]])
scheds = {
  'first1e-6': lambda k, nr: 1e-6 if k == 0 else 1e-11,
  'first1e-5': lambda k, nr: 1e-5 if k == 0 else 1e-11,
  'first1e-4': lambda k, nr: 1e-4 if k == 0 else 1e-11,
  'first1e-7': lambda k, nr: 1e-7 if k == 0 else 1e-11,
  'all1e-9': lambda k, nr: 1e-9,
}
for name, s in scheds.items():
    H, i, n, rh = run(s, M, N)
    e = [np.linalg.norm(H[j] - H0[j]) / np.linalg.norm(H0[j]) for j in range(1, M + 1)]
    print(f'{name}: its {i} ({i/i0:.3f}) newton {n} ({n/n0:.3f}) max rel traj err {max(e):.2e} final {e[-1]:.2e}')
    print('  step', M - 1, ['%.2e' % r for r in rh[M - 1]])
