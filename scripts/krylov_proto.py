"""NumPy prototype of the DCT-preconditioned Newton/BiCGStab forward step (study of the Krylov iteration count)."""
import sys, time, numpy as np, scipy.fft as sf

def lam1d(N, h):
    k = np.arange(N + 1)
    return 4.0 / h**2 * np.sin(np.pi * k / (2 * N))**2

class Grid:
    def __init__(s, N, kappa=1e-4, tau=0.05, gamma=10.0, c1=0.75, c2=1.0, dt=1e-2):
        s.N = N; s.h = 1.0 / N; s.kappa = kappa; s.tau = tau; s.gamma = gamma; s.c1 = c1; s.c2 = c2; s.dt = dt
        l = lam1d(N, s.h); s.lam = l[:, None] + l[None, :]
        s.eps = 5e-3; s.dsq = 1 - 1e-4
    def lap(s, v):
        p = np.pad(v, 1, mode='reflect')
        return (p[2:, 1:-1] + p[:-2, 1:-1] + p[1:-1, 2:] + p[1:-1, :-2] - 4 * v) / s.h**2
    def flog(s, f):
        fs = np.clip(f, -(1 - s.eps), 1 - s.eps)
        return np.log((1 + fs) / (1 - fs))
    def spec(s, x, fac):
        return sf.idctn(sf.dctn(x, type=1, workers=8) * fac, type=1, workers=8)

def bicgstab(G, a, b, abar, tol=1e-11, maxit=200):
    c0 = 1.0 / G.dt; c2 = G.kappa / 2
    sym = c0 + G.lam * (abar + c2 * G.lam)
    fP = 1.0 / sym; fK = G.lam / sym
    am = a - abar
    op = lambda x: x + G.spec(am * x, fK)
    pb = G.spec(b, fP)
    x = np.zeros_like(b); r = pb.copy(); r0 = r.copy(); nb = np.linalg.norm(pb)
    rho = alpha = omega = 1.0; v = np.zeros_like(b); p = np.zeros_like(b)
    for it in range(1, maxit + 1):
        rho_new = np.vdot(r0, r); beta = (rho_new / rho) * (alpha / omega); rho = rho_new
        p = r + beta * (p - omega * v)
        v = op(p); alpha = rho / np.vdot(r0, v)
        s_ = r - alpha * v
        if np.linalg.norm(s_) <= tol * nb:
            x += alpha * p; return x, it - 0.5
        t = op(s_); omega = np.vdot(t, s_) / np.vdot(t, t)
        x += alpha * p + omega * s_; r = s_ - omega * t
        if np.linalg.norm(r) <= tol * nb: return x, it
    return x, maxit

ABAR = {
    'geo': lambda a: np.sqrt(a.min() * a.max()),
    'mid': lambda a: 0.5 * (a.min() + a.max()),
    'mean': lambda a: a.mean(),
    'hmean': lambda a: 1.0 / np.mean(1.0 / a),
    'rms': lambda a: np.sqrt(np.mean(a * a)),
    'median': lambda a: np.median(a),
}

def step(G, phi0, mu0, w0, u0, u1, mode='geo', stats=None, floor=True):
    dt = G.dt; gdt = G.gamma / dt
    w1 = ((gdt - 0.5) * w0 + 0.5 * (u0 + u1)) / (gdt + 0.5)
    lf = G.lap(phi0)
    cphi = -G.tau * phi0 / dt - 0.5 * G.kappa * lf - 2 * G.c2 * phi0 - 0.5 * mu0 - 0.5 * (w1 + w0)
    cmu = -phi0 / dt - 0.5 * G.lap(mu0)
    phi = phi0.copy(); mu = -G.kappa * lf + G.c1 * G.flog(phi0) - 2 * G.c2 * phi0 - w1
    def resid(phi, mu):
        rp = G.tau / dt * phi - 0.5 * G.kappa * G.lap(phi) + G.c1 * G.flog(phi) - 0.5 * mu + cphi
        rm = phi / dt - 0.5 * G.lap(mu) + cmu
        return rp, rm
    rp, rm = resid(phi, mu); nr = np.sqrt(np.sum(rp**2) + np.sum(rm**2))
    for nit in range(50):
        fl = 2.2e-16 * (2 / G.h**2) * np.linalg.norm(mu) if floor else 0.0
        if nr < 1e-6 or nr <= 1.5 * fl: break
        a = G.tau / dt + 2 * G.c1 / (1 - np.minimum(phi**2, G.dsq))
        b = G.lap(rp) - rm
        if stats is not None and 'snap' in stats: stats['snap'].append((a.copy(), b.copy()))
        dphi, its = bicgstab(G, a, b, ABAR[mode](a))
        if stats is not None: stats['its'].append(its)
        dmu = 2 * (a * dphi - 0.5 * G.kappa * G.lap(dphi) + rp)
        al = 1.0
        pos = dphi > 0; neg = dphi < 0
        amax = 2.0
        if pos.any(): amax = min(amax, 0.9 * np.min((0.99 - phi[pos]) / dphi[pos]))
        if neg.any(): amax = min(amax, 0.9 * np.min((-0.99 - phi[neg]) / dphi[neg]))
        al = min(1.0, amax)
        for _ in range(13):
            pt, mt = phi + al * dphi, mu + al * dmu
            rpt, rmt = resid(pt, mt); nt = np.sqrt(np.sum(rpt**2) + np.sum(rmt**2))
            if nt <= (1 - 1e-4 * al) * nr: break
            al *= 0.5
        nr_old = nr
        phi, mu, rp, rm, nr = pt, mt, rpt, rmt, nt
        if floor and nr > 0.5 * nr_old and nr < 50 * fl: break
    if stats is not None: stats['newton'].append(nit)
    phi = np.clip(phi, -0.99, 0.99)
    return phi, mu, w1

def trapz_w(N):
    w = np.ones(N + 1); w[0] = w[-1] = 0.5; return w

if __name__ == '__main__':
    N = int(sys.argv[1]); M = int(sys.argv[2]); mode = sys.argv[3] if len(sys.argv) > 3 else 'geo'
    snaps = [int(s) for s in sys.argv[4].split(',')] if len(sys.argv) > 4 else []
    G = Grid(N)
    rng = np.random.default_rng(42)
    phi = 0.1 * rng.standard_normal((N + 1, N + 1)); W = np.outer(trapz_w(N), trapz_w(N)); phi -= (phi * W).sum() / W.sum()
    mu = np.zeros_like(phi); w = np.zeros_like(phi); z = np.zeros_like(phi)
    m0 = (phi * W).sum() * G.h**2
    st = {'its': [], 'newton': []}
    t0 = time.time()
    for n in range(M):
        phi, mu, w = step(G, phi, mu, w, z, z, mode, st)
        me = (phi * W).sum() * G.h**2 - m0
        inter = np.abs(phi) < 0.985
        wi = (W * inter).sum() * G.h**2
        if wi > 0: phi = np.where(inter, phi - me / wi, phi)
        if (n + 1) in snaps: np.save(f'/tmp/vch_proto_phi_{N}_{n+1}.npy', np.stack([phi, mu, w]))
        if (n + 1) % 10 == 0:
            k = st['its'][-20:]
            print(n + 1, f"t={time.time()-t0:.0f}s newton/step={np.mean(st['newton'][-10:]):.2f} its/solve={np.mean(k):.2f} max|phi|={np.abs(phi).max():.3f}", flush=True)
    print('mean its/solve', np.mean(st['its']), 'mean newton/step', np.mean(st['newton']))
