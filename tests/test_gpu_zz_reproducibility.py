"""Bit-reproducibility of the 2D engine on the GPU (through the C ABI): same inputs -> same bits, from call to call and with several
contexts busy at the same time.  In a file of its own, collected last: it is the only test whose failure mode would be intermittent."""
import numpy as np
import pytest

import vch_oracle as O

pytestmark = pytest.mark.gpu


def make_ctx(nat, P: O.Phys2D):
    return nat.Ctx2D(P.Nx, P.Ny, P.Lx / P.Nx, P.Ly / P.Ny, P.Lx, P.Ly, P.tau, P.gamma, P.c1, P.c2, P.kappa)


@pytest.mark.parametrize("N,M,reps", [(128, 20, 3), (1024, 4, 3)])
def test_results_are_bit_reproducible(native, N, M, reps):
    """Same inputs -> same bits: from call to call on one context, and with other contexts busy on the same GPU at the same time.
    Regression for a race in the deferred BiCGStab update (vch_fft16.cuh, RowPrologue mode 3): the odd line out of the last line
    pair used to enter its complex FFT paired with whatever line 0 held while another CTA rewrote line 0 in place; its imaginary
    part leaks into the real line at rounding level, so repeated Jacobian solves at 1024^2 differed by ~1e-14 and trajectories by
    ~1e-11 after a few steps (N + 1 lines is odd on every grid; on a 128^2 grid it only showed with concurrent contexts)."""
    import threading
    import torch
    P = O.Phys2D(Nx=N, Ny=N, T=M * 1e-2)
    dts = np.full(M, 1e-2)
    t = 1e-2 * np.arange(M + 1)
    x = np.linspace(0, 1, N + 1)
    X, Y = np.meshgrid(x, x, indexing="ij")
    K = 3 if N <= 256 else 2

    class Problem:
        def __init__(self, seed):
            self.stream = torch.cuda.Stream()
            with torch.cuda.stream(self.stream):
                self.c = make_ctx(native, P)
                self.phi0 = torch.from_numpy(O.init_phi_2d(N, N, seed=seed)).cuda()
                self.phiT = torch.from_numpy(0.7 * np.sin(2 * np.pi * X) * np.cos(np.pi * Y)).cuda()
            self.stream.synchronize()
            self.out = None

        def run(self):
            with torch.cuda.stream(self.stream):
                c = self.c
                h = c.forward(self.phi0, None, dts)[0]
                s = torch.from_numpy(t / t[-1]).cuda()[:, None, None]
                Q = ((1 - s) * h[0] + s * self.phiT).contiguous()
                r = c.adjoint(h, t, 5.0, 10.0, Q, self.phiT, want_pq=False)[2]
                u1 = native.grad_prox(torch.zeros_like(r), r, 1e-3, 50.0, 1e-3, -1.0, 1.0)[0]
                h1 = c.forward(self.phi0, u1, dts)[0]
                self.stream.synchronize()
            self.out = (h, r, u1, h1, int(c.last_stats["krylov_iterations"]))

    probs = [Problem(42 + k) for k in range(K)]
    for p in probs:
        p.run()
    ref = [p.out for p in probs]
    for rep_ in range(reps):
        if rep_ % 2 == 0:       # all contexts at the same time, one host thread each
            th = [threading.Thread(target=p.run) for p in probs]
            [x_.start() for x_ in th]; [x_.join() for x_ in th]
        else:
            for p in probs:
                p.run()
        torch.cuda.synchronize()
        for k, p in enumerate(probs):
            for name, a, b in zip(("phi0_hist", "r", "u1", "phi1_hist"), ref[k][:4], p.out[:4]):
                assert torch.equal(a, b), f"repeat {rep_}, problem {k}: {name} differs by {float((a - b).abs().max()):.2e}"
            assert ref[k][4] == p.out[4]
    # the linear solve on its own (the race showed here first)
    c = probs[0].c
    phi = ref[0][0][M // 2].contiguous()
    g = torch.Generator(device="cuda").manual_seed(1)
    Rp, Rm = (1e-3 * torch.randn(phi.shape, dtype=torch.float64, device="cuda", generator=g) for _ in range(2))
    d0, m0, its0 = c.jacobian_solve(phi, 1e-2, Rp, Rm)
    for _ in range(6):
        d1, m1, its1 = c.jacobian_solve(phi, 1e-2, Rp, Rm)
        assert torch.equal(d0, d1) and torch.equal(m0, m1) and its0 == its1
