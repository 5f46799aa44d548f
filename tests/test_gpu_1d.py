"""GPU parity tests, 1D path (batched ensembles; batch = 1 is the reference call).  CUDA kernels behind the C ABI
against golden vectors from the unmodified reference and against the CPU oracle.

The 1D default problem is in the spinodal regime and amplifies a 1e-13 perturbation of phi_0 to ~2e-11 over T = 1
(measured on the reference itself), so trajectory agreement sits at 1e-10, inside the 1e-8 tolerance.
"""
import numpy as np
import pytest

import vch_oracle as O
from conftest import rel

pytestmark = pytest.mark.gpu
TOL_TRAJ, TOL_GRAD, TOL_J = 1e-8, 1e-7, 1e-7


def make_ctx(nat, P: O.Phys1D):
    return nat.Ctx1D(P.N, P.Lx / P.N, P.Lx, P.tau, P.gamma, P.c1, P.c2, P.kappa)


def dt_list(P):
    out, t = [], 0.0
    while t < P.T - 1e-10:
        d = min(P.dt_initial, P.T - t)
        out.append(d)
        t += d
    return np.array(out)


def test_residual_and_newton_match_oracle(native):
    P = O.Phys1D()
    c = make_ctx(native, P)
    L = O.neumann_1d(P.N, P.Lx / P.N).toarray()
    rng = np.random.default_rng(3)
    phi0 = O.init_phi_1d(P.N)
    w0 = np.zeros_like(phi0)
    mu0 = -P.kappa * (L @ phi0) + P.c1 * O.flory_log(phi0, O.log_eps()) - 2 * P.c2 * phi0
    w1 = 0.01 * rng.standard_normal(phi0.shape)
    phi = phi0 + 1e-3 * rng.standard_normal(phi0.shape)
    mu = mu0 + 1e-3 * rng.standard_normal(phi0.shape)
    Rp, Rm = c.residual(phi, phi0, mu, mu0, w1, w0, 1e-2)
    Rp_o, Rm_o = O.residual_1d(P, L, phi, mu, phi0, mu0, w1, w0, 1e-2)
    assert rel(Rp, Rp_o) < 1e-12 and rel(Rm, Rm_o) < 1e-12
    p_o, m_o, h_o = O.newton_1d(P, L, phi0, mu0, w0, w1, 1e-2)
    p, m, h = c.newton(phi0, mu0, w0, w1, 1e-2)
    assert len(h) == len(h_o) and 3 <= len(h) < 10 and h[-1] < 1e-6      # cf. test_1d_forward.py:342-395
    np.testing.assert_allclose(h[:-1], h_o[:-1], rtol=1e-6)
    assert rel(p, p_o) < 1e-11 and rel(m, m_o) < 1e-11


@pytest.mark.parametrize("name", ["g1d_default", "g1d_n64"])
def test_forward_adjoint_cost_prox_match_reference_golden(native, golden, name):
    g = golden(name)
    P = O.from_json(O.Phys1D, g["cfg_json"])
    Op = O.from_json(O.Opt1D, g["opt_json"])
    c = make_ctx(native, P)
    phi_init = O.init_phi_1d(P.N)
    dts = dt_list(P)
    hist, mu, w = c.forward(phi_init, None, dts, want_mu=True, want_w=True)
    assert hist.shape == g["phi0"].shape                                   # M+2 rows, level 0 twice
    assert np.array_equal(hist[0], hist[1])
    assert rel(hist, g["phi0"]) < TOL_TRAJ and rel(mu, g["mu0"]) < TOL_TRAJ
    # adjoint uses the reference's hard-wired default physics (backward_solver.py:29-33)
    cadj = make_ctx(native, O.Phys1D(N=P.N, Lx=P.Lx))
    p, q, r = cadj.adjoint(g["phi0"], g["t"], Op.b1, Op.b2, g["phiQ"], g["phiT"])
    assert np.abs(p[0]).max() == 0 and np.abs(r[0]).max() == 0             # dt = 0 level is skipped
    assert rel(p, g["p0"]) < TOL_GRAD and rel(q, g["q0"]) < TOL_GRAD and rel(r, g["r0"]) < TOL_GRAD
    u0 = np.zeros_like(g["phi0"])
    J0 = c.cost(g["phi0"], u0, g["phiQ"], g["phiT"], g["x"], g["t"], Op.b1, Op.b2, Op.b3, Op.kappa_sparsity)
    assert abs(J0[0] - g["J"][0]) <= TOL_J * abs(g["J"][0])
    u1, _, red = native.grad_prox(u0, g["r0"], Op.b3, Op.alpha_max, Op.kappa_sparsity, Op.u_min, Op.u_max)
    assert np.array_equal(u1, g["u1"])                                      # same inputs -> bit-identical prox
    # controlled forward (exercises the one-row control offset and the w filter)
    hist1, _, w1 = c.forward(phi_init, g["u1"], dts, want_w=True)
    assert rel(hist1, g["phi1"]) < TOL_TRAJ and rel(w1, g["w1"]) < TOL_TRAJ
    J1 = c.cost(g["phi1"], g["u1"], g["phiQ"], g["phiT"], g["x"], g["t"], Op.b1, Op.b2, Op.b3, Op.kappa_sparsity)
    assert abs(J1[0] - g["J"][1]) <= TOL_J * abs(g["J"][1])
    # mass conservation <= 1e-12 (test_1d_forward.py:185-223)
    wts = (P.Lx / P.N) * O.trapz_w(P.N + 1)
    m = hist1 @ wts
    assert np.abs(m - m[0]).max() < 1e-12


def test_ensemble_equals_single_problem_calls(native, golden):
    """A batch is B independent problems: every member equals the batch-1 call bit for bit."""
    g = golden("g1d_default")
    P = O.from_json(O.Phys1D, g["cfg_json"])
    c = make_ctx(native, P)
    rng = np.random.default_rng(1234)
    B = 6
    dts = dt_list(P)
    phi_init = np.stack([O.init_phi_1d(P.N, seed=40 + b) for b in range(B)])
    u = 0.2 * rng.standard_normal((B, len(dts) + 2, P.N + 1))
    hist, _, _ = c.forward(phi_init, u, dts)
    for b in (0, 3, 5):
        hb, _, _ = c.forward(phi_init[b], u[b], dts)
        assert np.array_equal(hb, hist[b])
    t = np.concatenate([[0.0], np.concatenate([[0.0], np.cumsum(dts)])])
    b1 = rng.uniform(0.1, 1, B); b2 = rng.uniform(5, 20, B)
    phiT = np.stack([O.targets_1d(g["x"], t, hist[b, 0], P.Lx, A_T=rng.uniform(0.3, 0.8), choice_t=1 + b % 3)[0] for b in range(B)])
    phiQ = np.stack([(1 - (t / t[-1])[:, None]) * hist[b, 0] + (t / t[-1])[:, None] * phiT[b] for b in range(B)])
    p, q, r = c.adjoint(hist, t, b1, b2, phiQ, phiT)
    po, qo, ro = O.adjoint_1d(hist[2], g["x"], t, b1[2], b2[2], phiQ[2], phiT[2], P)
    assert rel(r[2], ro) < TOL_GRAD and rel(p[2], po) < TOL_GRAD
    b3 = np.exp(rng.uniform(np.log(1e-4), np.log(1e-2), B)); ks = np.exp(rng.uniform(np.log(1e-5), np.log(1e-3), B))
    J = c.cost(hist, u, phiQ, phiT, g["x"], t, b1, b2, b3, ks)
    Jo, parts = O.cost_1d(hist[4], u[4], phiQ[4], phiT[4], g["x"], t, b1[4], b2[4], b3[4], ks[4])
    assert abs(J[4, 0] - Jo) <= TOL_J * abs(Jo) and rel(J[4, 1:], parts) < 1e-9
    un, red = c.grad_prox(u, r, b3, 100.0, ks, -1.0, 1.0)
    for b in (1, 4):
        assert np.array_equal(un[b], O.soft_prox(u[b], r[b] + b3[b] * u[b], 100.0, ks[b], -1.0, 1.0))
        assert red[b, 2] == np.count_nonzero(un[b])


def test_temporal_order_and_symmetry(native):
    """Reference properties: symmetric IC stays symmetric (test_1d_forward.py:300-319); CN is ~2nd order (:253-296)."""
    P = O.Phys1D(N=128, T=0.02, dt_initial=1e-3)
    c = make_ctx(native, P)
    x = np.linspace(0, 1, 129)
    sym = 0.05 * np.cos(2 * np.pi * x)
    h, _, _ = c.forward(sym, None, dt_list(P))
    assert np.abs(h[-1] - h[-1][::-1]).max() < 1e-8
    errs = []
    ref = None
    for dt in (1e-3 / 8, 1e-3, 5e-4, 2.5e-4):
        Pd = O.Phys1D(N=128, T=0.02, dt_initial=dt)
        hh, _, _ = make_ctx(native, Pd).forward(sym, None, dt_list(Pd))
        if ref is None:
            ref = hh[-1]
        else:
            errs.append(np.abs(hh[-1] - ref).max())
    slope = np.polyfit(np.log([1e-3, 5e-4, 2.5e-4]), np.log(errs), 1)[0]
    assert 1.2 < slope < 2.3


def test_ensemble_members_match_reference_golden(native, golden):
    """BASELINE config 4: 16 randomly chosen members of the 1024-problem ensemble (varied targets, kappa_sp, b1..b3), each taken
    through one optimistic PGD iteration by the UNMODIFIED 1D reference (oracle/make_golden_ensemble.py), against ONE batched
    call per stage of the library — the way the ensemble driver runs it (GD_1D.optimistic_iteration_ensemble).  The members
    are cut out of the product's own make_ensemble(1024), so the ensemble definition itself is pinned as well."""
    import sys, os
    from conftest import load_dropin
    G = load_dropin("1D")["GD_1D"]
    g = golden("g1d_ensemble16")
    ens = G.make_ensemble(1024)
    mem = g["members"]
    for k in ("b1", "b2", "b3", "ksp"):
        np.testing.assert_allclose(ens[k][mem], g[k], rtol=1e-15)
    assert np.array_equal(ens["choice_t"][mem], g["choice_t"])
    P = O.Phys1D()
    c = make_ctx(native, P)
    sel = lambda a: np.ascontiguousarray(a[mem])
    phi_init, phiQ, phiT = sel(ens["phi_init"]), sel(ens["phi_Q"]), sel(ens["phi_T"])
    hist0, _, _ = c.forward(phi_init, None, ens["dts"])
    assert rel(hist0[0], g["phi0"]) < TOL_TRAJ                                # u = 0: every member has the same trajectory
    u0 = np.zeros_like(hist0)
    u1, hist1, J, red, r = G.optimistic_iteration_ensemble(c, c, u0, hist0, phiQ, phiT, ens["x"], ens["t_hist"], ens["dts"],
                                                           phi_init, g["b1"], g["b2"], g["b3"], g["ksp"], float(g["alpha"]))
    for k in range(len(mem)):
        assert rel(r[k], g["r0"][k]) < TOL_GRAD, (k, rel(r[k], g["r0"][k]))
        assert rel(u1[k], g["u1"][k]) < TOL_GRAD
        assert np.array_equal(u1[k] != 0, g["u1"][k] != 0), f"member {mem[k]}: control support differs"
        assert rel(hist1[k], g["phi1"][k]) < TOL_TRAJ, (k, rel(hist1[k], g["phi1"][k]))
        assert abs(J[k, 0] - g["J1"][k]) <= TOL_J * abs(g["J1"][k])
