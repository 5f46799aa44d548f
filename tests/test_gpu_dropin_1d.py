"""Drop-in acceptance tests, 1D: `Vch_control_1D` modules by bare name, properties of src/1D/tests_1D/* re-stated,
plus golden checks and the ensemble driver."""
import contextlib
import io

import numpy as np
import pytest

import vch_oracle as O
from conftest import load_dropin, rel

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def m(native):
    return load_dropin("1D")


def quiet(f, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return f(*a, **k)


def test_forward_contract_mass_energy(m, golden):
    F, C = m["Forward_solver"], m["config"]
    g = golden("g1d_default")
    phi, x, t = quiet(F.run_main_simulation, C.ForwardSolverConfig(), store_history=True, verbose=True)
    assert phi.shape == (102, 129) and np.array_equal(t[:3], [0.0, 0.0, 0.01]) and rel(phi, g["phi0"]) < 1e-8
    fin, x2, t2 = quiet(F.run_main_simulation, C.ForwardSolverConfig(), store_history=False, verbose=False)
    assert np.array_equal(fin, phi[-1]) and len(t2) == 102
    assert np.array_equal(quiet(F.run_main_simulation, None, store_history=True, verbose=False)[0], phi)   # fwd_config=None
    wts = (1 / 128) * F.trapz_weights(129)
    assert np.abs(phi @ wts - phi[0] @ wts).max() < 1e-12                                  # test_1d_forward.py:185-223
    E = [F.free_energy(p, 0.03 ** 2, 0.75, 1.0, 1 / 128) for p in phi[1:]]
    assert np.all(np.diff(E) <= 1e-9)                                                      # :225-251
    sym = 0.05 * np.cos(2 * np.pi * x)
    ps, _, _ = quiet(F.run_main_simulation, C.ForwardSolverConfig(T=0.1), store_history=True, verbose=False, initial_phi=sym)
    assert np.abs(ps[-1] - ps[-1][::-1]).max() < 1e-8                                      # :300-319
    big, _, _ = quiet(F.run_main_simulation, C.ForwardSolverConfig(dt_initial=1.0, T=3.0), store_history=True, verbose=False)
    assert np.isfinite(big).all()                                                          # :323-339


def test_building_blocks(m):
    F = m["Forward_solver"]
    N, h = 128, 1 / 128
    L = F.laplacian_matrix_neumann(N, h)
    assert isinstance(L, np.ndarray) and np.array_equal(L, O.neumann_1d(N, h).toarray())
    rng = np.random.default_rng(11)
    a, b, c = rng.standard_normal((3, 129))
    g = 10.0 / 1e-2
    np.testing.assert_allclose(F.solve_w(a, 1e-2, 10.0, b, c), ((g - 0.5) * a + 0.5 * (c + b)) / (g + 0.5), rtol=1e-15)
    phi = 0.5 * np.tanh(a)
    assert rel(F.initialize_mu(phi, b, 0.75, 1.0, L, 9e-4), -9e-4 * (L @ phi) + 0.75 * F.regularized_log(phi) - 2 * phi - b) < 1e-12
    phi0 = F.init_phi_random(N, 1e-2, amp=0.01, seed=42)
    mu0 = F.initialize_mu(phi0, 0 * phi0, 0.75, 1.0, L, 9e-4)
    p, mu, hist = F.newton_raphson(phi0, mu0, 0 * phi0, 0 * phi0, 1e-2, 0.05, 0.75, 1.0, h, 1e-2, L, 9e-4, return_residual_history=True)
    assert 3 <= len(hist) < 10 and hist[-1] < 1e-6 and all(y < x for x, y in zip(hist[1:], hist[2:]))   # :342-395
    Rp = F.solve_phi_residual(p, phi0, mu, mu0, 0 * p, 0 * p, 1e-2, 0.05, 0.75, 1.0, L, 9e-4)
    Rm = F.solve_mu_residual(p, phi0, mu, mu0, 1e-2, L)
    assert abs(np.sqrt((Rp ** 2).sum() + (Rm ** 2).sum()) - hist[-1]) < 1e-10


def test_backward_cost_prox(m, golden):
    B, Cst, G = m["backward_solver"], m["cost_and_function"], m["GD_1D"]
    g = golden("g1d_default")
    opt = m["config"].OptimizationConfig()
    p, q, r = B.run_backward(g["phi0"], g["x"], g["t"], opt.b1, opt.b2, g["phiQ"], g["phiT"])
    assert rel(r, g["r0"]) < 1e-7 and rel(p, g["p0"]) < 1e-7 and np.abs(r[0]).max() == 0 and np.abs(r[-1]).max() == 0
    # adjoint step identity with independently assembled A, B on a synthetic history (test_1d_backward.py:199-229)
    x = np.linspace(0, 1, 129); t = np.linspace(0, 0.2, 6)
    hist = np.array([0.2 * np.sin(np.pi * x) * (1 + 0.2 * np.cos(2 * np.pi * tn / 0.2)) for tn in t])
    ps, qs, rs = B.run_backward(hist, x, t, 0.3, 13.0)
    L = O.neumann_1d(128, x[1] - x[0]).toarray(); I = np.eye(129)
    for n in range(4, -1, -1):
        dt = t[n + 1] - t[n]
        A = I - 0.05 * L + 0.5 * dt * L @ L - 0.5 * dt * np.diag(B.fpp_log(hist[n])) @ L
        Bm = I - 0.05 * L - 0.5 * dt * L @ L + 0.5 * dt * np.diag(B.fpp_log(hist[n + 1])) @ L
        rhs = Bm @ ps[n + 1] + 0.5 * dt * 0.3 * (hist[n] + hist[n + 1])
        assert np.linalg.norm(A @ ps[n] - rhs) / np.linalg.norm(rhs) < 2e3 * np.finfo(float).eps * np.linalg.cond(A)
    u0 = np.zeros_like(g["phi0"])
    out = io.StringIO()
    with contextlib.redirect_stdout(out):
        J0 = Cst.calculate_cost(g["phi0"], u0, g["phiQ"], g["phiT"], g["x"], g["t"], opt.b1, opt.b2, opt.b3, opt.kappa_sparsity, verbose=False)
    assert "Total Cost" in out.getvalue() and abs(J0 - g["J"][0]) < 1e-7 * g["J"][0]
    assert quiet(Cst.calculate_cost, u0, u0, u0, u0[0], g["x"], g["t"], 1, 1, 1, 1) == 0.0       # test_1d_cost.py:140-162
    grad = Cst.calculate_gradient(g["r0"], u0, opt.b3)
    u1 = G.perform_proximal_and_projection(Cst.perform_gradient_step(u0, grad, opt.alpha_max), opt.alpha_max,
                                           opt.kappa_sparsity, opt.u_min, opt.u_max)
    assert np.array_equal(u1, g["u1"])
    rng = np.random.default_rng(3)
    v = rng.standard_normal((5, 33))
    np.testing.assert_allclose(G.perform_proximal_and_projection(v, 0.5, 0.4, -0.6, 0.9),
                               np.clip(np.sign(v) * np.maximum(np.abs(v) - 0.2, 0), -0.6, 0.9), atol=1e-12)


def test_driver_and_ensemble(m, golden):
    G, C = m["GD_1D"], m["config"]
    g = golden("g1d_default")
    res = quiet(G.optimize, C.ForwardSolverConfig(), C.OptimizationConfig(), 1, 1, max_iter=1, verbose=False)
    np.testing.assert_allclose(res["cost_history"], g["J"][:2], rtol=1e-7)
    assert abs(res["cost_history"][0] - 1.0977825116186202) < 1e-7 and abs(res["cost_history"][1] - 0.3218003723414755) < 1e-7
    # ensemble (BASELINE config 4, reduced to 24 members) == member-by-member single-problem calls
    import vch_b200_native as nat
    ens = G.make_ensemble(24)
    cfg = C.ForwardSolverConfig()
    c = nat.ctx1d(cfg.N, cfg.Lx / cfg.N, cfg.Lx, cfg.tau, cfg.gamma, cfg.c1, cfg.c2, cfg.kappa)
    hist0, _, _ = c.forward(ens["phi_init"], None, ens["dts"])
    u0 = np.zeros_like(hist0)
    u1, hist1, J, red, r = G.optimistic_iteration_ensemble(c, c, u0, hist0, ens["phi_Q"], ens["phi_T"], ens["x"], ens["t_hist"],
                                                           ens["dts"], ens["phi_init"], ens["b1"], ens["b2"], ens["b3"], ens["ksp"], 100.0)
    for b in (0, 7, 23):
        P = O.Phys1D()
        Op = O.Opt1D(b1=ens["b1"][b], b2=ens["b2"][b], b3=ens["b3"][b], kappa_sparsity=ens["ksp"][b])
        un, fw, Jo, ro = O.pgd_iter_1d(P, Op, u0[b], hist0[b], ens["t_hist"], ens["x"], ens["phi_Q"][b], ens["phi_T"][b], 100.0)
        assert rel(r[b], ro) < 1e-7 and rel(u1[b], un) < 1e-7 and np.array_equal(u1[b] != 0, un != 0)
        assert rel(hist1[b], fw["phi"]) < 1e-8 and abs(J[b, 0] - Jo) < 1e-7 * abs(Jo)
    assert G.shard_range(1024, 3, 8) == (384, 512) and G.shard_range(10, 0, 4) == (0, 3) and G.shard_range(10, 3, 4) == (8, 10)


def test_batched_line_search_and_fd_directions_equal_sequential(m, golden):
    """line-search trials and finite-difference directions as ONE multi-problem launch (batch axis of the 1D
    kernels) return exactly what the reference's sequential loops return (GD_1D.py:73-113, second_order_conditions.py:58-110)."""
    G, C, S, F, Cst, B = (m[k] for k in ("GD_1D", "config", "second_order_conditions", "Forward_solver", "cost_and_function", "backward_solver"))
    cfg, opt = C.ForwardSolverConfig(), C.OptimizationConfig()
    phi, x, t = quiet(F.run_main_simulation, cfg, True, None, False)
    phi_T, phi_Q = quiet(G.build_targets_1d, x, t, phi[0].copy(), cfg.Lx, cfg.T, False, 1, 1)
    u = np.zeros_like(phi)
    J0 = quiet(Cst.calculate_cost, phi, u, phi_Q, phi_T, x, t, opt.b1, opt.b2, opt.b3, opt.kappa_sparsity)
    _, _, r = quiet(B.run_backward, phi, x, t, opt.b1, opt.b2, phi_Q, phi_T)
    g = Cst.calculate_gradient(r, u, opt.b3)
    # costs of the six trial step sizes, then a reference cost that the first trials do NOT beat: the search has to shrink alpha
    alphas = [4.0e4 * 0.5 ** j for j in range(6)]
    us = np.stack([G.perform_proximal_and_projection(Cst.perform_gradient_step(u, g, a), a, opt.kappa_sparsity, opt.u_min, opt.u_max)
                   for a in alphas])
    Js = Cst.calculate_cost_batch(F.run_main_simulation_batch(cfg, us)[0], us, phi_Q, phi_T, x, t, opt.b1, opt.b2, opt.b3,
                                  opt.kappa_sparsity)[:, 0]
    first = next((j for j in range(1, 6) if Js[j] < Js[:j].min()), None)
    cost_k = 0.5 * (Js[first] + Js[:first].min()) if first is not None else Js.min() - 1.0     # (else: every trial fails)
    args = (u, cost_k, g, phi_Q, phi_T, x, t, opt.b1, opt.b2, opt.b3, opt.kappa_sparsity, opt.u_min, opt.u_max, cfg)
    seq = quiet(G.perform_backtracking_line_search, *args, alpha_init=4.0e4, beta=0.5, max_ls_iter=6, batch=1)
    bat = quiet(G.perform_backtracking_line_search, *args, alpha_init=4.0e4, beta=0.5, max_ls_iter=6, batch=4)
    assert seq[6] == bat[6] and seq[6] >= 2, (seq[6], bat[6])                 # same number of trials, more than one
    assert seq[0] == bat[0] and seq[2] == bat[2]                                # alpha and cost bit for bit
    assert np.array_equal(seq[1], bat[1]) and np.array_equal(seq[3], bat[3])    # control and trajectory bit for bit
    kw = dict(num_directions=4, epsilon=1e-4, seed=5)
    sargs = (cfg, seq[1], r, seq[3], x, t, opt.b1, opt.b2, opt.b3, opt.kappa_sparsity, phi_Q, phi_T, opt.u_min, opt.u_max)
    d_seq = quiet(S.approximate_second_order_condition, *sargs, batch=1, **kw)
    d_bat = quiet(S.approximate_second_order_condition, *sargs, **kw)           # default: all directions in one launch
    assert d_seq == d_bat and len(d_bat) == 4
