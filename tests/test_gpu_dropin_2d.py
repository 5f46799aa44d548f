"""Drop-in acceptance tests, 2D: the package's `Vch_control_2D` modules imported by bare name, exercised the way the
reference's own suites do (src/2D/tests_2D/*; the properties are re-stated here, not copied) plus golden checks."""
import contextlib
import io

import numpy as np
import pytest

import vch_oracle as O
from conftest import load_dropin, rel

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def m(native):
    return load_dropin("2D")


def quiet(f, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return f(*a, **k)


def test_laplacian_operator_and_eigenmode(m):
    F = m["Forward2_solver"]
    Nx = Ny = 64
    h = 1.0 / Nx
    L = F.laplacian_matrix_neumann(Nx, Ny, h, h)
    assert hasattr(L, "toarray") and (L @ L).shape == L.shape                      # SciPy sparse, as the tests require
    assert abs(L - O.neumann_2d(Nx, Ny, h, h)).max() == 0
    x = np.linspace(0, 1, Nx + 1)
    X, Y = np.meshgrid(x, x, indexing="ij")
    v = np.cos(2 * np.pi * X) * np.cos(np.pi * Y)
    lam = -((2 * np.pi) ** 2 + np.pi ** 2)
    assert rel(F.apply_laplacian(L, v, Nx, Ny), lam * v) < 5e-3                     # cf. test_2d_forward.py:155-173
    assert np.abs(F.apply_laplacian(L, np.ones_like(v), Nx, Ny)).max() < 1e-10
    with pytest.raises(ValueError):
        F.apply_laplacian(L, np.zeros((3, 3)), Nx, Ny)
    # a matrix that lost the cached spacing (e.g. after arithmetic) still works: spacing is read from its entries
    assert rel(F.apply_laplacian((1.0 * L).tocsr(), v, Nx, Ny), L @ v.ravel()) < 1e-13


def test_ic_solve_w_initialize_mu(m):
    F = m["Forward2_solver"]
    phi = F.init_phi_random(128, 128, 1e-2, amp=0.1, seed=42)
    assert np.array_equal(phi, O.init_phi_2d(128, 128))
    w = np.outer(F.trapz_weights(129), F.trapz_weights(129))
    assert abs((w * phi).sum() / w.sum()) < 5e-14 and np.abs(phi).max() <= 0.99
    rng = np.random.default_rng(0)
    a, b, c = rng.standard_normal((3, 17, 9))
    g = 10.0 / 1e-2
    np.testing.assert_allclose(F.solve_w(a, 1e-2, 10.0, b, c), ((g - 0.5) * a + 0.5 * (c + b)) / (g + 0.5), rtol=1e-15)
    L = F.laplacian_matrix_neumann(128, 128, 1 / 128, 1 / 128)
    mu = F.initialize_mu(phi, 0 * phi, 0.75, 1.0, 1e-4, L, 128, 128, 1e-2)
    ref = -1e-4 * (L @ phi.ravel()).reshape(phi.shape) + 0.75 * F.regularized_log(phi, 1e-2) - 2.0 * phi
    assert rel(mu, ref) < 1e-12                                                     # cf. test_2d_Cost.py:137-163


def test_run_main_simulation_contract_and_physics(m, golden):
    F, C = m["Forward2_solver"], m["config"]
    cfg = C.ForwardSolverConfig(Nx=32, Ny=32, T=0.2, dt_initial=1e-2)
    out = quiet(F.run_main_simulation, cfg, store_history=True, control_input=None, verbose=True)
    phi, (x, y), t = out
    g = golden("g2d_32")
    assert phi.shape == (21, 33, 33) and np.allclose(t, g["t"]) and rel(phi, g["phi0"]) < 1e-8
    with pytest.raises(ValueError):
        F.run_main_simulation(cfg, store_history=True, control_input=np.zeros((21, 5, 5)), verbose=False)
    assert quiet(F.run_main_simulation, C.ForwardSolverConfig(Nx=16, Ny=16, T=0.02), store_history=False, verbose=False) is None
    # mass conservation and energy decay (test_2d_forward.py:213-279)
    wts = (1 / 32) ** 2 * np.outer(F.trapz_weights(33), F.trapz_weights(33))
    mass = (wts * phi).sum(axis=(1, 2))
    assert np.abs(mass - mass[0]).max() < 1e-11
    E = [F.free_energy(p, cfg.kappa, cfg.c1, cfg.c2, 1 / 32, 1 / 32) for p in phi]
    assert np.all(np.diff(E) <= 1e-9)
    # symmetric IC via monkey-patching the module attribute, looked up at call time (test_2d_forward.py:282-299)
    X, Y = np.meshgrid(x, y, indexing="ij")
    orig = F.init_phi_random
    F.init_phi_random = lambda *a, **k: 0.05 * np.cos(2 * np.pi * X) * np.cos(2 * np.pi * Y)
    try:
        ps, _, _ = quiet(F.run_main_simulation, cfg, store_history=True, verbose=False)
    finally:
        F.init_phi_random = orig
    assert np.abs(ps[-1] - ps[-1][::-1, :]).max() < 1e-8 and np.abs(ps[-1] - ps[-1][:, ::-1]).max() < 1e-8


def test_newton_raphson_history_is_quadratic(m):
    F = m["Forward2_solver"]
    N, h, dt = 64, 1 / 64, 1e-2
    L = F.laplacian_matrix_neumann(N, N, h, h)
    phi0 = F.init_phi_random(N, N, 1e-2, amp=0.1, seed=42)
    w0 = np.zeros_like(phi0)
    mu0 = F.initialize_mu(phi0, w0, 0.75, 1.0, 1e-4, L, N, N, 1e-2)
    p, mu, hist = F.newton_raphson(phi0, mu0, w0, w0, dt, 0.05, 0.75, 1.0, 1e-4, 1e-2, L, N, N, h, h, return_residual_history=True)
    assert len(F.newton_raphson(phi0, mu0, w0, w0, dt, 0.05, 0.75, 1.0, 1e-4, 1e-2, L, N, N, h, h)) == 2
    assert hist[-1] < 1e-6 and all(b < a for a, b in zip(hist[1:], hist[2:]))
    e = np.array(hist[1:])
    if len(e) >= 3:                                                                 # log-log slope ~2 (test_2d_forward.py:404-491)
        slope = np.polyfit(np.log(e[:-1]), np.log(e[1:]), 1)[0]
        assert 1.5 < slope < 2.6
    Rp = F.solve_phi_residual(p, phi0, mu, mu0, w0, w0, dt, 0.05, 0.75, 1.0, 1e-4, L, N, N, 1e-2)
    Rm = F.solve_mu_residual(p, phi0, mu, mu0, dt, L, N, N)
    assert abs(np.sqrt((Rp ** 2).sum() + (Rm ** 2).sum()) - hist[-1]) < 1e-9
    J = F.assemble_jacobian(p, dt, 0.05, 0.75, 1e-4, L, 1e-2)
    assert J.shape == (2 * p.size, 2 * p.size)


def test_run_backward_cost_prox_and_kkt(m, golden):
    B, Cst, C, S = m["backward2_solver"], m["cost2_and_function"], m["config"], m["second_order_conditions_2d"]
    g = golden("g2d_32")
    cfg, opt = C.ForwardSolverConfig(Nx=32, Ny=32, T=0.2), C.OptimizationConfig()
    phiT, phiQ = O.targets_2d(g["x"], g["y"], g["t"], g["phi0"][0], 1.0, 1.0, 0.2)
    p, q, r = B.run_backward(g["phi0"], g["x"], g["y"], g["t"], cfg, opt.b1, opt.b2, phiQ, phiT)
    assert rel(r, g["r0"]) < 1e-7 and rel(p, g["p0"]) < 1e-7
    with pytest.raises(AssertionError):
        B.run_backward(g["phi0"][0], g["x"], g["y"], g["t"], cfg, 1.0, 1.0)
    u0 = np.zeros_like(g["phi0"])
    out = io.StringIO()
    with contextlib.redirect_stdout(out):
        J0 = Cst.calculate_cost(g["phi0"], u0, phiQ, phiT, g["x"], g["y"], g["t"], opt)
    assert "Tracking Cost (J1)" in out.getvalue() and abs(J0 - g["J"][0]) < 1e-7 * g["J"][0]
    # isolated terms against analytic values (test_2d_Cost.py:208-300): phi - phi_Q = c, phi(T) - phi_T = d, u = e
    z = np.zeros_like(g["phi0"])
    o1 = C.OptimizationConfig(b1=2.0, b2=0.0, b3=0.0, kappa_sparsity=0.0)
    assert abs(quiet(Cst.calculate_cost, z + 0.3, z, z, 0 * phiT + 0.3, g["x"], g["y"], g["t"], o1) - 0.5 * 2.0 * 0.09 * 0.2) < 1e-12
    o3 = C.OptimizationConfig(b1=0.0, b2=0.0, b3=4.0, kappa_sparsity=0.5)
    assert abs(quiet(Cst.calculate_cost, z, z - 0.2, z, 0 * phiT, g["x"], g["y"], g["t"], o3) - (0.5 * 4 * 0.04 * 0.2 + 0.5 * 0.2 * 0.2)) < 1e-12
    grad = Cst.calculate_gradient(r, u0 + 0.1, opt)
    assert np.array_equal(grad, r + opt.b3 * (u0 + 0.1))
    u1 = Cst.proximal_step(u0, Cst.calculate_gradient(r, u0, opt), opt.alpha_max, opt)
    assert np.array_equal(u1 != 0, g["u1"] != 0) and rel(u1, g["u1"]) < 1e-7
    # ISTA properties (test_2d_proximal.py:133-257): soft threshold, box, fixed point
    rng = np.random.default_rng(7)
    v, gr = rng.standard_normal((4, 9, 9)), rng.standard_normal((4, 9, 9))
    ob = C.OptimizationConfig(kappa_sparsity=0.3, u_min=-0.5, u_max=0.7)
    y = v - 0.2 * gr
    np.testing.assert_allclose(Cst.proximal_step(v, gr, 0.2, ob), np.clip(np.sign(y) * np.maximum(np.abs(y) - 0.06, 0), -0.5, 0.7), atol=1e-12)
    assert np.array_equal(Cst.proximal_step(0 * v, 0.29 * np.sign(gr), 1.0, ob), 0 * v)
    out = io.StringIO()
    with contextlib.redirect_stdout(out):
        S.verify_sparsity_condition(g["u2"], g["r1"], opt.kappa_sparsity)
    a, b, mm = O.kkt_counts(g["u2"], g["r1"], opt.kappa_sparsity)
    assert f"({a}/{g['u2'].size} points)" in out.getvalue()


def test_driver_optimize_matches_reference_iterations(m, golden):
    """Two PGD iterations of the drop-in driver reproduce the reference's cost sequence (device-resident and host paths)."""
    G, C = m["GD2_configured"], m["config"]
    g = golden("g2d_32")
    cfg, opt = C.ForwardSolverConfig(Nx=32, Ny=32, T=0.2), C.OptimizationConfig()
    for resident in (True, False):
        res = quiet(G.optimize, cfg, opt, 1, 1, max_iter=2, device_resident=resident, verbose=False)
        np.testing.assert_allclose(res["cost_history"], g["J"], rtol=1e-7)
        assert rel(res["u"], g["u2"]) < 1e-7 and rel(res["phi_hist"], g["phi2"]) < 1e-8
    phiT, phiQ = quiet(G.build_targets, g["x"], g["y"], g["t"], g["phi0"][0], 1.0, 1.0, 0.2, False, 1, 1)
    assert np.array_equal(phiT, g["phiT"])


def test_default_config_first_iteration_matches_reference(m, golden):
    """The reference's default 2D run (128^2, 100 steps): J0, J1 and sub-sampled phi / r / u1 (BASELINE.md anchors)."""
    G, C = m["GD2_configured"], m["config"]
    g = golden("g2d_128")
    res = quiet(G.optimize, C.ForwardSolverConfig(), C.OptimizationConfig(), 1, 1, max_iter=1, verbose=False)
    np.testing.assert_allclose(res["cost_history"], g["J"], rtol=1e-7)
    assert abs(res["cost_history"][0] - 3.0192834439912475) < 1e-7 and abs(res["cost_history"][1] - 1.5416947343810048) < 1e-7
    keep = g["keep"]
    assert rel(res["phi_hist"][keep], g["phi1"]) < 1e-8
    assert rel(res["u"][keep], g["u1"]) < 1e-7 and np.array_equal(res["u"][keep] != 0, g["u1"] != 0)


def test_driver_loop_matches_reference_loop(m, golden):
    """K iterations of the reference's own PGD loop (oracle/make_golden_loop.py, default 128^2 config): cost, step-size and
    control-change histories, line-search events included, reproduced by the device-resident driver."""
    import os
    from conftest import ROOT
    if not os.path.exists(os.path.join(ROOT, "tests", "golden", "g2d_128_loop.npz")):
        pytest.skip("loop golden not generated")
    g = golden("g2d_128_loop")
    G, C = m["GD2_configured"], m["config"]
    K = len(g["J"]) - 1
    res = quiet(G.optimize, C.ForwardSolverConfig(), C.OptimizationConfig(), 1, 1, max_iter=K, device_resident=True, verbose=False)
    np.testing.assert_allclose(res["cost_history"], g["J"], rtol=1e-7)
    np.testing.assert_allclose(res["alpha_history"], g["alpha"], rtol=1e-12)
    assert res["timers"]["ls_attempts"] == int(g["ls_attempts"].sum())
    assert rel(res["u"][[1, 50, 99]][:, ::4, ::4], g["u_final_sub"]) < 1e-6
    assert rel(res["phi_hist"][-1], g["phi_final_T"]) < 1e-7


def test_free_energy_and_second_order_condition_match_reference(native, golden):
    """`free_energy` is a fused device reduction (Forward2_solver.py:256-319) — against the NumPy formula of the
    reference; `approximate_second_order_condition_2d` (second_order_conditions_2d.py:120-235) — against the values the
    UNMODIFIED reference produced for the same iterate, directions and epsilon (oracle/make_golden.py g2d_32_soc)."""
    mods = load_dropin("2D")
    F, S, G = mods["Forward2_solver"], mods["second_order_conditions_2d"], mods["GD2_configured"]
    rng = np.random.default_rng(3)
    phi = 0.9 * np.tanh(rng.standard_normal((21, 13))); w = 0.1 * rng.standard_normal((21, 13))
    hx, hy, kappa, c1, c2 = 1 / 12, 1.5 / 20, 1e-4, 0.75, 1.0
    def ref(phi, w, eps=1e-8):
        tw = lambda n: np.r_[0.5, np.ones(n - 2), 0.5]
        wts = np.outer(tw(phi.shape[0]), tw(phi.shape[1]))
        E = kappa / (2 * hx) * np.sum(np.diff(phi, axis=1) ** 2) * hy + kappa / (2 * hy) * np.sum(np.diff(phi, axis=0) ** 2) * hx
        s = np.clip(phi, -1 + eps, 1 - eps)
        E += hx * hy * np.sum(wts * (c1 * ((1 + s) * np.log(1 + s) + (1 - s) * np.log(1 - s)) - c2 * s ** 2))
        return E - (hx * hy * np.sum(wts * w * phi) if w is not None else 0.0)
    assert abs(F.free_energy(phi, kappa, c1, c2, hx, hy) - ref(phi, None)) <= 1e-12 * abs(ref(phi, None))
    assert abs(F.free_energy(phi, kappa, c1, c2, hx, hy, w=w, eps=1e-3) - ref(phi, w, 1e-3)) <= 1e-12 * abs(ref(phi, w, 1e-3))
    g, gs = golden("g2d_32"), golden("g2d_32_soc")
    import json
    cfg = mods["config"].ForwardSolverConfig(**{k: v for k, v in json.loads(str(g["cfg_json"])).items()
                                                if k in mods["config"].ForwardSolverConfig.model_fields})
    opt = mods["config"].OptimizationConfig()
    phiT, phiQ = G.build_targets(g["x"], g["y"], g["t"], g["phi0"][0].copy(), cfg.Lx, cfg.Ly, cfg.T, False, 1, 1)
    d2 = quiet(S.approximate_second_order_condition_2d, g["u1"], g["r1"], g["phi1"], g["x"], g["y"], g["t"], opt_config=opt,
               phi_Q_target=phiQ, phi_T_target=phiT, u_min=opt.u_min, u_max=opt.u_max,
               num_directions=3, epsilon=float(gs["epsilon"]), seed=int(gs["seed"]), fwd_config=cfg)
    np.testing.assert_allclose(d2, gs["d2"], rtol=1e-4)


def test_concurrent_line_search_and_fd_directions_equal_sequential(native, golden):
    """trial step sizes / finite-difference directions evaluated concurrently (one worker thread and one
    library context per trial: independent streams and work vectors) return exactly what the sequential loops of the
    reference return (GD2_configured.py:71-146, second_order_conditions_2d.py:120-235)."""
    import json
    mods = load_dropin("2D")
    G, S, Cst, B = mods["GD2_configured"], mods["second_order_conditions_2d"], mods["cost2_and_function"], mods["backward2_solver"]
    g = golden("g2d_32")
    cfg = mods["config"].ForwardSolverConfig(**{k: v for k, v in json.loads(str(g["cfg_json"])).items()
                                                if k in mods["config"].ForwardSolverConfig.model_fields})
    opt = mods["config"].OptimizationConfig()
    phiT, phiQ = quiet(G.build_targets, g["x"], g["y"], g["t"], g["phi0"][0].copy(), cfg.Lx, cfg.Ly, cfg.T, False, 1, 1)
    u0 = np.zeros_like(g["phi0"])
    J0 = quiet(Cst.calculate_cost, g["phi0"], u0, phiQ, phiT, g["x"], g["y"], g["t"], opt)
    grad = Cst.calculate_gradient(g["r0"], u0, opt)
    args = (u0, J0, grad, phiQ, phiT, g["x"], g["y"], cfg, opt)
    seq = quiet(G.perform_backtracking_line_search_2D, *args, alpha_init=3.0e5, beta=0.25, max_ls_iter=6, batch=1)
    con = quiet(G.perform_backtracking_line_search_2D, *args, alpha_init=3.0e5, beta=0.25, max_ls_iter=6, batch=3)
    assert seq[6] == con[6], (seq[6], con[6])
    assert seq[0] == con[0] and seq[2] == con[2] and np.array_equal(seq[1], con[1]) and np.array_equal(seq[3], con[3])
    kw = dict(opt_config=opt, phi_Q_target=phiQ, phi_T_target=phiT, u_min=opt.u_min, u_max=opt.u_max, num_directions=4,
              epsilon=1e-4, seed=11, fwd_config=cfg)
    d_seq = quiet(S.approximate_second_order_condition_2d, g["u1"], g["r1"], g["phi1"], g["x"], g["y"], g["t"], batch=1, **kw)
    d_con = quiet(S.approximate_second_order_condition_2d, g["u1"], g["r1"], g["phi1"], g["x"], g["y"], g["t"], batch=4, **kw)
    assert d_seq == d_con and len(d_con) == 4
