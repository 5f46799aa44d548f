"""GPU edge cases through the C ABI: ragged sizes, single levels, empty horizons, argument errors — the shapes the
reference's tests poke at (tiny synthetic grids, M = 5) and the limits of the kernels' grid-stride / tail handling."""
import numpy as np
import pytest

import vch_oracle as O
from conftest import rel

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("count", [1, 2, 31, 255, 256, 257, 1025, 70001])
def test_prox_kkt_solve_w_ragged_counts(native, count):
    rng = np.random.default_rng(count)
    u, r = rng.standard_normal(count), 0.01 * rng.standard_normal(count)
    un, g, red = native.grad_prox(u, r, 0.3, 0.7, 0.05, -0.8, 0.9, want_grad=True)
    ref = O.soft_prox(u, r + 0.3 * u, 0.7, 0.05, -0.8, 0.9)
    assert np.array_equal(un, ref) and np.array_equal(g, r + 0.3 * u)
    assert red[2] == np.count_nonzero(ref) and abs(red[1] - np.sum(u * u)) <= 1e-12 * max(1.0, red[1])
    assert native.kkt_counts(un, r, 0.05) == O.kkt_counts(un, r, 0.05)
    w = native.solve_w(u, 1e-2, 10.0, r, un)
    assert np.array_equal(w, O.w_update(u, 1e-2, 10.0, r, un))


def test_degenerate_horizons_and_shapes(native):
    P = O.Phys2D(Nx=16, Ny=12, T=0.02)
    c = native.Ctx2D(P.Nx, P.Ny, 1 / 16, 1 / 12, 1.0, 1.0, P.tau, P.gamma, P.c1, P.c2, P.kappa)
    phi0 = O.init_phi_2d(16, 12)
    h0, _, _ = c.forward(phi0, None, np.zeros(0))                    # empty horizon: only level 0
    assert h0.shape == (1, 17, 13) and np.array_equal(h0[0], phi0)
    # control shorter than the horizon: rows beyond the end act as zeros (Forward2_solver.py:545-548)
    u_short = 0.3 * np.ones((2, 17, 13))
    dts = np.full(3, 1e-2)
    h_a, _, w_a = c.forward(phi0, u_short, dts, want_w=True)
    u_full = np.concatenate([u_short, np.zeros((2, 17, 13))])
    u_full[2:] = 0.0
    fw = O.forward_2d(O.Phys2D(Nx=16, Ny=12, T=0.03), u_short)
    assert rel(h_a, fw["phi"]) < 1e-8 and rel(w_a, fw["w"]) < 1e-12
    # single-level adjoint / cost
    x, y = np.linspace(0, 1, 17), np.linspace(0, 1, 13)
    p, q, r = c.adjoint(h_a[:1], np.zeros(1), 5.0, 10.0, None, None)
    po, qo, ro = O.adjoint_2d(P, h_a[:1], x, y, np.zeros(1), 5.0, 10.0)
    assert rel(p, po) < 1e-9 and rel(q, qo) < 1e-9 and np.abs(r).max() == 0
    J = c.cost(h_a[:1], 0 * h_a[:1], 0 * h_a[:1], 0 * phi0, x, y, np.zeros(1), 1.0, 2.0, 3.0, 4.0)
    Jo, parts = O.cost_2d(h_a[:1], 0 * h_a[:1], 0 * h_a[:1], 0 * phi0, x, y, np.zeros(1), O.Opt2D(b1=1, b2=2, b3=3, kappa_sparsity=4))
    assert abs(J[0] - Jo) <= 1e-12 * abs(Jo) and J[1] == 0.0
    # repeated time stamps: the adjoint copies the level (backward2_solver.py:214-216)
    t_rep = np.array([0.0, 0.01, 0.01, 0.02])
    p2, q2, r2 = c.adjoint(h_a, t_rep, 5.0, 10.0, None, None)
    assert np.array_equal(p2[1], p2[2]) and np.array_equal(r2[1], r2[2])
    po2, _, ro2 = O.adjoint_2d(P, h_a, x, y, t_rep, 5.0, 10.0)
    assert rel(p2, po2) < 1e-9 and rel(r2, ro2) < 1e-9
    with pytest.raises(ValueError):
        c.forward(phi0, np.zeros((3, 5, 5)), dts)
    with pytest.raises(ValueError):
        c.adjoint(np.zeros((4, 3, 3)), t_rep, 1.0, 1.0)
    with pytest.raises(TypeError):
        import torch
        c.residual(phi0, torch.zeros(17, 13, dtype=torch.float64, device="cuda"), phi0, phi0, phi0, phi0, 1e-2)


def test_1d_edge_cases(native):
    P = O.Phys1D(N=40, T=0.03)
    c = native.Ctx1D(P.N, P.Lx / P.N, P.Lx, P.tau, P.gamma, P.c1, P.c2, P.kappa)
    phi0 = O.init_phi_1d(40)
    h0, _, _ = c.forward(phi0, None, np.zeros(0))
    assert h0.shape == (2, 41) and np.array_equal(h0[0], phi0) and np.array_equal(h0[1], phi0)
    dts = np.full(3, 1e-2)
    u = 0.2 * np.sin(np.arange(5 * 41)).reshape(5, 41)
    h, mu, w = c.forward(phi0, u, dts, want_mu=True, want_w=True)
    fw = O.forward_1d(P, u)
    assert rel(h, fw["phi"]) < 1e-8 and rel(w, fw["w"]) < 1e-12
    u3 = u[:3]                                                        # rows == steps: the last row repeats (Forward_solver.py:351-353)
    h1, _, w1 = c.forward(phi0, u3, dts, want_w=True)
    fw1 = O.forward_1d(P, u3)
    assert rel(h1, fw1["phi"]) < 1e-8 and rel(w1, fw1["w"]) < 1e-12
    with pytest.raises(IndexError):                                   # fewer rows than steps: IndexError, like the reference
        c.forward(phi0, u[:2], dts)
    with pytest.raises(ValueError):
        c.forward(phi0, np.zeros((3, 7)), dts)
    big = native.Ctx1D(512, 1 / 512, 1.0, P.tau, P.gamma, P.c1, P.c2, P.kappa)      # N = 512 as in the reference's order test
    hb, _, _ = big.forward(O.init_phi_1d(512), None, np.full(2, 1e-3))
    fb = O.forward_1d(O.Phys1D(N=512, T=2e-3, dt_initial=1e-3))
    assert rel(hb, fb["phi"]) < 1e-8


def test_nonfinite_input_raises_and_context_recovers(native, golden):
    """the device-side non-finite flag must not stay set.  A NaN initial field raises RuntimeError
    (VCH_E_NONFINITE); the next valid call on the SAME context succeeds and reproduces the reference golden."""
    g = golden("g2d_32")
    P = O.from_json(O.Phys2D, g["cfg_json"])
    c = native.Ctx2D(P.Nx, P.Ny, P.Lx / P.Nx, P.Ly / P.Ny, P.Lx, P.Ly, P.tau, P.gamma, P.c1, P.c2, P.kappa)
    dts = np.diff(g["t"])
    bad = g["phi0"][0].copy()
    bad[3, 5] = np.nan
    with pytest.raises(RuntimeError):
        c.forward(bad, None, dts)
    with pytest.raises(RuntimeError):                                # NaN control: w -> residual -> non-finite
        u = np.zeros_like(g["phi0"]); u[1, 2, 2] = np.inf
        c.forward(g["phi0"][0], u, dts)
    h, _, _ = c.forward(g["phi0"][0], None, dts)                     # same context, valid input
    assert rel(h, g["phi0"]) < 1e-8
    p, q, r = c.adjoint(h, g["t"], 5.0, 10.0, None, None)
    assert np.isfinite(r).all() and c.last_stats["krylov_stalls"] == 0


def test_adjoint_krylov_stall_is_an_error(native, golden):
    """a BiCGStab solve of the adjoint sweep that stops above tolerance must not return silently (the reference's
    direct solve cannot fail that way): VCH_E_KRYLOV -> KrylovStall; the forward solve, guarded by Newton's own residual
    check, keeps returning normally; restoring the iteration cap restores the result."""
    g = golden("g2d_32")
    P = O.from_json(O.Phys2D, g["cfg_json"])
    c = native.Ctx2D(P.Nx, P.Ny, P.Lx / P.Nx, P.Ly / P.Ny, P.Lx, P.Ly, P.tau, P.gamma, P.c1, P.c2, P.kappa)
    dts = np.diff(g["t"])
    h, _, _ = c.forward(g["phi0"][0], None, dts)
    c.set_krylov(1e-17, 1)                                           # below fp64 resolution: cannot be reached
    with pytest.raises(native.KrylovStall):
        c.adjoint(h, g["t"], 5.0, 10.0, None, None)
    c.set_krylov(1e-11, 200)
    p, q, r = c.adjoint(h, g["t"], 5.0, 10.0, None, None)
    assert c.last_stats["krylov_stalls_adjoint"] == 0
    x, y = g["x"], g["y"]
    po, qo, ro = O.adjoint_2d(P, h, x, y, g["t"], 5.0, 10.0)
    assert rel(p, po) < 1e-7 and rel(r, ro) < 1e-7


@pytest.mark.timeout(120)
def test_zero_right_hand_side_ends_the_solve_graph(native):
    """A linear solve whose right-hand side is exactly zero is finished before its first iteration.  In the fused adjoint solve
    the start (||b||^2 -> done) lives in the kernel IN FRONT of the graph, so the WHILE body is entered once with `done` set
    and its last kernel has to clear the loop condition on its early return — otherwise the graph would spin.  b1 = b2 = 0 makes
    every adjoint right-hand side zero; the power-of-two grid takes the graph path with the radix-16 kernels."""
    P = O.Phys2D(Nx=32, Ny=32, T=0.04)
    c = native.Ctx2D(32, 32, 1 / 32, 1 / 32, 1.0, 1.0, P.tau, P.gamma, P.c1, P.c2, P.kappa)
    dts = np.full(4, 1e-2)
    hist, _, _ = c.forward(O.init_phi_2d(32, 32), None, dts)
    t = np.concatenate([[0.0], np.cumsum(dts)])
    p, q, r = c.adjoint(hist, t, 0.0, 0.0, np.zeros_like(hist), np.zeros_like(hist[0]))
    assert not p.any() and not q.any() and not r.any()
    assert c.last_stats["krylov_iterations"] == 0 and c.last_stats["krylov_stalls"] == 0
    # and the context is fine afterwards: an ordinary adjoint sweep equals the oracle's
    fw = O.forward_2d(P)
    phiT, phiQ = O.targets_2d(fw["x"], fw["y"], fw["t"], fw["phi"][0], P.Lx, P.Ly, P.T)
    Op = O.Opt2D()
    p2, q2, r2 = c.adjoint(hist, t, Op.b1, Op.b2, phiQ, phiT)
    ro = O.adjoint_2d(P, fw["phi"], fw["x"], fw["y"], fw["t"], Op.b1, Op.b2, phiQ, phiT)
    assert rel(r2, ro[2]) < 1e-7 and rel(hist, fw["phi"]) < 1e-8
