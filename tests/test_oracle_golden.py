"""CPU tests: the oracle restatement (oracle/vch_oracle.py) pinned against golden vectors that the UNMODIFIED
reference produced (oracle/make_golden.py).  This is what makes the oracle a trustworthy checker for the GPU tests."""
import numpy as np
import pytest

import vch_oracle as O
from conftest import rel


@pytest.mark.parametrize("name", ["g2d_rect", "g2d_32"])
def test_oracle_2d_matches_reference(golden, name):
    g = golden(name)
    P = O.from_json(O.Phys2D, g["cfg_json"])
    Op = O.from_json(O.Opt2D, g["opt_json"])
    fw = O.forward_2d(P)
    assert np.array_equal(fw["t"], g["t"]) and np.array_equal(fw["nres"], g["nres0"])
    assert rel(fw["phi"], g["phi0"]) < 1e-12 and rel(fw["mu"], g["mu0"]) < 1e-12
    phiT, phiQ = O.targets_2d(fw["x"], fw["y"], fw["t"], fw["phi"][0], P.Lx, P.Ly, P.T)
    assert np.array_equal(phiT, g["phiT"])
    u0 = np.zeros_like(fw["phi"])
    J0, _ = O.cost_2d(fw["phi"], u0, phiQ, phiT, fw["x"], fw["y"], fw["t"], Op)
    assert abs(J0 - g["J"][0]) < 1e-12 * abs(g["J"][0])
    p, q, r = O.adjoint_2d(P, fw["phi"], fw["x"], fw["y"], fw["t"], Op.b1, Op.b2, phiQ, phiT)
    assert rel(p, g["p0"]) < 1e-10 and rel(q, g["q0"]) < 1e-10 and rel(r, g["r0"]) < 1e-10
    u1, fw1, J1, _ = O.pgd_iter_2d(P, Op, u0, fw["phi"], fw["t"], fw["x"], fw["y"], phiQ, phiT, Op.alpha_max)
    assert rel(u1, g["u1"]) < 1e-10 and np.array_equal(u1 != 0, g["u1"] != 0)
    assert rel(fw1["phi"], g["phi1"]) < 1e-10 and rel(fw1["w"], g["w1"]) < 1e-10
    assert abs(J1 - g["J"][1]) < 1e-10 * abs(g["J"][1])


def test_oracle_1d_matches_reference(golden):
    g = golden("g1d_n64")
    P = O.from_json(O.Phys1D, g["cfg_json"])
    Op = O.from_json(O.Opt1D, g["opt_json"])
    fw = O.forward_1d(P)
    assert fw["phi"].shape == g["phi0"].shape and np.array_equal(fw["t"], g["t"])
    assert rel(fw["phi"], g["phi0"]) < 1e-9 and rel(fw["mu"], g["mu0"]) < 1e-9      # 1D is chaotic at 1e-11
    phiT, phiQ = O.targets_1d(fw["x"], fw["t"], fw["phi"][0], P.Lx)
    assert np.array_equal(phiT, g["phiT"]) and rel(phiQ, g["phiQ"]) < 1e-9
    p, q, r = O.adjoint_1d(g["phi0"], g["x"], g["t"], Op.b1, Op.b2, g["phiQ"], g["phiT"])
    assert rel(p, g["p0"]) < 1e-9 and rel(r, g["r0"]) < 1e-9 and np.abs(p[0]).max() == 0
    u0 = np.zeros_like(g["phi0"])
    J0, _ = O.cost_1d(g["phi0"], u0, g["phiQ"], g["phiT"], g["x"], g["t"], Op.b1, Op.b2, Op.b3, Op.kappa_sparsity)
    assert abs(J0 - g["J"][0]) < 1e-12 * abs(g["J"][0])
    u1 = O.soft_prox(u0, g["r0"] + Op.b3 * u0, Op.alpha_max, Op.kappa_sparsity, Op.u_min, Op.u_max)
    assert np.array_equal(u1, g["u1"])
    fw1 = O.forward_1d(P, g["u1"])
    assert rel(fw1["phi"], g["phi1"]) < 1e-9 and rel(fw1["w"], g["w1"]) < 1e-12
    J1, _ = O.cost_1d(g["phi1"], g["u1"], g["phiQ"], g["phiT"], g["x"], g["t"], Op.b1, Op.b2, Op.b3, Op.kappa_sparsity)
    assert abs(J1 - g["J"][1]) < 1e-12 * abs(g["J"][1])


def test_oracle_default_1d_anchor_values(golden):
    """BASELINE.md parity anchors for the reference's 1D default run."""
    g = golden("g1d_default")
    assert abs(g["J"][0] - 1.0977825116186202) < 1e-15 and abs(g["J"][1] - 0.3218003723414755) < 1e-15
    Op = O.Opt1D()
    J0, parts = O.cost_1d(g["phi0"], np.zeros_like(g["phi0"]), g["phiQ"], g["phiT"], g["x"], g["t"], Op.b1, Op.b2, Op.b3, Op.kappa_sparsity)
    assert abs(J0 - g["J"][0]) < 1e-12 and parts[2] == 0 and parts[3] == 0
    assert O.kkt_counts(g["u1"], g["r0"], Op.kappa_sparsity)[2] <= g["u1"].size


def test_golden_128_anchor_values(golden):
    g = golden("g2d_128")
    assert abs(g["J"][0] - 3.0192834439912475) < 1e-15 and abs(g["J"][1] - 1.5416947343810048) < 1e-15
    assert g["phi0"].shape == (5, 129, 129) and list(g["keep"]) == [0, 1, 50, 99, 100]
    frac_bounds = np.mean(np.abs(g["u1"]) == 1.0)
    assert 0.2 < frac_bounds < 0.35            # 28.3 % of u1 at the bounds (full array)
