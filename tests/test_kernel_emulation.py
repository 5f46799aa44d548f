"""The CUDA transform kernels themselves, run on the CPU (tests/emu/cuda_emu.h: blocks as groups of fibers, barriers, warp
shuffles, shared memory) and checked against NumPy/SciPy:

  * `dct_fft_kernel` rows + fused column solve = the exact constant-coefficient solve P^-1 (and lambda P^-1);
  * the fused BiCGStab prologues (p = r + beta q, s = r - alpha v, coefficient multiply) and epilogues (+addend, adjoint-side
    multiply, dot products -> alpha / omega, half-step exit, done-flag gating), forward and adjoint form;

The emulation pins kernel LOGIC (index maps, flag protocol, reductions); memory ordering and timing need the GPU suite."""
import os
import struct
import subprocess

import numpy as np
import pytest
import scipy.fft as sf

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "emu")
CUDA_INC = os.environ.get("CUDA_HOME", "/usr/local/cuda") + "/include"
CUDA_LIB = os.environ.get("CUDA_HOME", "/usr/local/cuda") + "/lib64"


def _build(tmp, name, flags):
    exe = os.path.join(tmp, name)
    cmd = ["g++", "-std=c++17", "-O1", "-I" + CUDA_INC] + flags + [os.path.join(EMU, "dct_emu_harness.cpp"), "-o", exe,
           "-L" + CUDA_LIB, "-lcudart", "-Wl,-rpath," + CUDA_LIB]
    subprocess.run(cmd, check=True, capture_output=True)
    return exe


def _run(exe, lg, tmp, tag):
    out = os.path.join(tmp, f"out_{tag}_{lg}.bin")
    subprocess.run([exe, str(lg), out], check=True, timeout=600)
    res = {}
    with open(out, "rb") as fh:
        while True:
            hdr = fh.read(40)
            if len(hdr) < 40:
                break
            name = hdr[:32].split(b"\0")[0].decode()
            (cnt,) = struct.unpack("<Q", hdr[32:])
            res[name] = np.frombuffer(fh.read(8 * cnt), dtype=np.float64).copy()
    return res


@pytest.fixture(scope="module")
def builds(tmp_path_factory):
    if not os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h")):
        pytest.skip("CUDA headers not found")
    tmp = str(tmp_path_factory.mktemp("emu"))
    return tmp, {"default": _build(tmp, "h_default", [])}


rel = lambda a, b: float(np.linalg.norm(a - b) / np.linalg.norm(b))


@pytest.mark.parametrize("lg", [6, 7])
def test_transform_kernels_on_cpu_match_numpy(builds, lg):
    tmp, exe = builds
    R = _run(exe["default"], lg, tmp, "default")
    N = (1 << lg) // 2
    n = N + 1
    pitch = (n + 3) & ~3
    sq = lambda k: R[k].reshape(n, n)
    x, r, q, r0, a, lam = sq("in"), sq("r"), sq("q"), sq("r0"), sq("a"), R["lam"]
    L = lam[:, None] + lam[None, :]
    # A. row transform = unnormalised DCT-I of every line; fused solve = exact inverse of c0 + abar*lam + c2*lam^2
    rows = R["A_rows"].reshape(n, pitch)[:, :n]
    assert rel(rows, sf.dct(x, type=1, axis=1)) < 1e-14
    sym = 100.0 + L * (7.5 + 5e-5 * L)
    assert rel(sq("A_solve"), sf.idctn(sf.dctn(x, type=1) / sym, type=1)) < 1e-13
    assert rel(sq("A_lamsolve"), sf.idctn(sf.dctn(x, type=1) * L / sym, type=1)) < 1e-13
    # B/C. forward operator applications with the fused vector updates and dot products
    abar = 7.25
    symb = 100.0 + L * (abar + 5e-5 * L)
    K = lambda y: sf.idctn(sf.dctn(y, type=1) * L / symb, type=1)
    beta = (1.7 / 0.9) * (0.6 / 1.3)
    p = r + beta * q
    v = K((a - abar) * p) + p
    assert rel(sq("B_p"), p) < 1e-15 and rel(sq("B_v"), v) < 1e-13
    alpha, r0v, rho, done, half = R["B_scal"][:5]
    assert abs(r0v - np.vdot(r0, v)) < 1e-11 * abs(r0v) and abs(alpha - 1.7 / r0v) < 1e-12 * abs(alpha) and rho == 1.7
    assert done == 0 and half == 0
    s = r - 0.6 * q
    t = K((a - abar) * s) + s
    assert rel(sq("C_s"), s) < 1e-15 and rel(sq("C_t"), t) < 1e-13
    omega, ts, tt = R["C_scal"][:3]
    assert abs(ts - np.vdot(t, s)) < 1e-11 * abs(ts) and abs(tt - np.vdot(t, t)) < 1e-11 * tt and abs(omega - ts / tt) < 1e-12 * abs(omega)
    # D. beta = 0: p = r bit for bit, nothing of the poisoned q arrives anywhere
    assert np.array_equal(sq("D_p"), r) and np.isfinite(sq("D_v")).all()
    assert rel(sq("D_v"), K((a - abar) * r) + r) < 1e-13
    # E. adjoint (right-preconditioned) form: multiply after the transform
    pa = r + beta * q
    va = (a - abar) * K(pa) + pa
    assert rel(sq("E_p"), pa) < 1e-15 and rel(sq("E_v"), va) < 1e-13
    al = R["E_scal"][0]
    assert abs(al - 1.7 / np.vdot(r0, va)) < 1e-11 * abs(al)
    sa = r - al * va
    ta = (a - abar) * K(sa) + sa
    assert rel(sq("E_s"), sa) < 1e-12 and rel(sq("E_t"), ta) < 1e-12
    assert abs(R["E_scal2"][0] - np.vdot(ta, sa) / np.vdot(ta, ta)) < 1e-10 * abs(R["E_scal2"][0])
    # F. half-step exit fires, and kernels gated on the done flag leave their outputs alone
    alpha_f, done_f, half_f, r0v_rel = R["F_scal"][:4]
    assert abs(alpha_f - 1.0) < 1e-12 and done_f == 1 and half_f == 1 and abs(r0v_rel - 1.0) < 1e-12
    assert np.all(R["F_keep"] == 42.0) and np.all(R["F_keepw"] == 43.0)


# ------------------------------------------------------------------------------------------------ Newton linear solve / adjoint step
import sys
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import krylov_proto as KP   # noqa: E402  (NumPy model of the algorithm; here only its stencil and spectral helpers)


@pytest.fixture(scope="module")
def krylov_exe(tmp_path_factory):
    if not os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h")):
        pytest.skip("CUDA headers not found")
    tmp = str(tmp_path_factory.mktemp("emuk"))
    exe = os.path.join(tmp, "k_default")
    subprocess.run(["g++", "-std=c++17", "-O1", "-I" + CUDA_INC, os.path.join(EMU, "krylov_emu_harness.cpp"), "-o", exe,
                    "-L" + CUDA_LIB, "-lcudart", "-Wl,-rpath," + CUDA_LIB], check=True, capture_output=True)
    return tmp, exe


def _read(path):
    res = {}
    with open(path, "rb") as fh:
        while True:
            hdr = fh.read(40)
            if len(hdr) < 40:
                break
            (cnt,) = struct.unpack("<Q", hdr[32:])
            res[hdr[:32].split(b"\0")[0].decode()] = np.frombuffer(fh.read(8 * cnt), dtype=np.float64).copy()
    return res


@pytest.mark.parametrize("lg,tol,half", [(6, 1e-11, 1), (6, 1e-11, 0), (6, 1e-6, 1), (7, 1e-11, 1)])
def test_newton_linear_solve_and_adjoint_step_on_cpu(krylov_exe, lg, tol, half):
    """residual -> Schur rhs -> P^-1 b -> BiCGStab (device-side done / half-step flags) -> delta-mu, ceiling, trial iterate, and one
    adjoint step, through the real kernels on the CPU emulator, against NumPy."""
    tmp, exe = krylov_exe
    out = os.path.join(tmp, f"k_{lg}_{tol}_{half}.bin")
    subprocess.run([exe, str(lg), repr(tol), str(half), out], check=True, timeout=900)
    R = _read(out)
    N = (1 << lg) // 2
    n = N + 1
    G = KP.Grid(N, kappa=1e-4, tau=0.05, gamma=10.0, c1=0.75, c2=1.0, dt=1e-2)
    sq = lambda k: R[k].reshape(n, n)
    phi, mu, cphi, cmu = sq("phi"), sq("mu"), sq("cphi"), sq("cmu")
    dt, tdt = 1e-2, 0.05 / 1e-2
    # residual, Jacobian diagonal, reductions, published mirror
    rp = tdt * phi - 0.5 * G.kappa * G.lap(phi) + G.c1 * G.flog(phi) - 0.5 * mu + cphi
    rm = phi / dt - 0.5 * G.lap(mu) + cmu
    a = tdt + 2 * G.c1 / (1 - np.minimum(phi ** 2, 1 - 1e-4))
    assert rel(sq("Rphi"), rp) < 1e-13 and rel(sq("Rmu"), rm) < 1e-13 and rel(sq("a"), a) < 1e-15
    res2, amin, amax, abar, mu2, mirror_ok = R["res_scal"]
    assert abs(res2 - (np.sum(rp ** 2) + np.sum(rm ** 2))) < 1e-12 * res2 and amin == sq("a").min() and amax == sq("a").max()
    assert abs(abar - np.sqrt(amin * amax)) < 1e-15 * abar and abs(mu2 - np.sum(mu ** 2)) < 1e-12 * mu2 and mirror_ok == 1
    # Schur right-hand side and the solve
    b = G.lap(sq("Rphi")) - sq("Rmu")
    assert rel(sq("b"), b) < 1e-12
    A = lambda x: x / dt - G.lap(sq("a") * x) + 0.5 * G.kappa * G.lap(G.lap(x))
    dphi = sq("dphi")
    x_ref, its_ref = KP.bicgstab(G, sq("a"), sq("b"), abar, tol=tol)
    its, done, half_exits, half_flag, solves, launched, unchanged, nonfinite = R["fwd_scal"]
    assert done == 1 and half_flag == 0 and solves == 1 and unchanged == 1 and nonfinite == 0
    assert its == np.ceil(its_ref) if half else its >= np.ceil(its_ref)
    assert half_exits == (1 if (half and its_ref != np.floor(its_ref)) else 0)
    assert rel(dphi, x_ref) < 50 * tol and np.linalg.norm(A(dphi) - sq("b")) < 1e3 * tol * np.linalg.norm(sq("b"))
    # delta-mu, full-step trial iterate, step ceiling
    dmu = 2 * (sq("a") * dphi - 0.5 * G.kappa * G.lap(dphi) + sq("Rphi"))
    assert rel(sq("dmu"), dmu) < 1e-12
    assert np.array_equal(sq("phit"), phi + dphi) and np.array_equal(sq("mut"), mu + sq("dmu"))
    pos, neg = dphi > 0, dphi < 0
    assert abs(R["ceil"][0] - np.min((0.99 - phi[pos]) / dphi[pos])) < 1e-12 * abs(R["ceil"][0])
    assert abs(R["ceil"][1] - np.min((-0.99 - phi[neg]) / dphi[neg])) < 1e-12 * abs(R["ceil"][1])
    # adjoint step
    fpp = lambda f: 2 * G.c1 / (1 - np.clip(f, -(1 - 1e-8), 1 - 1e-8) ** 2) - 2 * G.c2
    p1, q1, phi1 = sq("p1"), sq("q1"), sq("phi1")
    assert rel(q1, -G.lap(p1)) < 1e-13
    rhs = p1 + 0.05 * q1 + 0.5 * dt * G.lap(q1) - 0.5 * dt * fpp(phi1) * q1 + 0.5 * dt * 5.0 * (phi + phi1)
    aa = 0.05 + 0.5 * dt * fpp(phi)
    assert rel(sq("adj_rhs"), rhs) < 1e-12 and rel(sq("adj_a"), aa) < 1e-14
    Aadj = lambda p: p - sq("adj_a") * G.lap(p) + 0.5 * dt * G.lap(G.lap(p))
    ap = sq("adj_p")
    assert np.linalg.norm(Aadj(ap) - sq("adj_rhs")) < 1e-10 * np.linalg.norm(sq("adj_rhs"))
    a_its, a_done, a_half, a_abar = R["adj_scal"]
    assert a_done == 1 and a_half == half_exits and 1 <= a_its <= 6      # the counter is cumulative: the adjoint solve adds no half-step exit and abs(a_abar - np.sqrt(aa.min() * aa.max())) < 1e-14 * a_abar
