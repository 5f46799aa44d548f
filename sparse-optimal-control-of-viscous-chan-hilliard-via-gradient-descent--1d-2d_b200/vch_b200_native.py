"""ctypes binding of libvch_b200.so (include/vch_b200.h) — the only door between the Python drop-in
modules and the sm_100a CUDA kernels.

There is no CPU fallback: if the library is missing, or no CUDA device is visible, every compute call raises.
Arrays may be NumPy (host; the library stages H2D/D2H on its stream) or torch CUDA tensors (device pointers are
passed through, nothing is copied); one call uses one kind for all of its array arguments.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvch_b200.so")

MEM_HOST, MEM_DEVICE = 0, 1
E_CUDA, E_SHAPE, E_NONFINITE, E_KRYLOV, E_ARG = 1, 2, 3, 4, 5


class KrylovStall(ArithmeticError):
    """The matrix-free linear solve stopped above its tolerance (no analogue in the reference's direct solve)."""


class Params2D(C.Structure):
    _fields_ = [("Nx", C.c_int), ("Ny", C.c_int), ("hx", C.c_double), ("hy", C.c_double), ("Lx", C.c_double),
                ("Ly", C.c_double), ("tau", C.c_double), ("gamma", C.c_double), ("c1", C.c_double), ("c2", C.c_double),
                ("kappa", C.c_double), ("delta_sep", C.c_double)]


class Params1D(C.Structure):
    _fields_ = [("N", C.c_int), ("h", C.c_double), ("Lx", C.c_double), ("tau", C.c_double), ("gamma", C.c_double),
                ("c1", C.c_double), ("c2", C.c_double), ("kappa", C.c_double), ("delta_sep", C.c_double)]


class Stats(C.Structure):
    _fields_ = [("newton_residual_evals", C.c_longlong), ("newton_linear_solves", C.c_longlong),
                ("krylov_iterations", C.c_longlong), ("krylov_max_iterations", C.c_longlong),
                ("kernel_launches", C.c_longlong), ("krylov_stalls", C.c_longlong), ("last_newton_residual", C.c_double),
                ("krylov_half_exits", C.c_longlong), ("krylov_stalls_adjoint", C.c_longlong)]

    def as_dict(self):
        return {f: getattr(self, f) for f, _ in self._fields_}


# every symbol include/vch_b200.h declares (tests check the library exports exactly these)
EXPORTS = [
    "vch_last_error", "vch_device_count", "vch_version",
    "vch2d_create", "vch2d_destroy", "vch2d_set_stream", "vch2d_set_krylov", "vch2d_set_krylov_first", "vch2d_set_newton", "vch2d_set_stream_budget", "vch2d_launch_count", "vch2d_profile", "vch2d_profile_report",
    "vch2d_apply_laplacian", "vch2d_initialize_mu", "vch_solve_w", "vch2d_residual", "vch2d_jacobian_solve",
    "vch2d_newton", "vch2d_forward", "vch2d_adjoint", "vch2d_cost", "vch_grad_prox", "vch_kkt_counts", "vch_free_energy",
    "vch2d_pgd_iteration", "vch2d_forward_ckpt", "vch2d_pgd_iteration_ckpt",
    "vch2d_slab_create", "vch2d_slab_rows", "vch2d_slab_ipc_handle", "vch2d_slab_attach", "vch2d_slab_selftest",
    "vch1d_create", "vch1d_destroy", "vch1d_set_stream", "vch1d_launch_count", "vch1d_residual", "vch1d_initialize_mu", "vch1d_newton",
    "vch1d_forward", "vch1d_adjoint", "vch1d_cost", "vch1d_grad_prox",
]

_lib = None
_lock = threading.Lock()


def lib():
    """Load the shared library (once).  Fails loudly when it has not been built."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise ImportError(
                    f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                    "(nvcc, sm_100a).  vch_b200 has no CPU fallback.")
            L = C.CDLL(LIB_PATH)
            L.vch_last_error.restype = C.c_char_p
            L.vch2d_launch_count.restype = C.c_longlong
            L.vch1d_launch_count.restype = C.c_longlong
            L.vch2d_launch_count.argtypes = [C.c_void_p]
            L.vch1d_launch_count.argtypes = [C.c_void_p]
            L.vch2d_destroy.argtypes = [C.c_void_p]
            L.vch1d_destroy.argtypes = [C.c_void_p]
            L.vch2d_destroy.restype = None
            L.vch1d_destroy.restype = None
            _lib = L
    return _lib


def device_count() -> int:
    return int(lib().vch_device_count())


def require_device():
    if device_count() < 1:
        raise RuntimeError("vch_b200: no CUDA device visible; this build has no CPU fallback")


def _check(rc):
    if rc == 0:
        return
    msg = lib().vch_last_error().decode("utf-8", "replace")
    if rc == E_SHAPE:
        raise ValueError(msg)
    if rc == E_NONFINITE:
        raise RuntimeError(msg)
    if rc == E_KRYLOV:
        raise KrylovStall(msg)
    raise RuntimeError(f"vch_b200 error {rc}: {msg}")


def _is_dev(a):
    return a is not None and hasattr(a, "data_ptr")


def _mem_of(*arrays):
    kinds = {_is_dev(a) for a in arrays if a is not None}
    if len(kinds) > 1:
        raise TypeError("mix of NumPy arrays and CUDA tensors in one call")
    return MEM_DEVICE if (kinds and kinds.pop()) else MEM_HOST


class _Args:
    """Keeps converted arrays alive for the duration of a call."""

    def __init__(self):
        self.keep = []

    def inp(self, a, shape=None):
        if a is None:
            return C.c_void_p(None)
        if _is_dev(a):
            import torch
            assert a.is_cuda and a.dtype == torch.float64 and a.is_contiguous(), "device arrays must be contiguous float64 CUDA tensors"
            if shape is not None and tuple(a.shape) != tuple(shape):
                raise ValueError(f"expected shape {tuple(shape)}, got {tuple(a.shape)}")
            self.keep.append(a)
            return C.c_void_p(a.data_ptr())
        b = np.ascontiguousarray(a, dtype=np.float64)
        if shape is not None and b.shape != tuple(shape):
            raise ValueError(f"expected shape {tuple(shape)}, got {b.shape}")
        self.keep.append(b)
        return b.ctypes.data_as(C.c_void_p)

    def out(self, like, shape):
        """Allocate an output of the same kind (NumPy / torch) as `like`."""
        if _is_dev(like):
            import torch
            t = torch.empty(tuple(shape), dtype=torch.float64, device=like.device)
            self.keep.append(t)
            return t, C.c_void_p(t.data_ptr())
        o = np.empty(tuple(shape), dtype=np.float64)
        self.keep.append(o)
        return o, o.ctypes.data_as(C.c_void_p)

    def host(self, a):
        b = np.ascontiguousarray(a, dtype=np.float64)
        self.keep.append(b)
        return b.ctypes.data_as(C.c_void_p)


def _current_stream_ptr():
    """torch's current CUDA stream when torch is in use (so torch.cuda.Event brackets our kernels), else the default stream."""
    try:
        import sys
        if "torch" in sys.modules:
            import torch
            if torch.cuda.is_available():
                return C.c_void_p(torch.cuda.current_stream().cuda_stream)
    except Exception:
        pass
    return C.c_void_p(None)


# ------------------------------------------------------------------------------------------------ context-free ops
def solve_w(w_old, dt, gamma, u_n, u_np1):
    """Forward2_solver.py:170-181 / Forward_solver.py:88-91."""
    a = _Args()
    mem = _mem_of(w_old, u_n, u_np1)
    shape = tuple(w_old.shape)
    out, po = a.out(w_old, shape)
    n = int(np.prod(shape))
    _check(lib().vch_solve_w(_current_stream_ptr(), C.c_longlong(n), a.inp(w_old), C.c_double(dt), C.c_double(gamma),
                             a.inp(u_n, shape), a.inp(u_np1, shape), po, mem))
    return out


def grad_prox(u, r, b3, alpha, kappa_sp, u_min, u_max, want_grad=False):
    """r + b3 u, then soft-threshold + box.  Returns (u_new, grad|None, red[4])."""
    a = _Args()
    mem = _mem_of(u, r)
    shape = tuple(u.shape)
    un, pun = a.out(u, shape)
    g, pg = a.out(u, shape) if want_grad else (None, C.c_void_p(None))
    red = np.zeros(4)
    _check(lib().vch_grad_prox(_current_stream_ptr(), C.c_longlong(int(np.prod(shape))), a.inp(u), a.inp(r, shape),
                               C.c_double(b3), C.c_double(alpha), C.c_double(kappa_sp), C.c_double(u_min), C.c_double(u_max),
                               pg, pun, red.ctypes.data_as(C.c_void_p), mem))
    return un, g, red


def kkt_counts(u, r, kappa_sp, tol=1e-6):
    a = _Args()
    mem = _mem_of(u, r)
    cnt = np.zeros(3, dtype=np.int64)
    _check(lib().vch_kkt_counts(_current_stream_ptr(), C.c_longlong(int(np.prod(u.shape))), a.inp(u), a.inp(r, tuple(u.shape)),
                                C.c_double(kappa_sp), C.c_double(tol), cnt.ctypes.data_as(C.c_void_p), mem))
    return int(cnt[0]), int(cnt[1]), int(cnt[2])


def free_energy(phi, kappa, c1, c2, hx, hy, w=None, eps=1e-8):
    """Discrete free energy, one fused reduction kernel (Forward2_solver.py:256-319).  phi: (Ny+1, Nx+1) array or CUDA tensor."""
    a = _Args()
    mem = _mem_of(phi, w)
    n0, n1 = int(phi.shape[0]), int(phi.shape[1])
    E = np.zeros(1)
    _check(lib().vch_free_energy(_current_stream_ptr(), n0, n1, a.inp(phi), a.inp(w, (n0, n1) if w is not None else None),
                                 C.c_double(kappa), C.c_double(c1), C.c_double(c2), C.c_double(hx), C.c_double(hy),
                                 C.c_double(eps), E.ctypes.data_as(C.c_void_p), mem))
    return float(E[0])


# ------------------------------------------------------------------------------------------------ 2D context
class Ctx2D:
    def __init__(self, Nx, Ny, hx, hy, Lx, Ly, tau, gamma, c1, c2, kappa, delta_sep=1e-2, device=0):
        require_device()
        self.p = Params2D(int(Nx), int(Ny), float(hx), float(hy), float(Lx), float(Ly), float(tau), float(gamma),
                          float(c1), float(c2), float(kappa), float(delta_sep))
        self.shape = (int(Nx) + 1, int(Ny) + 1)
        self.h = C.c_void_p()
        _check(lib().vch2d_create(C.byref(self.p), int(device), C.byref(self.h)))
        self.last_stats = {}

    def __del__(self):
        try:
            if getattr(self, "h", None) and self.h.value:
                lib().vch2d_destroy(self.h)
                self.h = C.c_void_p()
        except Exception:
            pass

    def _stream(self):
        _check(lib().vch2d_set_stream(self.h, _current_stream_ptr()))

    def set_krylov(self, rel_tol=1e-11, max_iter=200):
        _check(lib().vch2d_set_krylov(self.h, C.c_double(rel_tol), int(max_iter)))

    def set_stream_budget(self, nbytes=0):
        """Cap the device staging of the host-buffer pgd_iteration (0 = automatic); see include/vch_b200.h."""
        _check(lib().vch2d_set_stream_budget(self.h, C.c_longlong(int(nbytes))))

    def set_krylov_first(self, rel_tol=1e-6):
        """Tolerance of the first linear solve of each Newton solve in the time loop (0 = always the full tolerance)."""
        _check(lib().vch2d_set_krylov_first(self.h, C.c_double(rel_tol)))

    def set_newton(self, floor_aware=True):
        _check(lib().vch2d_set_newton(self.h, 1 if floor_aware else 0))

    def launches(self):
        return int(lib().vch2d_launch_count(self.h))

    def profile(self, enable=True):
        _check(lib().vch2d_profile(self.h, 1 if enable else 0))

    def profile_report(self):
        """{kernel name: (total ms, launches)} since profiling was enabled."""
        names = C.create_string_buffer(4096); ms = (C.c_double * 64)(); cnt = (C.c_longlong * 64)(); n = C.c_int(0)
        _check(lib().vch2d_profile_report(self.h, names, 4096, ms, cnt, 64, C.byref(n)))
        keys = names.value.decode().split(";") if n.value else []
        return {k: (ms[i], int(cnt[i])) for i, k in enumerate(keys)}

    def apply_laplacian(self, v):
        a = _Args(); self._stream()
        out, po = a.out(v, self.shape)
        _check(lib().vch2d_apply_laplacian(self.h, a.inp(v, self.shape), po, _mem_of(v)))
        return out

    def initialize_mu(self, phi, w):
        a = _Args(); self._stream()
        out, po = a.out(phi, self.shape)
        _check(lib().vch2d_initialize_mu(self.h, a.inp(phi, self.shape), a.inp(w, self.shape), po, _mem_of(phi, w)))
        return out

    def residual(self, phi_new, phi_old, mu_new, mu_old, w_new, w_old, dt):
        a = _Args(); self._stream()
        s = self.shape
        rp, prp = a.out(phi_new, s); rm, prm = a.out(phi_new, s)
        _check(lib().vch2d_residual(self.h, a.inp(phi_new, s), a.inp(phi_old, s), a.inp(mu_new, s), a.inp(mu_old, s),
                                    a.inp(w_new, s), a.inp(w_old, s), C.c_double(dt), prp, prm,
                                    _mem_of(phi_new, phi_old, mu_new, mu_old, w_new, w_old)))
        return rp, rm

    def jacobian_solve(self, phi, dt, Rphi, Rmu):
        """Solve J(phi) [dphi; dmu] = -[Rphi; Rmu].  Returns (dphi, dmu, krylov_iterations)."""
        a = _Args(); self._stream()
        s = self.shape
        d1, p1 = a.out(phi, s); d2, p2 = a.out(phi, s)
        its = C.c_int(0)
        _check(lib().vch2d_jacobian_solve(self.h, a.inp(phi, s), C.c_double(dt), a.inp(Rphi, s), a.inp(Rmu, s), p1, p2,
                                          C.byref(its), _mem_of(phi, Rphi, Rmu)))
        return d1, d2, its.value

    def newton(self, phi_old, mu_old, w_old, w_new, dt, hist_cap=512):
        a = _Args(); self._stream()
        s = self.shape
        pn, ppn = a.out(phi_old, s); mn, pmn = a.out(phi_old, s)
        hist = np.zeros(hist_cap); nh = C.c_int(0); st = Stats()
        _check(lib().vch2d_newton(self.h, a.inp(phi_old, s), a.inp(mu_old, s), a.inp(w_old, s), a.inp(w_new, s),
                                  C.c_double(dt), ppn, pmn, hist.ctypes.data_as(C.c_void_p), hist_cap, C.byref(nh),
                                  C.byref(st), _mem_of(phi_old, mu_old, w_old, w_new)))
        self.last_stats = st.as_dict()
        return pn, mn, [float(v) for v in hist[:min(nh.value, hist_cap)]]

    def forward(self, phi0, u, dt_steps, want_mu=False, want_w=False):
        """Time loop.  Returns (phi_hist (M+1,..), mu_hist|None, w_hist|None)."""
        a = _Args(); self._stream()
        s = self.shape
        dts = np.ascontiguousarray(dt_steps, dtype=np.float64)
        M = int(dts.shape[0])
        if u is not None and (u.ndim != 3 or tuple(u.shape[1:]) != s):
            raise ValueError(f"control_input must have shape (M, {s[0]}, {s[1]})")
        hist, ph = a.out(phi0, (M + 1,) + s)
        mh, pm = a.out(phi0, (M,) + s) if want_mu else (None, C.c_void_p(None))
        wh, pw = a.out(phi0, (M,) + s) if want_w else (None, C.c_void_p(None))
        st = Stats()
        _check(lib().vch2d_forward(self.h, a.inp(phi0, s), a.inp(u), int(u.shape[0]) if u is not None else 0, M,
                                   a.host(dts), ph, pm, pw, C.byref(st), _mem_of(phi0, u)))
        self.last_stats = st.as_dict()
        return hist, mh, wh

    def adjoint(self, phi_hist, t_hist, b1, b2, phiQ=None, phiT=None, want_pq=True):
        a = _Args(); self._stream()
        s = self.shape
        lv = int(phi_hist.shape[0])
        full = (lv,) + s
        r, pr = a.out(phi_hist, full)
        p, pp = a.out(phi_hist, full) if want_pq else (None, C.c_void_p(None))
        q, pq = a.out(phi_hist, full) if want_pq else (None, C.c_void_p(None))
        st = Stats()
        _check(lib().vch2d_adjoint(self.h, a.inp(phi_hist, full), lv, a.host(t_hist), C.c_double(b1), C.c_double(b2),
                                   a.inp(phiQ, full if phiQ is not None else None), a.inp(phiT, s if phiT is not None else None),
                                   pp, pq, pr, C.byref(st), _mem_of(phi_hist, phiQ, phiT)))
        self.last_stats = st.as_dict()
        return p, q, r

    def cost(self, phi_hist, u, phiQ, phiT, x, y, t_hist, b1, b2, b3, kappa_sp):
        a = _Args(); self._stream()
        s = self.shape
        lv = int(phi_hist.shape[0])
        full = (lv,) + s
        J = np.zeros(9)
        _check(lib().vch2d_cost(self.h, a.inp(phi_hist, full), a.inp(u, full), a.inp(phiQ, full), a.inp(phiT, s), lv,
                                a.host(x), a.host(y), a.host(t_hist), C.c_double(b1), C.c_double(b2), C.c_double(b3),
                                C.c_double(kappa_sp), J.ctypes.data_as(C.c_void_p), _mem_of(phi_hist, u, phiQ, phiT)))
        return J

    def pgd_iteration(self, u, phi_hist, phiQ, phiT, t_hist, dt_steps, x, y, b1, b2, b3, kappa_sp, u_min, u_max, alpha,
                      u_out=None, phi_out=None, r_out=None):
        """One optimistic PGD iteration (GD2_configured.py:299-313).  Returns (u_new, phi_hist_new, J[5], red[4], stats)."""
        a = _Args(); self._stream()
        s = self.shape
        lv = int(phi_hist.shape[0])
        full = (lv,) + s
        if u_out is None:
            u_out, pun = a.out(u, full)
        else:
            pun = a.inp(u_out, full)
        if phi_out is None:
            phi_out, phn = a.out(u, full)
        else:
            phn = a.inp(phi_out, full)
        pr = a.inp(r_out, full) if r_out is not None else C.c_void_p(None)
        J = np.zeros(9); red = np.zeros(4); st = Stats()
        _check(lib().vch2d_pgd_iteration(self.h, lv, a.host(t_hist), a.host(dt_steps), a.host(x), a.host(y),
                                         a.inp(u, full), a.inp(phi_hist, full), a.inp(phiQ, full), a.inp(phiT, s),
                                         C.c_double(b1), C.c_double(b2), C.c_double(b3), C.c_double(kappa_sp),
                                         C.c_double(u_min), C.c_double(u_max), C.c_double(alpha), pun, phn, pr,
                                         J.ctypes.data_as(C.c_void_p), red.ctypes.data_as(C.c_void_p), C.byref(st),
                                         _mem_of(u, phi_hist, phiQ, phiT)))
        self.last_stats = st.as_dict()
        return u_out, phi_out, J, red, self.last_stats


def _ckpt_count(levels, stride):
    return (levels - 1 + stride - 1) // stride + 1


def _ctx2d_forward_ckpt(self, phi0, u, dt_steps, stride):
    """Forward solve that keeps only checkpoints every `stride` levels (include/vch_b200.h).  Returns (ck_phi, ck_mu, ck_w)."""
    a = _Args(); self._stream()
    s = self.shape
    dts = np.ascontiguousarray(dt_steps, dtype=np.float64)
    M = int(dts.shape[0])
    ncp = _ckpt_count(M + 1, int(stride))
    if u is not None and (u.ndim != 3 or tuple(u.shape[1:]) != s):
        raise ValueError(f"control_input must have shape (M, {s[0]}, {s[1]})")
    cp, pcp = a.out(phi0, (ncp,) + s); cm, pcm = a.out(phi0, (ncp,) + s); cw, pcw = a.out(phi0, (ncp,) + s)
    st = Stats()
    _check(lib().vch2d_forward_ckpt(self.h, a.inp(phi0, s), a.inp(u), int(u.shape[0]) if u is not None else 0, M, a.host(dts),
                                    int(stride), pcp, pcm, pcw, C.byref(st), _mem_of(phi0, u)))
    self.last_stats = st.as_dict()
    return cp, cm, cw


def _ctx2d_pgd_iteration_ckpt(self, u, ck, stride, phiQ, phiT, t_hist, dt_steps, x, y, b1, b2, b3, kappa_sp, u_min, u_max, alpha,
                              u_out=None, r_out=None):
    """One optimistic PGD iteration on a checkpointed trajectory.  ck = (ck_phi, ck_mu, ck_w).
    Returns (u_new, (ck_phi, ck_mu, ck_w) of the new trajectory, J[9], red[4], stats)."""
    a = _Args(); self._stream()
    s = self.shape
    lv = int(u.shape[0])
    full = (lv,) + s
    ncp = _ckpt_count(lv, int(stride))
    cshape = (ncp,) + s
    if u_out is None:
        u_out, pun = a.out(u, full)
    else:
        pun = a.inp(u_out, full)
    cp, pcp = a.out(u, cshape); cm, pcm = a.out(u, cshape); cw, pcw = a.out(u, cshape)
    pr = a.inp(r_out, full) if r_out is not None else C.c_void_p(None)
    J = np.zeros(9); red = np.zeros(4); st = Stats()
    _check(lib().vch2d_pgd_iteration_ckpt(self.h, lv, a.host(t_hist), a.host(dt_steps), a.host(x), a.host(y), a.inp(u, full),
                                          a.inp(ck[0], cshape), a.inp(ck[1], cshape), a.inp(ck[2], cshape), int(stride),
                                          a.inp(phiQ, full if phiQ is not None else None), a.inp(phiT, s if phiT is not None else None),
                                          C.c_double(b1), C.c_double(b2), C.c_double(b3), C.c_double(kappa_sp), C.c_double(u_min),
                                          C.c_double(u_max), C.c_double(alpha), pun, pcp, pcm, pcw, pr,
                                          J.ctypes.data_as(C.c_void_p), red.ctypes.data_as(C.c_void_p), C.byref(st),
                                          _mem_of(u, ck[0], ck[1], ck[2], phiQ, phiT)))
    self.last_stats = st.as_dict()
    return u_out, (cp, cm, cw), J, red, self.last_stats


Ctx2D.forward_ckpt = _ctx2d_forward_ckpt
Ctx2D.pgd_iteration_ckpt = _ctx2d_pgd_iteration_ckpt


def slab_partition(N, nranks):
    """Row ranges [(row0, rows)] of the slab decomposition of an (N+1, N+1) field: N // nranks rows per rank, the last
    rank one more (the same rule the library applies, csrc/vch2d.cu create_ctx)."""
    N, nranks = int(N), int(nranks)
    if nranks < 1 or N % nranks or (nranks > 1 and (N & (N - 1) or N < 32 or N > 4096 or nranks not in (2, 4, 8) or N // nranks < 8)):
        raise ValueError("slab decomposition needs N = 2^k (32..4096) and 2, 4 or 8 ranks with at least 8 rows each")
    rw = N // nranks
    return [(r * rw, rw + (1 if r == nranks - 1 else 0)) for r in range(nranks)]


class SlabCtx2D(Ctx2D):
    """One rank of a row-slab decomposition of ONE square 2D problem over 2/4/8 GPUs (include/vch_b200.h, "slab mode").

    Every array handed to the inherited methods is this rank's slab: (rows, N+1) fields, (levels, rows, N+1)
    trajectories, rows = self.shape[0] starting at global row self.row0; `x` (cost / pgd_iteration) stays the global
    abscissa vector.  All ranks must make the same calls in the same order."""

    def __init__(self, N, h, L, tau, gamma, c1, c2, kappa, delta_sep, rank, nranks, device=0):
        require_device()
        self.p = Params2D(int(N), int(N), float(h), float(h), float(L), float(L), float(tau), float(gamma),
                          float(c1), float(c2), float(kappa), float(delta_sep))
        self.h = C.c_void_p()
        self.rank, self.nranks = int(rank), int(nranks)
        _check(lib().vch2d_slab_create(C.byref(self.p), int(device), self.rank, self.nranks, C.byref(self.h)))
        r0, nr = C.c_int(0), C.c_int(0)
        _check(lib().vch2d_slab_rows(self.h, C.byref(r0), C.byref(nr)))
        self.row0, self.rows = r0.value, nr.value
        self.shape = (self.rows, int(N) + 1)
        self.last_stats = {}

    def ipc_handle(self) -> bytes:
        buf = C.create_string_buffer(64)
        _check(lib().vch2d_slab_ipc_handle(self.h, buf))
        return buf.raw

    def attach(self, handles):
        blob = b"".join(handles)
        assert len(blob) == 64 * self.nranks
        _check(lib().vch2d_slab_attach(self.h, C.c_char_p(blob)))

    def selftest(self):
        out = np.zeros(5)
        self._stream()
        _check(lib().vch2d_slab_selftest(self.h, out.ctypes.data_as(C.c_void_p)))
        return out

    @classmethod
    def create_distributed(cls, N, h, L, tau, gamma, c1, c2, kappa, delta_sep=1e-2, group=None, device=None):
        """Collective over a torch.distributed group (any backend): creates the rank's context and wires the peers."""
        import torch
        import torch.distributed as dist
        rank, nranks = dist.get_rank(group), dist.get_world_size(group)
        if device is None:
            device = torch.cuda.current_device()
        c = cls(N, h, L, tau, gamma, c1, c2, kappa, delta_sep, rank, nranks, device)
        handles = [None] * nranks
        dist.all_gather_object(handles, c.ipc_handle(), group=group)
        c.attach(handles)
        dist.barrier(group)
        return c


# ------------------------------------------------------------------------------------------------ 1D context
class Ctx1D:
    def __init__(self, N, h, Lx, tau, gamma, c1, c2, kappa, delta_sep=1e-2, device=0):
        require_device()
        self.p = Params1D(int(N), float(h), float(Lx), float(tau), float(gamma), float(c1), float(c2), float(kappa),
                          float(delta_sep))
        self.n = int(N) + 1
        self.h = C.c_void_p()
        _check(lib().vch1d_create(C.byref(self.p), int(device), C.byref(self.h)))

    def __del__(self):
        try:
            if getattr(self, "h", None) and self.h.value:
                lib().vch1d_destroy(self.h)
                self.h = C.c_void_p()
        except Exception:
            pass

    def _stream(self):
        _check(lib().vch1d_set_stream(self.h, _current_stream_ptr()))

    def launches(self):
        return int(lib().vch1d_launch_count(self.h))

    @staticmethod
    def _batched(a, tail):
        """View `a` with a leading batch axis; returns (array, batch, had_batch)."""
        if a.ndim == len(tail):
            return a.reshape((1,) + tuple(a.shape)), 1, False
        return a, int(a.shape[0]), True

    def residual(self, phi_new, phi_old, mu_new, mu_old, w_new, w_old, dt):
        a = _Args(); self._stream()
        x, B, had = self._batched(phi_new, (self.n,))
        sh = (B, self.n)
        rs = lambda v: v.reshape(sh)
        rp, prp = a.out(phi_new, sh); rm, prm = a.out(phi_new, sh)
        _check(lib().vch1d_residual(self.h, B, a.inp(rs(phi_new)), a.inp(rs(phi_old)), a.inp(rs(mu_new)), a.inp(rs(mu_old)),
                                    a.inp(rs(w_new)), a.inp(rs(w_old)), C.c_double(dt), prp, prm,
                                    _mem_of(phi_new, phi_old, mu_new, mu_old, w_new, w_old)))
        return (rp, rm) if had else (rp[0], rm[0])

    def initialize_mu(self, phi, w):
        a = _Args(); self._stream()
        x, B, had = self._batched(phi, (self.n,))
        out, po = a.out(phi, (B, self.n))
        _check(lib().vch1d_initialize_mu(self.h, B, a.inp(phi.reshape(B, self.n)), a.inp(w.reshape(B, self.n)), po, _mem_of(phi, w)))
        return out if had else out[0]

    def newton(self, phi_old, mu_old, w_old, w_new, dt, hist_cap=64):
        a = _Args(); self._stream()
        x, B, had = self._batched(phi_old, (self.n,))
        sh = (B, self.n)
        rs = lambda v: v.reshape(sh)
        pn, ppn = a.out(phi_old, sh); mn, pmn = a.out(phi_old, sh)
        hist = np.zeros((B, hist_cap)); nh = np.zeros(B, dtype=np.int32); status = np.zeros(B, dtype=np.int32)
        _check(lib().vch1d_newton(self.h, B, a.inp(rs(phi_old)), a.inp(rs(mu_old)), a.inp(rs(w_old)), a.inp(rs(w_new)),
                                  C.c_double(dt), ppn, pmn, hist.ctypes.data_as(C.c_void_p), hist_cap,
                                  nh.ctypes.data_as(C.c_void_p), status.ctypes.data_as(C.c_void_p),
                                  _mem_of(phi_old, mu_old, w_old, w_new)))
        hists = [[float(v) for v in hist[b, :min(int(nh[b]), hist_cap)]] for b in range(B)]
        return (pn, mn, hists) if had else (pn[0], mn[0], hists[0])

    def forward(self, phi0, u, dt_steps, want_mu=False, want_w=False):
        """phi0 (N+1,) or (B, N+1); u (rows, N+1) / (B, rows, N+1) / None.  Returns (phi_hist, mu_hist|None, w_hist|None)
        with phi_hist (M+2, N+1) or (B, M+2, N+1) — level 0 stored twice like the reference."""
        a = _Args(); self._stream()
        p0, B, had = self._batched(phi0, (self.n,))
        dts = np.ascontiguousarray(dt_steps, dtype=np.float64)
        M = int(dts.shape[0])
        rows = 0
        if u is not None:
            ub = u if had else u.reshape((1,) + tuple(u.shape))
            if ub.ndim != 3 or ub.shape[0] != B or ub.shape[2] != self.n:
                raise ValueError(f"control_input must have shape (rows, {self.n})")
            rows = int(ub.shape[1])
            if M > rows:     # the reference indexes control_input[step] for every step (Forward_solver.py:347-353)
                raise IndexError(f"index {rows} is out of bounds for axis 0 with size {rows}")
        else:
            ub = None
        hist, ph = a.out(phi0, (B, M + 2, self.n))
        mh, pm = a.out(phi0, (B, M, self.n)) if want_mu else (None, C.c_void_p(None))
        wh, pw = a.out(phi0, (B, M, self.n)) if want_w else (None, C.c_void_p(None))
        status = np.zeros(B, dtype=np.int32); st = Stats()
        _check(lib().vch1d_forward(self.h, B, a.inp(p0), a.inp(ub), rows, M, a.host(dts), ph, pm, pw,
                                   status.ctypes.data_as(C.c_void_p), C.byref(st), _mem_of(phi0, u)))
        if had:
            return hist, mh, wh
        return hist[0], (mh[0] if mh is not None else None), (wh[0] if wh is not None else None)

    def adjoint(self, phi_hist, t_hist, b1, b2, phiQ=None, phiT=None):
        a = _Args(); self._stream()
        F, B, had = self._batched(phi_hist, (0, self.n))
        lv = int(F.shape[1])
        full = (B, lv, self.n)
        b1v = np.ascontiguousarray(np.broadcast_to(np.asarray(b1, dtype=np.float64), (B,)))
        b2v = np.ascontiguousarray(np.broadcast_to(np.asarray(b2, dtype=np.float64), (B,)))
        Q = None if phiQ is None else phiQ.reshape(full)
        T = None if phiT is None else phiT.reshape((B, self.n))
        p, pp = a.out(phi_hist, full); q, pq = a.out(phi_hist, full); r, pr = a.out(phi_hist, full)
        _check(lib().vch1d_adjoint(self.h, B, a.inp(F), lv, a.host(t_hist), a.host(b1v), a.host(b2v), a.inp(Q), a.inp(T),
                                   pp, pq, pr, _mem_of(phi_hist, phiQ, phiT)))
        return (p, q, r) if had else (p[0], q[0], r[0])

    def cost(self, phi_hist, u, phiQ, phiT, x, t_hist, b1, b2, b3, kappa_sp):
        a = _Args(); self._stream()
        F, B, had = self._batched(phi_hist, (0, self.n))
        lv = int(F.shape[1])
        full = (B, lv, self.n)
        wts = np.ascontiguousarray(np.stack([np.broadcast_to(np.asarray(v, dtype=np.float64), (B,)) for v in (b1, b2, b3, kappa_sp)], axis=1))
        J = np.zeros((B, 5))
        _check(lib().vch1d_cost(self.h, B, a.inp(F), a.inp(None if u is None else u.reshape(full)),
                                a.inp(None if phiQ is None else phiQ.reshape(full)),
                                a.inp(None if phiT is None else phiT.reshape((B, self.n))), lv, a.host(x), a.host(t_hist),
                                a.host(wts), J.ctypes.data_as(C.c_void_p), _mem_of(phi_hist, u, phiQ, phiT)))
        return J if had else J[0]

    def grad_prox(self, u, r, b3, alpha, kappa_sp, u_min, u_max):
        """Per-problem parameters (scalars broadcast).  u, r: (B, levels, N+1).  Returns (u_new, red (B,4))."""
        a = _Args(); self._stream()
        B = int(u.shape[0])
        per = int(np.prod(u.shape[1:]))
        par = np.ascontiguousarray(np.stack([np.broadcast_to(np.asarray(v, dtype=np.float64), (B,))
                                             for v in (b3, alpha, kappa_sp, u_min, u_max, 0.0)], axis=1))
        un, pun = a.out(u, tuple(u.shape))
        red = np.zeros((B, 4))
        _check(lib().vch1d_grad_prox(self.h, B, C.c_longlong(per), a.inp(u), a.inp(r, tuple(u.shape)), a.host(par), pun,
                                     red.ctypes.data_as(C.c_void_p), _mem_of(u, r)))
        return un, red


# ------------------------------------------------------------------------------------------------ context cache
# Contexts handed to the drop-in modules.  They are shared, so a lookup always returns one with the DEFAULT solver settings
# (a caller that changed tolerances, the stream budget or profiling on a cached context must not leak that into later
# run_main_simulation / run_backward calls with the same physics), and eviction is least-recently-used, one entry at a time
# (parameter sweeps such as the reference's convergence-order tests keep their working set instead of rebuilding DCT tables
# and CUDA graphs after every 16th configuration).
from collections import OrderedDict
import threading

_ctx_cache: "OrderedDict" = OrderedDict()
_CTX_CACHE_MAX = 16
_ctx_lock = threading.Lock()
_tls = threading.local()


def set_ctx_slot(slot: int = 0) -> None:
    """Context slot of the calling thread.  The drop-in modules look their contexts up per (parameters, slot), so worker
    threads with different slots drive DIFFERENT contexts (own stream, own work vectors) and their kernels overlap on the
    GPU — how independent 2D problems (line-search trials, finite-difference directions) are run concurrently."""
    _tls.slot = int(slot)


def ctx_slot() -> int:
    return int(getattr(_tls, "slot", 0))


def run_concurrent(fn, n_jobs: int, workers: int):
    """[fn(j) for j in range(n_jobs)], `workers` at a time, worker w on context slot w (ctypes releases the GIL inside
    the library calls).  workers <= 1: a plain loop on the caller's slot."""
    if workers <= 1 or n_jobs <= 1:
        return [fn(j) for j in range(n_jobs)]
    from concurrent.futures import ThreadPoolExecutor
    import queue
    free = queue.SimpleQueue()
    for w in range(min(workers, n_jobs)):
        free.put(w)

    def job(j):
        w = free.get()
        try:
            set_ctx_slot(w)
            return fn(j)
        finally:
            set_ctx_slot(0)
            free.put(w)
    with ThreadPoolExecutor(min(workers, n_jobs)) as pool:
        return list(pool.map(job, range(n_jobs)))


def _cached(key, make, reset):
    with _ctx_lock:
        c = _ctx_cache.get(key)
        fresh = c is None
        if fresh:
            while len(_ctx_cache) >= _CTX_CACHE_MAX:
                _ctx_cache.popitem(last=False)
            c = _ctx_cache[key] = make()
        else:
            _ctx_cache.move_to_end(key)
    if not fresh:
        reset(c)
    return c


def _reset2d(c):
    c.set_krylov(1e-11, 200)
    c.set_krylov_first(float(os.environ.get("VCH_KRYLOV_FIRST_RTOL", 1e-6)))
    c.set_newton(not os.environ.get("VCH_NEWTON_STRICT"))
    c.set_stream_budget(0)
    c.profile(False)


def ctx2d(Nx, Ny, hx, hy, Lx, Ly, tau, gamma, c1, c2, kappa, delta_sep=1e-2, device=0) -> Ctx2D:
    key = ("2d", int(Nx), int(Ny), float(hx), float(hy), float(Lx), float(Ly), float(tau), float(gamma), float(c1),
           float(c2), float(kappa), float(delta_sep), int(device))
    return _cached(key + (ctx_slot(),), lambda: Ctx2D(*key[1:]), _reset2d)


def ctx1d(N, h, Lx, tau, gamma, c1, c2, kappa, delta_sep=1e-2, device=0) -> Ctx1D:
    key = ("1d", int(N), float(h), float(Lx), float(tau), float(gamma), float(c1), float(c2), float(kappa),
           float(delta_sep), int(device))
    return _cached(key + (ctx_slot(),), lambda: Ctx1D(*key[1:]), lambda c: None)
