"""CPU tests: the C-ABI library loads and exports every symbol include/vch_b200.h declares (no compute without a GPU),
the product path fails loudly without a device, and the host-side logic of the drop-in modules."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, PKG, load_dropin


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "vch_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(vch(?:1d|2d)?_[a-z_0-9]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    import vch_b200_native as nat
    if not os.path.exists(nat.LIB_PATH):
        sys.path.insert(0, ROOT)
        import __graft_entry__ as ge
        ge.build()
    L = ctypes.CDLL(nat.LIB_PATH)
    declared = _header_symbols()
    assert len(declared) >= 30
    for s in declared:
        assert hasattr(L, s), f"{s} declared in include/vch_b200.h but not exported"
    assert sorted(nat.EXPORTS) == declared, "vch_b200_native.EXPORTS out of sync with the header"
    out = subprocess.run(["nm", "-D", "--defined-only", nat.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (vch\w+)", out))
    assert exported == set(declared), exported ^ set(declared)


def test_no_cpu_fallback():
    """Without a CUDA device every compute entry raises; nothing silently routes to NumPy or the oracle."""
    import vch_b200_native as nat
    if nat.device_count() > 0:
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CUDA device"):
        nat.Ctx2D(32, 32, 1 / 32, 1 / 32, 1, 1, 0.05, 10, 0.75, 1, 1e-4)
    with pytest.raises(RuntimeError, match="no CUDA device"):
        nat.Ctx1D(128, 1 / 128, 1, 0.05, 10, 0.75, 1, 9e-4)
    with pytest.raises(RuntimeError):
        nat.solve_w(np.zeros(4), 1e-2, 10.0, np.zeros(4), np.zeros(4))
    with pytest.raises(RuntimeError):
        nat.grad_prox(np.zeros(4), np.zeros(4), 0.0, 1.0, 0.1, -1.0, 1.0)
    src = "".join(open(os.path.join(PKG, d, f)).read() for d in ("", "Vch_control_1D", "Vch_control_2D")
                  for f in os.listdir(os.path.join(PKG, d)) if f.endswith(".py"))
    assert "vch_oracle" not in src and "import oracle" not in src, "product code must never import the oracle"


@pytest.mark.parametrize("dim", ["1D", "2D"])
def test_config_models_match_reference_defaults_and_validators(dim, tmp_path, capsys):
    m = load_dropin(dim) if False else None
    # import only the config module (the solver modules need the GPU library but not a GPU to import)
    mods = load_dropin(dim)
    C = mods["config"]
    f, o = C.ForwardSolverConfig(), C.OptimizationConfig()
    if dim == "2D":
        assert (f.Nx, f.Ny, f.Lx, f.Ly, f.T, f.dt_initial, f.tau, f.gamma, f.c1, f.c2) == (128, 128, 1.0, 1.0, 1.0, 1e-2, 0.05, 10.0, 0.75, 1.0)
        assert f.kappa == 0.01 ** 2 and (o.b1, o.b2, o.b3, o.kappa_sparsity, o.alpha_max, o.max_iter) == (5.0, 10.0, 1e-4, 1e-4, 50.0, 500)
        with pytest.raises(Exception):
            C.ForwardSolverConfig(Nx=10)
    else:
        assert (f.N, f.Lx, f.T, f.dt_initial, f.tau, f.gamma, f.c1, f.c2) == (128, 1.0, 1.0, 1e-2, 0.05, 10.0, 0.75, 1.0)
        assert f.kappa == 0.03 ** 2 and (o.b1, o.b2, o.b3, o.kappa_sparsity, o.alpha_max, o.max_iter) == (0.3, 13.0, 0.0019, 9e-5, 100.0, 1000)
        with pytest.raises(Exception):
            C.ForwardSolverConfig(N=10)
    assert (o.u_min, o.u_max) == (-1.0, 1.0)
    for bad in (dict(c1=1.0, c2=0.9), dict(T=-1.0), dict(gamma=0.0)):
        with pytest.raises(Exception):
            C.ForwardSolverConfig(**bad)
    with pytest.raises(Exception):
        C.OptimizationConfig(u_min=0.5, u_max=0.5)
    assert set(f.dict().keys()) == set(type(f).model_fields.keys())              # tests use cfg.dict()
    path = str(tmp_path / "cfg.json")
    C.save_params(f, o, 17, path)
    back = C.load_params(path)
    assert back.last_run_iterations == 17 and back.forward_solver == f and back.optimization == o
    assert C.load_params(str(tmp_path / "missing.json")).last_run_iterations == 0
    capsys.readouterr()


def test_host_side_helpers_2d():
    mods = load_dropin("2D")
    F, G = mods["Forward2_solver"], mods["GD2_configured"]
    import vch_oracle as O
    dts, t = F._time_grid(1.0, 1e-2)
    ref = O.forward_2d.__globals__  # noqa: F841  (oracle only used as a checker of host logic)
    tt, ts = 0.0, [0.0]
    while tt < 1.0 - 1e-10:
        d = min(1e-2, 1.0 - tt); tt += d; ts.append(min(tt, 1.0))
    assert len(dts) == 100 and np.array_equal(t, np.array(ts))
    dts2, t2 = F._time_grid(0.25, 0.1)
    assert len(dts2) == 3 and abs(dts2[-1] - 0.05) < 1e-15 and t2[-1] == 0.25
    assert np.array_equal(F.init_phi_random(20, 12, 1e-2, amp=0.1, seed=42), O.init_phi_2d(20, 12))
    L = F.laplacian_matrix_neumann(12, 20, 1 / 12, 1.5 / 20)
    assert abs(L - O.neumann_2d(12, 20, 1 / 12, 1.5 / 20)).max() == 0 and F._spacing((2.0 * L * 0.5).tocsr(), 12, 20) == pytest.approx((1 / 12, 1.5 / 20))
    x, y, th = np.linspace(0, 1, 13), np.linspace(0, 1.5, 21), np.linspace(0, 0.5, 6)
    p0 = np.random.default_rng(0).standard_normal((13, 21))
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        pT, pQ = G.build_targets(x, y, th, p0, 1.0, 1.5, 0.5, False, 1, 1)
        pT2, pQ2 = G.build_targets(x, y, th, p0, 1.0, 1.5, 0.5, False, 2, 2)
    oT, oQ = O.targets_2d(x, y, th, p0, 1.0, 1.5, 0.5)
    assert np.array_equal(pT, oT) and np.array_equal(pQ, oQ)
    assert set(np.unique(pT2)) == {-1.0, 1.0} and not pQ2.any()
    import vch_b200_native as nat
    if nat.device_count() == 0:      # free_energy is a device reduction (vch_free_energy): no CPU fallback, fails loudly
        with pytest.raises(RuntimeError):
            F.free_energy(0 * p0, 1e-4, 0.75, 1.0, 1 / 12, 1.5 / 20)
    else:
        assert F.free_energy(0 * p0, 1e-4, 0.75, 1.0, 1 / 12, 1.5 / 20) == 0.0
    with pytest.raises(ValueError):  # an operator that is not the Neumann Laplacian of this grid is rejected, not replaced
        F._spacing((3.0 * L).tocsr() + 1e-3 * __import__("scipy.sparse", fromlist=["eye"]).eye(L.shape[0], format="csr"), 12, 20)
    with pytest.raises(ValueError):
        F._spacing(F.laplacian_matrix_neumann(6, 5, 0.1, 0.1), 12, 20)
    assert F.instability_report.__call__ and len(F.regularized_log(np.array([2.0, -2.0, 0.0]), 1e-2)) == 3


def test_host_side_helpers_1d():
    mods = load_dropin("1D")
    F, G = mods["Forward_solver"], mods["GD_1D"]
    import vch_oracle as O
    dts, t = F._time_grid(1.0, 1e-2)
    assert len(dts) == 100 and len(t) == 102 and t[0] == t[1] == 0.0
    assert np.array_equal(F.init_phi_random(128, 1e-2, amp=0.01, seed=42), O.init_phi_1d(128))
    x = np.linspace(0, 1, 129)
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        for ct in (1, 2, 3):
            pT, pQ = G.build_targets_1d(x, t, 0.01 * np.cos(x), 1.0, 1.0, False, ct, 1)
            oT, oQ = O.targets_1d(x, t, 0.01 * np.cos(x), 1.0, choice_t=ct)
            assert np.array_equal(pT, oT) and np.array_equal(pQ, oQ)
    ens = G.make_ensemble(16)
    assert ens["phi_Q"].shape == (16, 102, 129) and ens["b1"].min() >= 0.1 and set(ens["choice_t"]) <= {1, 2, 3}
    parts = [G.shard_range(1024, r, 8) for r in range(8)]
    assert parts[0][0] == 0 and parts[-1][1] == 1024 and all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
    sizes = [b - a for a, b in (G.shard_range(10, r, 4) for r in range(4))]
    assert sizes == [3, 3, 2, 2]


def test_context_slots_and_concurrent_runner():
    """Worker threads of vch_b200_native.run_concurrent drive distinct context slots (so that batched 2D trials use distinct library
    contexts), results come back in job order, and the caller's slot is untouched."""
    import threading
    import vch_b200_native as nat
    seen = []
    lock = threading.Lock()

    def job(j):
        with lock:
            seen.append((j, nat.ctx_slot(), threading.get_ident()))
        return j * j
    assert nat.run_concurrent(job, 7, 3) == [j * j for j in range(7)]
    assert nat.ctx_slot() == 0
    slots = {s for _, s, _ in seen}
    assert slots <= {0, 1, 2} and len(seen) == 7
    # two jobs that overlap in time never share a slot: with as many workers as jobs every job has its own
    seen.clear()
    barrier = threading.Barrier(3)

    def job2(j):
        barrier.wait(timeout=10)
        return nat.ctx_slot()
    assert sorted(nat.run_concurrent(job2, 3, 3)) == [0, 1, 2]
    # workers <= 1: a plain loop on the caller's slot
    assert nat.run_concurrent(lambda j: nat.ctx_slot(), 3, 1) == [0, 0, 0]


def test_bench_and_entry_scripts_parse_and_expose_the_contract():
    """bench.py and __graft_entry__.py only ever run on the GPU box; a typo there would cost the round's measurement.  Here: both
    compile, bench.py's command line has the driver's flags (--gpus / --steps / --warmup / --impl) and the workload switches the docs
    name, the B200 arm refuses to run without a device instead of falling back, and the fit used by the CPU arm recovers a power law."""
    import importlib.util
    import subprocess
    for name in ("bench.py", "__graft_entry__.py"):
        with open(os.path.join(ROOT, name)) as fh:
            compile(fh.read(), name, "exec")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--help"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr[-2000:]
    for flag in ("--gpus", "--steps", "--warmup", "--impl", "--workload", "--horizon", "--rank-seeds", "--no-e2e", "--no-slab",
                 "--no-ensemble", "--no-concurrent", "--no-parity", "--no-cpu", "--profile-steps"):
        assert flag in out.stdout, flag
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    p, c = bench._fit_power([(1e4, 2.0 * 1e4 ** 1.5), (4e4, 2.0 * 4e4 ** 1.5), (9e4, 2.0 * 9e4 ** 1.5)])
    assert abs(p - 1.5) < 1e-9 and abs(c - 2.0) < 1e-6
    assert bench.METRIC.startswith("PGD iters/s") and bench.UNIT == "it/s"
    import torch
    if not torch.cuda.is_available():      # no CPU fallback: the B200 arm must fail loudly without a device
        run = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0", "--horizon", "2", "--n", "32"],
                             capture_output=True, text=True, timeout=300)
        assert run.returncode != 0 and not any(l.startswith('{"metric"') for l in run.stdout.splitlines())
