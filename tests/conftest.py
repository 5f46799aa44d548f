"""Test plumbing: marker registration, import paths for the package / oracle, golden fixtures."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "sparse-optimal-control-of-viscous-chan-hilliard-via-gradient-descent--1d-2d_b200")
for p in (PKG, os.path.join(ROOT, "oracle"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # tests/test_library_emulation.py re-runs part of the GPU suite in a subprocess against the CPU-emulation build of the
    # library (tests/emu): TEST plumbing only — the product binding has no such switch and no CPU fallback.
    emu = os.environ.get("VCH_TEST_EMU_LIB")
    if emu:
        os.environ["VCH_NO_GRAPHS"] = "1"          # the emulated runtime has no CUDA graphs
        import vch_b200_native as nat
        nat.LIB_PATH = emu


def rel(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return float(np.linalg.norm(a.ravel() - b.ravel()) / max(np.linalg.norm(b.ravel()), 1e-300))


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    return load


@pytest.fixture(scope="session")
def native():
    import vch_b200_native as nat
    if nat.device_count() < 1:
        pytest.skip("no CUDA device")
    return nat


DROPIN_NAMES = ("config", "Forward_solver", "backward_solver", "cost_and_function", "GD_1D", "second_order_conditions",
                "Forward2_solver", "backward2_solver", "cost2_and_function", "GD2_configured", "second_order_conditions_2d")


def load_dropin(dim):
    """Put <package>/Vch_control_<dim> first on sys.path (the reference's modules import each other by bare name,
    SURVEY §1) and drop bare-named modules of the other dimension from sys.modules."""
    import importlib
    for m in DROPIN_NAMES:
        sys.modules.pop(m, None)
    for d in ("1D", "2D"):
        p = os.path.join(PKG, f"Vch_control_{d}")
        while p in sys.path:
            sys.path.remove(p)
    sys.path.insert(0, os.path.join(PKG, f"Vch_control_{dim}"))
    names = [n for n in DROPIN_NAMES if ("2" in n) == (dim == "2D") or n == "config"]
    return {n: importlib.import_module(n) for n in names}
