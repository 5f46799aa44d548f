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


def rel(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300))


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
