"""The WHOLE C-ABI library on the CPU: tests/emu/build_emu_lib.py compiles the product sources (csrc/*.cu, *.cuh — after two
syntactic rewrites, see there) with g++ against the CUDA execution-model emulator (tests/emu/cuda_emu.h) and a host-memory
stand-in for the CUDA runtime (tests/emu/cudart_shim.cpp).  The result exports the same C ABI and is driven through the
UNMODIFIED ctypes binding, so a selection of the GPU parity tests (`-m gpu`) runs here, without a GPU, against the same
reference goldens and oracle: host control flow (Newton loop with the forcing term, polled BiCGStab path, adjoint sweep, PGD
iteration, staging) AND device code (every kernel those tests reach, 2D and 1D).

What it cannot show: CUDA-graph execution (the emulated runtime has none: VCH_NO_GRAPHS=1), memory ordering, timing — the GPU
suite on the B200 remains the parity gate.  The emulation library is test infrastructure: nothing in the product can load it
(the binding has no override; conftest.py patches the path in the subprocess only when VCH_TEST_EMU_LIB is set)."""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))

# test ids that need neither torch.cuda nor CUDA graphs and finish in seconds under emulation
FAST = [
    "tests/test_gpu_2d.py::test_building_blocks",
    "tests/test_gpu_2d.py::test_jacobian_solve_matches_direct",
    "tests/test_gpu_2d.py::test_newton_matches_oracle",
    "tests/test_gpu_2d.py::test_forward_matches_reference_golden[g2d_rect]",
    "tests/test_gpu_2d.py::test_forward_matches_reference_golden[g2d_32]",
    "tests/test_gpu_2d.py::test_adjoint_cost_prox_match_reference_golden",
    "tests/test_gpu_2d.py::test_adjoint_step_identity_small_rect",
    "tests/test_gpu_1d.py::test_residual_and_newton_match_oracle",
    "tests/test_gpu_1d.py::test_forward_adjoint_cost_prox_match_reference_golden",
    "tests/test_gpu_1d.py::test_temporal_order_and_symmetry",
    "tests/test_gpu_edge_cases.py::test_prox_kkt_solve_w_ragged_counts",
    "tests/test_gpu_edge_cases.py::test_1d_edge_cases",
]
# VCH_EMU_FULL=1 adds the long ones (about 8 more minutes)
SLOW = [
    "tests/test_gpu_2d.py::test_pgd_iteration_matches_reference_golden[g2d_32]",
    "tests/test_gpu_2d.py::test_inexact_first_newton_solve_keeps_trajectory_and_saves_iterations",
    "tests/test_gpu_1d.py::test_ensemble_equals_single_problem_calls",
    "tests/test_gpu_2d.py::test_checkpointed_pgd_iteration_equals_fully_stored[5-True-True]",
    "tests/test_gpu_dropin_1d.py::test_batched_line_search_and_fd_directions_equal_sequential",
    "tests/test_gpu_edge_cases.py::test_zero_right_hand_side_ends_the_solve_graph",
]


@pytest.fixture(scope="module")
def emu_lib(tmp_path_factory):
    cuda_home = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    if not os.path.exists(os.path.join(cuda_home, "include", "cuda_runtime.h")):
        pytest.skip("CUDA headers not found")
    import build_emu_lib
    return build_emu_lib.build(str(tmp_path_factory.mktemp("emulib")))


def _run(emu_lib, ids):
    env = dict(os.environ, VCH_TEST_EMU_LIB=emu_lib)
    p = subprocess.run([sys.executable, "-m", "pytest", "-q", "-m", "gpu", "-p", "no:cacheprovider"] + ids, cwd=ROOT, env=env,
                       capture_output=True, text=True, timeout=3000)
    tail = p.stdout[-3000:]
    m = re.search(r"(\d+) passed", tail)
    return p.returncode, int(m.group(1)) if m else 0, tail


def test_gpu_parity_tests_pass_on_the_emulated_library(emu_lib):
    rc, passed, tail = _run(emu_lib, FAST)
    assert rc == 0 and passed >= 25 and "failed" not in tail and "skipped" not in tail, tail


# the 7-launch BiCGStab iteration (VCH_BICG6=0: the form slab mode runs; the 6-launch one is the default everywhere else and is
# what every other test here exercises) against the same goldens
BICG7 = [
    "tests/test_gpu_2d.py::test_jacobian_solve_matches_direct",
    "tests/test_gpu_2d.py::test_newton_matches_oracle",
    "tests/test_gpu_2d.py::test_forward_matches_reference_golden[g2d_rect]",
    "tests/test_gpu_2d.py::test_adjoint_cost_prox_match_reference_golden[g2d_rect]",
    "tests/test_gpu_2d.py::test_adjoint_step_identity_small_rect",
]
BICG7_SLOW = [
    "tests/test_gpu_2d.py::test_forward_matches_reference_golden[g2d_32]",
    "tests/test_gpu_2d.py::test_adjoint_cost_prox_match_reference_golden[g2d_32]",
]


def test_seven_launch_iteration_on_the_emulated_library(emu_lib):
    lib = emu_lib
    env = dict(os.environ, VCH_TEST_EMU_LIB=lib, VCH_BICG6="0")
    ids = BICG7 + (BICG7_SLOW if os.environ.get("VCH_EMU_FULL") else [])
    p = subprocess.run([sys.executable, "-m", "pytest", "-q", "-m", "gpu", "-p", "no:cacheprovider", "--timeout", "600"] + ids, cwd=ROOT,
                       env=env, capture_output=True, text=True, timeout=3000)
    tail = p.stdout[-3000:]
    m = re.search(r"(\d+) passed", tail)
    assert p.returncode == 0 and m and int(m.group(1)) >= 8 and "failed" not in tail, tail
    # the switch must actually select the form: fewer kernel launches with the 6-launch iteration for the same solver statistics
    probe = ("import os, sys, numpy as np; sys.path.insert(0, %r); sys.path.insert(0, %r); import vch_b200_native as nat, vch_oracle as O;"
             "nat.LIB_PATH = %r; P = O.Phys2D(Nx=32, Ny=32, T=0.03); c = nat.Ctx2D(32, 32, 1/32, 1/32, 1.0, 1.0, P.tau, P.gamma, P.c1, P.c2, P.kappa);"
             "c.forward(O.init_phi_2d(32, 32), None, np.full(3, 1e-2)); s = c.last_stats; print(s['kernel_launches'], s['krylov_iterations'])"
             % (os.path.join(ROOT, "sparse-optimal-control-of-viscous-chan-hilliard-via-gradient-descent--1d-2d_b200"), os.path.join(ROOT, "oracle"), lib))
    out = {}
    for flag in ("0", "1"):
        q = subprocess.run([sys.executable, "-c", probe], env=dict(os.environ, VCH_NO_GRAPHS="1", VCH_BICG6=flag), capture_output=True,
                           text=True, timeout=600, check=True)
        out[flag] = [int(v) for v in q.stdout.split()[-2:]]
    assert out["1"][1] == out["0"][1] and out["1"][0] < out["0"][0], out


@pytest.mark.skipif(not os.environ.get("VCH_EMU_FULL"), reason="long (about 8 minutes): set VCH_EMU_FULL=1")
def test_long_gpu_parity_tests_pass_on_the_emulated_library(emu_lib):
    rc, passed, tail = _run(emu_lib, SLOW)
    assert rc == 0 and passed >= 3 and "failed" not in tail, tail
