"""Slab mode (BASELINE config 5): ONE 2D problem decomposed over several GPUs must reproduce the
single-GPU solution of the same library (the reference cannot run these sizes in parallel at all).  Needs >= 2 GPUs; the
check itself lives in scripts/slab_check.py so that it can also be launched by hand under torchrun."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _gpus():
    import torch
    return torch.cuda.device_count()


def _run(nproc, *args, timeout=600):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(29600 + nproc), os.path.join(ROOT, "scripts", "slab_check.py"), *map(str, args)]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)


@pytest.mark.parametrize("N,M", [(128, 3), (512, 4)])
def test_two_rank_slab_matches_single_gpu(N, M):
    if _gpus() < 2:
        pytest.skip("slab mode needs at least 2 GPUs")
    out = _run(2, N, M)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "FAIL" not in out.stdout, out.stdout[-3000:]      # torchrun exits non-zero when any rank fails


def test_four_rank_slab_matches_single_gpu():
    if _gpus() < 4:
        pytest.skip("needs 4 GPUs")
    out = _run(4, 256, 3)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "FAIL" not in out.stdout, out.stdout[-3000:]


def test_slab_context_rejects_bad_shapes():
    import vch_b200_native as nat
    with pytest.raises(ValueError):         # not a power of two (VCH_E_SHAPE)
        nat.SlabCtx2D(100, 0.01, 1.0, 0.05, 10.0, 0.75, 1.0, 1e-4, 1e-2, 0, 2)
    with pytest.raises(RuntimeError):       # 3 ranks
        nat.SlabCtx2D(128, 1 / 128, 1.0, 0.05, 10.0, 0.75, 1.0, 1e-4, 1e-2, 0, 3)
    c = nat.SlabCtx2D(128, 1 / 128, 1.0, 0.05, 10.0, 0.75, 1.0, 1e-4, 1e-2, 1, 2)   # creating one rank alone is fine
    assert (c.row0, c.rows) == nat.slab_partition(128, 2)[1] == (64, 65)
    with pytest.raises(RuntimeError):       # but it cannot run collectives before the peers are attached
        c.selftest()


def test_missing_peer_raises_instead_of_hanging():
    """One rank skips a collective call: the other must come back with VCH_E_COMM once the 10 s wait limit expires."""
    if _gpus() < 2:
        pytest.skip("slab mode needs at least 2 GPUs")
    out = _run(2, 128, 2, "desync", timeout=300)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "desync: raised after" in out.stdout and "FAIL" not in out.stdout, out.stdout[-3000:]
