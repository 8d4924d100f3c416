"""N>1 on hardware: the shipped CUDA library over its peer-memory / NCCL path on 2 / 3 / 4 / 8 B200s of one box (one process per GPU,
torch.distributed.run), against the reference's goldens, the one-GPU run and the oracle.  Skipped for world sizes the
box does not have; tests/test_multi_rank.py covers the host-side logic of the same path over gloo on CPU."""
import os
import socket
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run_world(world, cases=(), env=None):
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    if torch.cuda.device_count() < world:
        pytest.skip("%d GPUs needed, %d visible" % (world, torch.cuda.device_count()))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "gpu_multi_rank_worker.py")] + list(cases)
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900, cwd=ROOT,
                       env=dict(os.environ, **(env or {})))
    assert p.returncode == 0 and "GPU_MULTI_RANK_OK" in p.stdout, p.stdout[-6000:]
    return p.stdout


@pytest.mark.parametrize("world", [2, 3, 4, 8])
def test_sharded_solve_over_nccl(product_lib, oracle, world):
    """Peer-memory back end (the default): every case of the worker."""
    out = _run_world(world)
    assert "backend=2" in out, out[-2000:]


def test_sharded_solve_collective_fallback(product_lib, oracle):
    """CUPPEN_P2P=0: the NCCL-collective back end of round 1 stays a working fallback (two ranks, three cases)."""
    out = _run_world(2, cases=("goe_n4096_p8", "s1_n1000_p8", "wilk"), env={"CUPPEN_P2P": "0"})
    assert "backend=1" in out, out[-2000:]
