"""Index logic of the peer-memory exchange (csrc/p2p.h) on the CPU: the CUPPEN_HD functors run with the ranks' symmetric
heaps emulated as buffers of one process (tests/host/p2p_host.cpp) -- halo rows of the slice layout land in the heap of the
rank that needs them, subtree vectors are replicated everywhere, residual partial sums agree on every rank -- for power-of-
two and odd rank counts and for more subtrees than ranks."""
import ctypes
import os
import subprocess

import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def p2plib():
    subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "host")], check=True, stdout=subprocess.DEVNULL)
    lib = ctypes.CDLL(os.path.join(ROOT, "tests", "host", "_build", "libp2p_host.so"))
    lib.p2p_host_check.argtypes = [ctypes.c_int] * 3
    lib.p2p_host_check.restype = ctypes.c_int
    return lib


@pytest.mark.parametrize("n,G,S", [(256, 2, 2), (300, 3, 4), (1000, 4, 4), (777, 5, 8), (512, 6, 8), (901, 7, 8), (1024, 8, 8), (257, 2, 8)])
def test_peer_memory_functors(p2plib, n, G, S):
    assert p2plib.p2p_host_check(n, G, S) == 0
