"""N>1 path on CPU: world_size 2 and 4 over gloo (SPMD, one process per rank), driving the
TEST-ONLY host build of the product sources through the C ABI's communication callbacks."""
import os
import socket
import subprocess
import sys

import pytest

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run(world, gen, n, P, vectors, hostemu):
    port = _free_port()
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "multi_rank_worker.py"), gen, str(n), str(P),
                                       vectors if isinstance(vectors, str) else ("1" if vectors else "0")], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = []
    for p in procs:
        try:
            out, _ = p.communicate(timeout=600)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        outs.append(out)
    for r, p in enumerate(procs):
        assert p.returncode == 0, "rank %d failed:\n%s" % (r, outs[r][-3000:])
    assert "MULTI_RANK_OK" in outs[0]


@pytest.mark.parametrize("world,gen,n,P,vectors", [
    (2, "goe", 300, 4, True),
    (2, "s1", 512, 2, True),
    (2, "s2", 256, 1, True),       # accurate tree, cooperative top merge
    (4, "goe", 400, 4, True),
    (4, "rand_u", 333, 8, False),  # eigenvalue-only mode
    (2, "goe", 333, 1, True),      # odd sizes: slices start at odd rows
    (4, "s2", 1000, 8, True),      # reference leaves of 125 rows, Givens-heavy
    (8, "goe", 640, 8, True),      # one rank per reference leaf, three cooperative levels
    (3, "goe", 400, 4, True),      # any rank count: 4 subtrees over 3 ranks (1 / 1 / 2)
    (3, "s1", 333, 3, True),       # reference tree of 3 leaves on 3 ranks
    (5, "goe", 777, 8, True),      # 8 subtrees over 5 ranks
    (6, "s2", 600, 1, True),       # accurate tree, 8 subtrees over 6 ranks
    (7, "rand_u", 900, 2, False),  # eigenvalue-only mode on 7 ranks
    (2, "goe", 301, 3, True),      # `-g 2 -p 3`: unequal subtrees (2 reference leaves | 1)
    (2, "goe", 300, 4, "sel"),     # -eFILE (selected-eigenvector mode) on several ranks
    (3, "s1", 500, 8, "sel"),
    (4, "s2", 256, 1, "sel"),
])
def test_sharded_solve_over_gloo(hostemu, oracle, world, gen, n, P, vectors):
    _run(world, gen, n, P, vectors, hostemu)
