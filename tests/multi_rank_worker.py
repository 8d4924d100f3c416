"""Worker of tests/test_multi_rank.py: one process per rank, torch.distributed (gloo) carries the
vector exchanges of the TEST-ONLY host build through the C ABI's communication callbacks -- the
same code path (row-block sharding, cooperative merges, root-range split, halo + all-reduce for
the residuals) that NCCL drives on the GPUs."""
import ctypes
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import symmetric_eigenvalue_b200 as se  # noqa: E402
from symmetric_eigenvalue_b200 import api  # noqa: E402
import oracle  # noqa: E402


def as_tensor(ptr, nbytes):
    arr = np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_uint8)), shape=(nbytes,))
    return torch.from_numpy(arr)


def make_callbacks(rank, world):
    def group_bcast(user, buf, nbytes, root, lo, cnt):
        try:
            t = as_tensor(buf, nbytes)
            if rank == root:
                for r in range(lo, lo + cnt):
                    if r != root:
                        dist.send(t, r)
            else:
                dist.recv(t, root)
            return 0
        except Exception as e:      # pragma: no cover
            print("group_bcast failed:", e, flush=True)
            return 1

    def allreduce(user, buf, count):
        t = torch.from_numpy(np.ctypeslib.as_array(buf, shape=(count,)))
        dist.all_reduce(t)
        return 0

    def allgather(user, send, recv, nbytes):
        s = as_tensor(send, nbytes)
        out = [torch.empty(nbytes, dtype=torch.uint8) for _ in range(world)]
        dist.all_gather(out, s.clone())
        as_tensor(recv, nbytes * world).copy_(torch.cat(out))
        return 0

    def allreduce_i32(user, buf, count):
        t = torch.from_numpy(np.ctypeslib.as_array(buf, shape=(count,)))
        dist.all_reduce(t)
        return 0

    def alltoallv(user, send, sbytes, recv, rbytes):
        try:
            reqs = []
            keep = []
            for k in range(1, world):
                dst, src = (rank + k) % world, (rank - k) % world
                if sbytes[dst]:
                    t = as_tensor(send[dst], sbytes[dst]).clone()
                    keep.append(t)
                    reqs.append(dist.isend(t, dst))
                if rbytes[src]:
                    reqs.append(dist.irecv(as_tensor(recv[src], rbytes[src]), src))
            for r in reqs:
                r.wait()
            return 0
        except Exception as e:      # pragma: no cover
            print("alltoallv failed:", e, flush=True)
            return 1

    cb = api._Callbacks()
    cb.user = None
    cb.group_bcast = api._BCAST_FN(group_bcast)
    cb.allreduce_sum_f64 = api._ALLRED_FN(allreduce)
    cb.allgather = api._ALLGATHER_FN(allgather)
    cb.allreduce_sum_i32 = api._ALLRED_I32_FN(allreduce_i32)
    cb.alltoallv = api._ALLTOALLV_FN(alltoallv)
    return cb


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    gen, n, P, vectors = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4] == "1"
    select_mode = sys.argv[4] == "sel"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lib = api._declare(ctypes.CDLL(os.path.join(ROOT, "tests", "host", "_build", "libcuppen_hostemu.so")))
    D, E = {"goe": oracle.goe, "s1": lambda k: oracle.scheme(1, k), "s2": lambda k: oracle.scheme(2, k),
            "rand_u": oracle.rand_u}[gen](n)
    cb = make_callbacks(rank, world)
    if select_mode:
        # -eFILE on several ranks: replicated eigenvalue-only solve, the selected vectors dealt to the ranks and gathered
        rng = np.random.default_rng(n)
        sel = [0, n - 1] + rng.integers(0, n, size=11).tolist()       # 13 vectors: uneven over 2 / 3 / 4 ranks, duplicates allowed
        s = se.CuppenSolver(n, ref_leaves=P, rank=rank, world=world, callbacks=cb, lib=lib, select=True)
        s.set_tridiagonal(D, E)
        s.select(sel)
        s.solve()
        one = se.cuppens(D, E, ref_leaves=P, lib=lib, select=sel)      # the same on one rank
        assert np.array_equal(s.eigenvalues(), one["lam"])
        assert np.array_equal(s.selected_eigenvectors(), one["V"]), "gathered vectors differ from the one-rank run"
        assert np.array_equal(s.residuals(sel), one["resid"])
        s.select(sel[:1])                                              # fewer vectors than ranks
        s.solve()
        assert np.array_equal(s.selected_eigenvectors(), one["V"][:, :1])
        s.close()
        dist.barrier()
        if rank == 0:
            print("MULTI_RANK_OK", flush=True)
        dist.destroy_process_group()
        return
    s = se.CuppenSolver(n, ref_leaves=P, vectors=vectors, rank=rank, world=world, callbacks=cb, lib=lib)
    s.set_tridiagonal(D, E)
    s.solve()
    lam = s.eigenvalues()
    r0, rows = s.local_rows()
    rowmap = s.local_row_map() if vectors else None
    # every rank must hold the same eigenvalues
    t = torch.from_numpy(lam.copy()); ref = t.clone()
    dist.broadcast(ref, 0)
    assert torch.equal(t, ref), "eigenvalues differ between ranks"
    single = se.cuppens(D, E, ref_leaves=P, vectors=vectors, lib=lib)       # one-rank run of the same build
    assert np.abs(lam - single["lam"]).max() <= 1e-14 * (np.abs(D).max() + 2 * np.abs(E).max())
    mine = sorted((m.m, m.offset, m.zdefl, m.givens) for m in s.merge_stats() if m.mode == 1)
    # a rank only sees the merges that touch its rows; all of them must match the one-rank run
    allst = sorted((m.m, m.offset, m.zdefl, m.givens) for m in single["stats"] if m.mode == 1)
    assert set(mine) <= set(allst), (mine, allst)
    if vectors:
        V = s.eigenvectors()
        res = s.residuals()
        assert V.shape == (rows, n) and rowmap.size == rows
        # all ranks together hold every row exactly once
        cnt = torch.zeros(n, dtype=torch.int64); cnt[rowmap] += 1
        dist.all_reduce(cnt)
        assert bool((cnt == 1).all()), "row ownership is not a partition"
        assert np.abs(V - single["V"][rowmap]).max() <= 1e-13
        assert np.abs(res - single["resid"]).max() <= 1e-12 + 1e-6 * single["resid"].max()
        # -c and -v on several ranks: Gram check over the gathered row slices, eigenvector file written by rank 0
        dev, _ = s.orthogonality()
        want = np.abs(single["V"].T @ single["V"] - np.eye(n)).max()
        assert abs(dev - want) <= 1e-13, (dev, want)
        import tempfile
        path = os.path.join(tempfile.gettempdir(), "cuppen_mr_%d_%d.bin" % (os.getppid(), n)) if rank == 0 else None
        s.write_eigenvectors(path)
        if rank == 0:
            ranks, lamf, Vf = se.read_eigenvector_file(path)
            os.unlink(path)
            assert np.array_equal(lamf, lam) and Vf.shape == (n, n)
            assert np.abs(Vf - single["V"]).max() <= 1e-13
    o = oracle.solve(D, E, P)
    assert np.abs(lam - o["lam"]).max() <= 1e-12 * (np.abs(D).max() + 2 * np.abs(E).max())
    s.close()
    dist.barrier()
    if rank == 0:
        print("MULTI_RANK_OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
