"""Worker of tests/test_gpu_multi_rank.py: one process per B200, launched by torch.distributed.run.  Drives the
SHIPPED library (libcuppen_b200.so) over its NCCL path -- row redistribution into slice layout, cooperative merges
with the root-range split, GEMMs on local row slices (odd offsets included), halo exchange + all-reduce of the
residuals -- and checks every case against the reference's golden output, the one-GPU run of the same library and the
CPU oracle.  torch.distributed (NCCL) is only the test's own plumbing (unique-id broadcast, gathers)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import symmetric_eigenvalue_b200 as se  # noqa: E402
from symmetric_eigenvalue_b200 import api  # noqa: E402
import oracle  # noqa: E402
from conftest import load_golden, norm_T  # noqa: E402

# (kind, name-or-generator, n, P, vectors)
CASES = [
    ("golden", "s1_n4096_p8", 0, 0, True),        # BASELINE configs[1]
    ("golden", "goe_n4096_p8", 0, 0, True),       # GEMM-heavy
    ("golden", "s1_n1000_p8", 0, 0, True),        # reference leaves of 125 rows: odd row offsets (cp.async GEMM path)
    ("golden", "s2_n4096_p8", 0, 0, True),        # Givens-heavy (2051 rotations at the top merge)
    ("golden", "randu_n16384_p8", 0, 0, False),   # eigenvalue-only mode at a BASELINE size
    ("gen", "goe", 2049, 1, True),                # accurate rule, odd size
    ("gen", "wilk", 1536, 4, True),
    ("gen", "goe", 777, 2, True),                 # small and ragged: slices of a few rows
]


def gen(name, n):
    return {"goe": oracle.goe, "rand_u": oracle.rand_u, "wilk": lambda k: oracle.wilkinson(k, norm=64.0),
            "s1": lambda k: oracle.scheme(1, k), "s2": lambda k: oracle.scheme(2, k)}[name](n)


def nccl_id(rank):
    t = torch.zeros(api.NCCL_ID_BYTES, dtype=torch.uint8, device="cuda")
    if rank == 0:
        t.copy_(torch.frombuffer(bytearray(se.nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(t, 0)
    return bytes(t.cpu().numpy().tobytes())


def gather(obj, world):
    out = [None] * world
    dist.all_gather_object(out, obj)
    return out


def run_case(rank, world, local, case):
    kind, name, n, P, vectors = case
    g = None
    if kind == "golden":
        g = load_golden(name)
        D, E, P = g["D"], g["E"], g["P"]
        n = len(D)
    else:
        D, E = gen(name, n)
    nT = norm_T(D, E)
    try:
        s = se.CuppenSolver(n, ref_leaves=P, vectors=vectors, device=local, rank=rank, world=world, nccl_id=nccl_id(rank))
    except se.CuppenError as e:
        # a tree without a level of `world` subtrees is refused at create on every rank alike
        assert "no level with" in str(e) or "too small" in str(e), e
        return "refused (%s)" % str(e)[:60]
    s.set_tridiagonal(D, E)
    for rep in range(2):          # the second solve re-uses every buffer (stale data of the first must not leak)
        s.solve()
    lam = s.eigenvalues()
    t = torch.from_numpy(lam.copy()).cuda(); ref = t.clone()
    dist.broadcast(ref, 0)
    assert torch.equal(t, ref), "eigenvalues differ between ranks"
    single = se.cuppens(D, E, ref_leaves=P, vectors=vectors, device=local)      # one-GPU run on this rank's device
    assert np.abs(lam - single["lam"]).max() <= 1e-13 * nT, np.abs(lam - single["lam"]).max()
    mine = [(m.m, m.offset, m.zdefl, m.givens) for m in s.merge_stats() if m.mode == 1]
    allst = sorted((m.m, m.offset, m.zdefl, m.givens) for m in single["stats"] if m.mode == 1)
    union = sorted(set(x for part in gather(mine, world) for x in part))
    assert union == allst, (union, allst)
    if g is not None:
        assert np.abs(lam - g["lam"]).max() <= 1e-12 * nT, (name, np.abs(lam - g["lam"]).max())
        assert union == [tuple(int(x) for x in r) for r in g["merges"].tolist()]
    else:
        o = oracle.solve(D, E, P)
        assert np.abs(lam - o["lam"]).max() <= 1e-12 * nT
    if vectors:
        V = s.eigenvectors()
        res = s.residuals()
        r0, rows = s.local_rows()
        rowmap = s.local_row_map()
        assert V.shape == (rows, n) and rowmap.size == rows
        cnt = torch.zeros(n, dtype=torch.int64, device="cuda"); cnt[torch.from_numpy(rowmap.astype(np.int64)).cuda()] += 1
        dist.all_reduce(cnt)
        assert bool((cnt == 1).all()), "row ownership is not a partition"
        ref_rows = single["V"][rowmap]
        dv = float(np.minimum(np.abs(V - ref_rows).max(axis=0), np.abs(V + ref_rows).max(axis=0)).max())
        assert dv <= 1e-10, dv
        assert np.abs(res - single["resid"]).max() <= 1e-12 * nT + 1e-6 * single["resid"].max()
        # assemble V on rank 0 from the ranks' row slices: residuals and orthogonality recomputed with numpy
        sel = np.unique(np.concatenate([np.linspace(0, n - 2, 16).astype(np.int32), np.linspace(0, n - 2, 16).astype(np.int32) + 1]))
        parts = gather((rowmap, s.eigenvector_columns(sel)), world)
        gram = torch.from_numpy(V.T @ V).cuda() if n <= 4096 else None
        if gram is not None:
            dist.all_reduce(gram)
            assert float((gram - torch.eye(n, dtype=torch.float64, device="cuda")).abs().max()) < 1e-12
        if rank == 0:
            Vs = np.empty((n, len(sel)))
            for rm, vl in parts:
                Vs[rm] = vl
            assert np.abs(Vs - single["V"][:, sel]).max() <= 1e-10 or np.abs(np.abs(Vs) - np.abs(single["V"][:, sel])).max() <= 1e-10
            TV = D[:, None] * Vs
            TV[1:] += E[:, None] * Vs[:-1]
            TV[:-1] += E[:, None] * Vs[1:]
            r2 = np.linalg.norm(TV - Vs * lam[sel][None, :], axis=0)
            assert np.allclose(r2, res[sel], rtol=1e-5, atol=1e-13 * nT), np.abs(r2 - res[sel]).max()
            assert np.abs(Vs.T @ Vs - np.eye(len(sel))).max() < 1e-12
    t = s.timers()
    s.close()
    return "ok launches=%d backend=%d" % (t["kernel_launches"], t["comm_mode"])


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    only = sys.argv[1:]
    for case in CASES:
        if only and case[1] not in only:
            continue
        msg = run_case(rank, world, local, case)
        dist.barrier()
        if rank == 0:
            print("CASE %s n=%s P=%s vectors=%s world=%d: %s" % (case[1], case[2] or "golden", case[3] or "golden", case[4], world, msg), flush=True)
    if rank == 0:
        print("GPU_MULTI_RANK_OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
