"""Generate tests/golden/*.npz by running the UNMODIFIED reference (oracle/_ref/cuppens_ref, built by
oracle/Makefile from /root/reference with the MPI/MKL shims) on the inputs listed in CASES.

Run in the build container (needs /root/reference):   python tests/golden/make_golden.py
Each file holds: D, E (input), P, lam (column 1 of the reference's output file), resid (column 2,
NaN where the reference printed none), merges (m, offset, zdefl, givens per merge from the
CUPPEN_ORACLE_STATS hook) and rhos; the `*_sel` cases also hold sel, the 1-based indices of the -eFILE run.
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
import oracle  # noqa: E402
from oracle.oracle import read_output  # noqa: E402

CASES = [
    # name, generator, P, with_vectors
    ("tinyL_p1", ("mtx", "/root/reference/tinyL.mtx"), 1, True),
    ("tinyL_p2", ("mtx", "/root/reference/tinyL.mtx"), 2, True),
    ("tinyL_p4", ("mtx", "/root/reference/tinyL.mtx"), 4, True),
    ("s1_n256_p4", ("scheme", 1, 256), 4, True),
    ("s2_n256_p4", ("scheme", 2, 256), 4, True),
    ("s2_n100_p3", ("scheme", 2, 100), 3, True),
    ("s2_n100_p5", ("scheme", 2, 100), 5, False),
    ("s2_n100_p6", ("scheme", 2, 100), 6, False),
    ("s1_n1024_p4", ("scheme", 1, 1024), 4, False),
    ("s2_n1024_p4", ("scheme", 2, 1024), 4, False),
    ("s1_n1000_p8", ("scheme", 1, 1000), 8, False),
    ("s1_n4096_p8", ("scheme", 1, 4096), 8, False),
    ("s2_n4096_p8", ("scheme", 2, 4096), 8, False),
    ("goe_n256_p4", ("goe", 256), 4, True),
    ("goe_n1024_p4", ("goe", 1024), 4, False),
    ("randu_n1024_p4", ("randu", 1024), 4, False),
    ("wilk64_n1024_p4", ("wilk", 1024), 4, False),
    # BASELINE sizes; eigenvalues only -- the reference's -e is O(n^4) here
    ("s2_n16384_p8", ("scheme", 2, 16384), 8, False),
    ("goe_n4096_p8", ("goe", 4096), 8, False),
    ("wilk64_n16384_p8", ("wilk", 16384), 8, False),
    ("randu_n16384_p8", ("randu", 16384), 8, False),
    # BASELINE configs[2] as bench.py runs it (the headline workload) and the north-star target size
    ("goe_n16384_p8", ("goe", 16384), 8, False),
    ("goe_n32768_p8", ("goe", 32768), 8, False),
    # -eFILE at BASELINE configs[1] / a GEMM-heavy input: the reference back-transforms only the listed (1-based)
    # indices (src/filehandling.c:165-239,339-345; ~4 s per vector here) -- pins the selected-eigenvector mode
    ("s1_n4096_p8_sel", ("scheme", 1, 4096), 8, [1, 1000, 2048, 4096]),
    ("goe_n4096_p8_sel", ("goe", 4096), 8, [2, 2049, 4095]),
]


def make_input(gen):
    if gen[0] == "mtx":
        D = np.full(4, 2.0); E = np.full(3, -1.0)      # tinyL.mtx is the 4x4 [-1 2 -1] matrix
        return D, E, ["-i", gen[1]]
    if gen[0] == "scheme":
        D, E = oracle.scheme(gen[1], gen[2])
        return D, E, ["-s", str(gen[1]), "-n", str(gen[2])]
    D, E = {"goe": oracle.goe, "randu": oracle.rand_u, "wilk": lambda n: oracle.wilkinson(n, norm=64.0)}[gen[0]](gen[1])
    return D, E, None


def main():
    oracle.build(ref=True)
    only = sys.argv[1:]
    for name, gen, P, vec in CASES:
        if only and name not in only:
            continue
        D, E, args = make_input(gen)
        with tempfile.TemporaryDirectory() as td:
            if args is None:
                mtx = os.path.join(td, "in.mtx")
                oracle.write_mtx(mtx, D, E)
                args = ["-i", mtx]
            out = os.path.join(td, "out.txt")
            sel = vec if isinstance(vec, list) else None
            if sel is not None:
                evf = os.path.join(td, "ev.txt")
                open(evf, "w").write("".join("%d\n" % i for i in sel))
                eopt = ["-e" + evf]
            else:
                eopt = ["-e"] if vec else []
            r = oracle.run_reference(args + eopt + [out], P=P, threads=2 if sel is not None else 1, timeout=1800)
            assert r["rc"] == 0 and "Program finished successfully!" in r["stdout"], (name, r["rc"], r["stderr"])
            lam, res = read_output(out)
        merges = np.array([[m["m"], m["off"], m["zdefl"], m["givens"]] for m in r["merges"]], dtype=np.int32).reshape(-1, 4)
        rhos = np.array([m["rho"] for m in r["merges"]])
        extra = {"sel": np.array(sel, dtype=np.int32)} if sel is not None else {}
        np.savez_compressed(os.path.join(HERE, name + ".npz"), D=D, E=E, P=P, lam=lam, resid=res, merges=merges, rhos=rhos, **extra)
        print(name, "n=%d P=%d merges=%s max resid=%s" % (len(D), P, merges[:, [0, 2, 3]].tolist()[-1:] if len(merges) else [],
                                                           np.nanmax(res) if vec else None), flush=True)


if __name__ == "__main__":
    main()
