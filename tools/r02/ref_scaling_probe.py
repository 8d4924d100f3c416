"""Per-eigenvector back-transformation cost of the unmodified reference on leading blocks of the GOE n=16384 matrix
(P=8): validates the extrapolation used by bench.py --impl reference (exponent fitted from n/4 and n/2)."""
import os, sys, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import oracle
T = int(sys.argv[1]) if len(sys.argv) > 1 else 1
D, E = oracle.goe(16384)
for m in (2048, 4096, 8192):
    with tempfile.TemporaryDirectory() as td:
        mtx = os.path.join(td, "in.mtx"); out = os.path.join(td, "o.txt"); ev = os.path.join(td, "ev.txt")
        oracle.write_mtx(mtx, D[:m], E[:m - 1])
        open(ev, "w").write("%d\n" % (m // 2))
        t0 = time.time(); r = oracle.run_reference(["-i", mtx, "-e" + ev, out], P=8, threads=T, timeout=7200, stats=False)
        bt = [l for l in r["stdout"].splitlines() if "Required time for backtransformation" in l]
        print(m, "wall", time.time() - t0, bt, flush=True)
