"""Times the unmodified reference (oracle/_ref/cuppens_ref) on GOE n, P=8: eigenvalue phase and one eigenvector."""
import os, sys, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import oracle
n = int(sys.argv[1]); P = int(sys.argv[2]); T = int(sys.argv[3])
D, E = oracle.goe(n)
with tempfile.TemporaryDirectory() as td:
    mtx = os.path.join(td, "in.mtx"); out = os.path.join(td, "o.txt"); ev = os.path.join(td, "ev.txt")
    oracle.write_mtx(mtx, D, E)
    t0 = time.time(); r = oracle.run_reference(["-i", mtx, out], P=P, threads=T, timeout=7200, stats=False); t1 = time.time()
    print("eigenvalue-only wall", t1 - t0, flush=True); print(r["stdout"][-600:], flush=True)
    open(ev, "w").write("%d\n" % (n // 2))
    t0 = time.time(); r = oracle.run_reference(["-i", mtx, "-e" + ev, out], P=P, threads=T, timeout=7200, stats=False); t1 = time.time()
    print("one-vector wall", t1 - t0, flush=True); print(r["stdout"][-600:], flush=True)
