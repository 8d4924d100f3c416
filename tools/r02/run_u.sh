#!/bin/bash
# r02 call U (1 GPU): half tiles for an under-filled last GEMM wave: parity suite, GEMM time with / without on workloads whose
# upper levels end in a short wave (s2 n=4096: 512 tiles = 3 waves + 68; GOE n=4096 P=4: 464 = 3 waves + 20)
O=gpurun_out/r02; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_u.txt 2>&1; echo "pytest rc $?" >> $O/pytest_u.txt; tail -3 $O/pytest_u.txt
cat > /tmp/ab.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import symmetric_eigenvalue_b200 as se
from bench import make_matrix
for mat, n, P in (("s2", 4096, 8), ("goe", 4096, 4), ("goe", 16384, 8), ("goe", 32768, 8)):
    D, E = make_matrix(mat, n)
    s = se.CuppenSolver(n, ref_leaves=P, vectors=True)
    s.set_tridiagonal(D, E)
    best = None
    for it in range(7):
        s.solve(); t = s.timers()
        if it >= 2 and (best is None or t["gemm_s"] < best["gemm_s"]): best = t
    print(sys.argv[1], mat, n, P, "device_ms %.4f gemm_ms %.4f" % (best["device_s"] * 1e3, best["gemm_s"] * 1e3), "resid %.3e" % s.residuals().max(), flush=True)
    s.close()
PY
CUPPEN_SPLIT_TAIL=0 python /tmp/ab.py whole > $O/ab_u.txt 2>&1
python /tmp/ab.py split >> $O/ab_u.txt 2>&1
cat $O/ab_u.txt
