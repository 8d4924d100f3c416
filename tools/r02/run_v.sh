#!/bin/bash
# r02 call V (1 GPU): short levels (fewer tiles than half the SMs) split into half tiles: parity suite, times with / without
O=gpurun_out/r02; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_v.txt 2>&1; echo "pytest rc $?" >> $O/pytest_v.txt; tail -3 $O/pytest_v.txt
cat > /tmp/ab.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import symmetric_eigenvalue_b200 as se
from bench import make_matrix
for mat, n, P in (("s1", 4096, 8), ("wilk", 16384, 8), ("s2", 4096, 8), ("goe", 16384, 8)):
    D, E = make_matrix(mat, n)
    s = se.CuppenSolver(n, ref_leaves=P, vectors=True)
    s.set_tridiagonal(D, E)
    best = None
    for it in range(8):
        s.solve(); t = s.timers()
        if it >= 2 and (best is None or t["device_s"] < best["device_s"]): best = t
    print(sys.argv[1], mat, n, P, "device_ms %.4f gemm_ms %.4f" % (best["device_s"] * 1e3, best["gemm_s"] * 1e3), "resid %.3e" % s.residuals().max(), flush=True)
    s.close()
PY
CUPPEN_SPLIT_TAIL=0 python /tmp/ab.py whole > $O/ab_v.txt 2>&1
python /tmp/ab.py split >> $O/ab_v.txt 2>&1
cat $O/ab_v.txt
