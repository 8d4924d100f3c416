#!/bin/bash
# r02 call L (1 GPU): parity suite on the pack-kernel prefetch / L2-hint defaults, phase times of three workloads
O=gpurun_out/r02; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_l.txt 2>&1; echo "pytest rc $?" >> $O/pytest_l.txt; tail -3 $O/pytest_l.txt
cat > /tmp/ab.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import symmetric_eigenvalue_b200 as se
from bench import make_matrix
for mat, n in (("goe", 16384), ("wilk", 16384), ("s1", 4096), ("goe", 32768)):
    D, E = make_matrix(mat, n)
    s = se.CuppenSolver(n, ref_leaves=8, vectors=True)
    s.set_tridiagonal(D, E)
    best = None
    for it in range(5):
        s.solve(); t = s.timers()
        if it >= 2 and (best is None or t["device_s"] < best["device_s"]): best = t
    print(sys.argv[1], mat, n, "device_ms %.4f" % (best["device_s"] * 1e3), {k: round(best[k] * 1e3, 3) for k in ("pack_s", "gemm_s", "residual_s", "deflation_s", "root_finding_s", "ev_extract_s") if k in best}, "launches", best["kernel_launches"], "resid %.3e" % s.residuals().max(), flush=True)
    s.close()
PY
python /tmp/ab.py new > $O/ab_l.txt 2>&1; cat $O/ab_l.txt
