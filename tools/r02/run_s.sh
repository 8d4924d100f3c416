#!/bin/bash
# r02 call S (1 GPU): tiled Loewner / norms kernels from m = 256 (variant) against the warp functors below 2048 (current)
O=gpurun_out/r02; mkdir -p $O
L=symmetric_eigenvalue_b200/lib/libcuppen_b200.so
cat > /tmp/ab.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import symmetric_eigenvalue_b200 as se
from bench import make_matrix
for mat, n, vec in (("goe", 16384, True), ("s1", 4096, True), ("wilk", 16384, True), ("goe", 65536, False), ("s1", 4096, False)):
    D, E = make_matrix(mat, n)
    s = se.CuppenSolver(n, ref_leaves=8, vectors=vec)
    s.set_tridiagonal(D, E)
    best = None
    for it in range(6):
        s.solve(); t = s.timers()
        if it >= 2 and (best is None or t["device_s"] < best["device_s"]): best = t
    print(sys.argv[1], mat, n, "vectors" if vec else "eigenvalues only", "device_ms %.4f" % (best["device_s"] * 1e3), {k: round(best[k] * 1e3, 3) for k in ("root_finding_s", "deflation_s", "ev_extract_s", "backtransform_ev_s") if k in best}, "lam[0] %.17g lam[-1] %.17g" % (s.eigenvalues()[0], s.eigenvalues()[-1]), flush=True)
    s.close()
PY
cp $L /tmp/new.so
python /tmp/ab.py base > $O/ab_s.txt 2>&1
cp gpurun_tmp/libcuppen_b200_tiled256.so $L; python /tmp/ab.py tiled256 >> $O/ab_s.txt 2>&1
cp /tmp/new.so $L
cat $O/ab_s.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $O/pytest_s.txt 2>&1; echo "pytest rc $?" >> $O/pytest_s.txt; tail -3 $O/pytest_s.txt
