#!/bin/bash
# r02 call R (1 GPU): secular kernel variants (err from the sums = current build; + unroll 8; + 3 CTAs per SM), eigenvalue-only n=65536
O=gpurun_out/r02; mkdir -p $O
L=symmetric_eigenvalue_b200/lib/libcuppen_b200.so
cat > /tmp/ab.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import symmetric_eigenvalue_b200 as se
from bench import make_matrix
for mat, n, vec in (("goe", 16384, True), ("goe", 65536, False)):
    D, E = make_matrix(mat, n)
    s = se.CuppenSolver(n, ref_leaves=8, vectors=vec)
    s.set_tridiagonal(D, E)
    best = None
    for it in range(5):
        s.solve(); t = s.timers()
        if it >= 2 and (best is None or t["device_s"] < best["device_s"]): best = t
    print(sys.argv[1], mat, n, "vectors" if vec else "eigenvalues only", "device_ms %.4f" % (best["device_s"] * 1e3), {k: round(best[k] * 1e3, 3) for k in ("root_finding_s", "deflation_s", "ev_extract_s") if k in best}, "lam[0] %.17g lam[-1] %.17g" % (s.eigenvalues()[0], s.eigenvalues()[-1]), flush=True)
    s.close()
PY
cp $L /tmp/new.so
python /tmp/ab.py base > $O/ab_r.txt 2>&1
cp gpurun_tmp/libcuppen_b200_unroll8.so $L; python /tmp/ab.py unroll8 >> $O/ab_r.txt 2>&1
cp gpurun_tmp/libcuppen_b200_sec3cta.so $L; python /tmp/ab.py sec3cta >> $O/ab_r.txt 2>&1
cp /tmp/new.so $L
cat $O/ab_r.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $O/pytest_r.txt 2>&1; echo "pytest rc $?" >> $O/pytest_r.txt; tail -3 $O/pytest_r.txt
