#!/bin/bash
# r02 call P (1 GPU): cluster compact kernel: parity suite, phase times, launch list of GOE n=16384
O=gpurun_out/r02; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_p.txt 2>&1; echo "pytest rc $?" >> $O/pytest_p.txt; tail -3 $O/pytest_p.txt
cat > /tmp/ab.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import symmetric_eigenvalue_b200 as se
from bench import make_matrix
for mat, n in (("goe", 16384), ("wilk", 16384), ("s1", 4096)):
    D, E = make_matrix(mat, n)
    s = se.CuppenSolver(n, ref_leaves=8, vectors=True)
    s.set_tridiagonal(D, E)
    best = None
    for it in range(6):
        s.solve(); t = s.timers()
        if it >= 2 and (best is None or t["device_s"] < best["device_s"]): best = t
    print(sys.argv[1], mat, n, "device_ms %.4f" % (best["device_s"] * 1e3), {k: round(best[k] * 1e3, 3) for k in ("pack_s", "gemm_s", "residual_s", "deflation_s", "root_finding_s", "ev_extract_s", "backtransform_ev_s") if k in best}, "resid %.3e" % s.residuals().max(), flush=True)
    s.close()
PY
python /tmp/ab.py cluster > $O/ab_p.txt 2>&1; cat $O/ab_p.txt
python tools/profile_step.py --size 16384 --matrix goe > $O/prof_plain_p.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_goe16k_p.csv python tools/profile_step.py --size 16384 --matrix goe > $O/ncu_launch_p.log 2>&1
python - <<'PY'
import csv, collections
rows=list(csv.reader(open("gpurun_out/r02/launches_goe16k_p.csv")))
hi=[i for i,r in enumerate(rows) if "Kernel Name" in r][0]
h=rows[hi]; kn=h.index("Kernel Name"); mv=h.index("Metric Value")
data=[r for r in rows[hi+1:] if len(r)>mv]
names=[r[kn] for r in data]
last=max(i for i,nm in enumerate(names) if nm.startswith("leaf_ql"))
agg=collections.OrderedDict()
for r in data[last:]:
    a=agg.setdefault(r[kn].split("(")[0][:50],[0,0.0]); a[0]+=1; a[1]+=float(r[mv].replace(",",""))/1e3
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1]): print("%-52s %3d %10.1f us" % (k, v[0], v[1]))
PY
