"""One warm-up solve + `--reps` solves of a BASELINE configuration, for ncu (profiles/README.md)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import symmetric_eigenvalue_b200 as se  # noqa: E402
from bench import make_matrix  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=4096)
ap.add_argument("--matrix", default="s1")
ap.add_argument("--ref-leaves", type=int, default=8)
ap.add_argument("--reps", type=int, default=1)
ap.add_argument("--select", type=int, default=0, help="selected-eigenvector mode on this many evenly spaced ranks")
ap.add_argument("--orth", action="store_true", help="also run the on-GPU orthogonality check")
a = ap.parse_args()
D, E = make_matrix(a.matrix, a.size)
if a.select:
    import numpy as np
    s = se.CuppenSolver(a.size, ref_leaves=a.ref_leaves, select=True)
    s.set_tridiagonal(D, E)
    sel = np.unique(np.linspace(0, a.size - 1, a.select).astype(np.int32))
    s.select(sel)
    for _ in range(1 + a.reps):
        s.solve()
    t = s.timers()
    print("select K=%d device_s %.6f apply_s %.6f launches %d max resid %.3e" % (len(sel), t["device_s"], t["apply_s"],
          t["kernel_launches"], s.residuals(sel).max()))
    sys.exit(0)
s = se.CuppenSolver(a.size, ref_leaves=a.ref_leaves, vectors=True)
s.set_tridiagonal(D, E)
for _ in range(1 + a.reps):
    s.solve()
t = s.timers()
if a.orth:
    print("orthogonality %.3e in %.6f s" % s.orthogonality())
print("device_s %.6f launches %d gemm_s %.6f gemm_tflops %.2f" % (t["device_s"], t["kernel_launches"], t["gemm_s"],
      t["gemm_flop"] / max(t["gemm_s"], 1e-12) * 1e-12), "max resid %.3e" % s.residuals().max())
