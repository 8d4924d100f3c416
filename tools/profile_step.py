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
a = ap.parse_args()
D, E = make_matrix(a.matrix, a.size)
s = se.CuppenSolver(a.size, ref_leaves=a.ref_leaves, vectors=True)
s.set_tridiagonal(D, E)
for _ in range(1 + a.reps):
    s.solve()
t = s.timers()
print("device_s %.6f launches %d gemm_s %.6f gemm_tflops %.2f" % (t["device_s"], t["kernel_launches"], t["gemm_s"],
      t["gemm_flop"] / max(t["gemm_s"], 1e-12) * 1e-12), "max resid %.3e" % s.residuals().max())
