"""Dense front end (cuppen_dense_eigh) on random symmetric matrices: phase times, residual and orthogonality."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from symmetric_eigenvalue_b200 import api  # noqa: E402

for n in [int(x) for x in (sys.argv[1:] or ["1024", "4096", "8192"])]:
    rng = np.random.default_rng(n)
    B = rng.normal(size=(n, n))
    A = (B + B.T) / 2
    api.dense_eigh(A[:256, :256])                      # warm the context / kernels
    t0 = time.perf_counter()
    w, Z, t = api.dense_eigh(A)
    wall = time.perf_counter() - t0
    R = A @ Z - Z * w[None, :]
    nA = np.abs(A).sum(axis=0).max()
    print("dense n=%d wall %.3f s  tridiagonalise %.3f s  tridiagonal solve %.3f s (device %.4f)  back-transformation %.3f s  "
          "max|AZ-ZW|/(n eps |A|) %.2f  max|Z^T Z-I|/(n eps) %.2f" % (n, wall, t["tridiagonalise_s"], t["tridiagonal_solve_s"], t["tridiagonal_device_s"],
          t["backtransform_s"], np.abs(R).max() / (n * 2.2e-16 * nA), np.abs(Z.T @ Z - np.eye(n)).max() / (n * 2.2e-16)), flush=True)
