"""GEMM kernels alone on random data (profiles/): 0 cp.async 128x128 (odd rows), 1 TMA bulk-copy lines (even rows),
3 TMA bulk-copy lines with odd rows (lines fetched from one row earlier), 4 / 5 TMA tensor maps (even / odd rows)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from symmetric_eigenvalue_b200 import api  # noqa: E402

for (M, N, K) in [(8192, 8192, 8192), (4096, 4096, 4096), (1024, 14000, 7000), (2048, 4096, 2048)]:
    for v in (0, 1, 3, 4, 5):
        err, tf = api.selftest_gemm(v, M, N, K, reps=3)
        print("gemm variant", v, (M, N, K), "err %.2e" % err, "TF/s %.2f" % tf, flush=True)
