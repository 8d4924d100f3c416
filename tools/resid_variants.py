#!/usr/bin/env python
"""Time the residual_kernel variants (columns per block x min blocks per SM) on random data."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from symmetric_eigenvalue_b200 import api
for n in (4096, 16384, 32768):
    row = {"n": n}
    for v in (14, 16, 23, 24, 25, 42, 43, 44, 82, 83):
        err, sec = api.selftest_residual(n, 0, 0, n, variant=v)
        row["v%d" % v] = {"us": round(sec * 1e6, 1), "GBps": round(8.0 * n * n / sec * 1e-9, 1), "err": err}
    print(json.dumps(row), flush=True)
