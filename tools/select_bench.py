#!/usr/bin/env python
"""Selected-eigenvector mode (CUPPEN_FLAG_SELECT) timings: eigenvalue-only solve + back-application of K
columns, CUDA-event time inside the library.  One JSON line per (matrix, n, K).
usage: python tools/select_bench.py [--sizes 16384,65536] [--ks 1,8,16,64] [--matrix goe] [--ref-leaves 8]"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import symmetric_eigenvalue_b200 as se  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--sizes", default="16384,65536")
ap.add_argument("--ks", default="1,8,16,64")
ap.add_argument("--matrix", default="goe")
ap.add_argument("--ref-leaves", type=int, default=8)
ap.add_argument("--reps", type=int, default=4)
a = ap.parse_args()
for n in [int(x) for x in a.sizes.split(",")]:
    D, E = bench.make_matrix(a.matrix, n)
    s = se.CuppenSolver(n, ref_leaves=a.ref_leaves, select=True)
    s.set_tridiagonal(D, E)
    for K in [int(x) for x in a.ks.split(",")]:
        sel = np.unique(np.linspace(0, n - 1, K).astype(np.int32))
        s.select(sel)
        dev, app = [], []
        for it in range(a.reps + 2):
            s.solve()
            t = s.timers()
            if it >= 2:
                dev.append(t["device_s"]); app.append(t["apply_s"])
        pairs = sum(float(m.k) ** 2 for m in s.merge_stats())
        passes = (len(sel) + 7) // 8
        print(json.dumps({"matrix": a.matrix, "n": n, "ref_leaves": a.ref_leaves, "K": int(len(sel)),
                          "device_s": float(np.mean(dev)), "apply_s": float(np.mean(app)),
                          "eigenvalue_phase_s": float(np.mean(dev) - np.mean(app)),
                          "pole_root_pairs_per_pass": pairs, "passes": passes,
                          "gpairs_per_s": pairs * passes / max(np.mean(app), 1e-12) * 1e-9,
                          "max_residual": float(s.residuals(sel).max())}), flush=True)
    s.close()
