#!/usr/bin/env python
"""Back-transformation GEMM of one BASELINE workload under the tile-order / L2-hint settings of gemm_tma.h
(CUPPEN_GEMM_HINT, CUPPEN_SUPERCOL_MB are read when a handle is created): one handle per setting, `--reps` timed solves
each.  Run plain for the times, under `ncu --metrics dram__bytes_read.sum,... -k regex:dgemm_tma` for the DRAM traffic
(the launches appear in the order of the settings, (1 + reps) solves each).
usage: python tools/gemm_hint_sweep.py [--size 16384] [--matrix goe] [--configs 0:48,1:48,3:48,1:72] [--reps 3]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import symmetric_eigenvalue_b200 as se  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=16384)
ap.add_argument("--matrix", default="goe")
ap.add_argument("--ref-leaves", type=int, default=8)
ap.add_argument("--configs", default="0:48,1:48,3:48,1:72")
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
D, E = bench.make_matrix(a.matrix, a.size)
for cfg in a.configs.split(","):
    hint, mb = cfg.split(":")
    os.environ["CUPPEN_GEMM_HINT"] = hint
    os.environ["CUPPEN_SUPERCOL_MB"] = mb
    s = se.CuppenSolver(a.size, ref_leaves=a.ref_leaves, vectors=True)
    s.set_tridiagonal(D, E)
    dev, gemm, tf = [], [], []
    for it in range(1 + a.reps):
        s.solve()
        t = s.timers()
        if it >= 1:
            dev.append(t["device_s"]); gemm.append(t["gemm_s"]); tf.append(t["gemm_flop"] / max(t["gemm_s"], 1e-12) * 1e-12)
    print(json.dumps({"matrix": a.matrix, "n": a.size, "hint": int(hint), "supercol_mb": int(mb), "device_s": min(dev) if dev else None,
                      "gemm_s": min(gemm) if gemm else None, "gemm_tflops": max(tf) if tf else None,
                      "max_residual": float(s.residuals().max())}), flush=True)
    s.close()
