import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import symmetric_eigenvalue_b200 as se
n, P = 512, 8
D, E = se.createMatrixScheme2(n)
full = se.cuppens(D, E, ref_leaves=P)
sel = [499, 433, 131, 100]
out = se.cuppens(D, E, ref_leaves=P, select=sel)
lam = full["lam"]
for t, r in enumerate(sel):
    a, b = out["V"][:, t], full["V"][:, r]
    print("rank", r, "|a-b|", np.abs(a - b).max(), "|a+b|", np.abs(a + b).max(), "gaps", lam[r] - lam[r - 1], lam[r + 1] - lam[r] if r + 1 < n else None,
          "vs rank-1", min(np.abs(a - full["V"][:, r - 1]).max(), np.abs(a + full["V"][:, r - 1]).max()),
          "vs rank+1", min(np.abs(a - full["V"][:, r + 1]).max(), np.abs(a + full["V"][:, r + 1]).max()) if r + 1 < n else None,
          "dlam", out["lam"][r] - lam[r], flush=True)
st = [s for s in full["stats"] if s.mode == 1]
print(st)
V = full["V"]
print("full orth", np.abs(V.T @ V - np.eye(n)).max())
