"""TEST INFRASTRUCTURE ONLY -- ctypes front end of oracle/_ref/libcuppen_oracle.so
(the plain-C restatement in cuppen_oracle.c) and a runner for oracle/_ref/cuppens_ref
(the unmodified reference built with the shims).  See oracle/cuppen_oracle.c header."""
import ctypes
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
_LIB = None


def build(ref=True):
    """(Re)build the oracle library and, if /root/reference is present, the reference binary."""
    target = "all" if ref else "oracle"
    subprocess.run(["make", "-C", HERE, target], check=True, stdout=subprocess.DEVNULL)


def load():
    global _LIB
    if _LIB is None:
        path = os.path.join(OUT, "libcuppen_oracle.so")
        if not os.path.exists(path):
            build(ref=False)
        lib = ctypes.CDLL(path)
        dp = ctypes.POINTER(ctypes.c_double)
        ip = ctypes.POINTER(ctypes.c_int)
        lib.cuppen_oracle_solve.restype = ctypes.c_int
        lib.cuppen_oracle_solve.argtypes = [ctypes.c_int, dp, dp, ctypes.c_int, dp, dp, dp, ip, dp, ip]
        lib.cuppen_oracle_merge.restype = ctypes.c_int
        lib.cuppen_oracle_merge.argtypes = [ctypes.c_int, dp, dp, ctypes.c_double, ip, dp, dp, ip, dp, dp]
        lib.cuppen_oracle_scheme.restype = None
        lib.cuppen_oracle_scheme.argtypes = [ctypes.c_int, ctypes.c_int, dp, dp]
        lib.cuppen_oracle_tql2.restype = ctypes.c_int
        lib.cuppen_oracle_tql2.argtypes = [ctypes.c_int, dp, dp, dp, ctypes.c_int, ctypes.c_int]
        _LIB = lib
    return _LIB


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _ip(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int))


def scheme(s, n):
    """createMatrixScheme1/2, /root/reference/src/helper.c:7-33."""
    D = np.empty(n); E = np.empty(max(n - 1, 1))
    load().cuppen_oracle_scheme(s, n, _dp(D), _dp(E))
    return D, E[: n - 1]


def solve(D, E, P, vectors=False, residuals=True):
    """Reference algorithm with a P-leaf tree.  Returns dict(lam, resid, V, stats, rhos)."""
    lib = load()
    D = np.ascontiguousarray(D, dtype=np.float64); E = np.ascontiguousarray(E, dtype=np.float64)
    n = D.size
    lam = np.empty(n); res = np.empty(n) if residuals else None
    V = np.empty((n, n), order="F") if vectors else None
    stats = np.zeros((max(2 * P, 2), 4), dtype=np.int32); rhos = np.zeros(max(2 * P, 2))
    nm = ctypes.c_int(0)
    Epad = E if E.size else np.zeros(1)
    rc = lib.cuppen_oracle_solve(n, _dp(D), _dp(Epad), P, _dp(lam), _dp(res) if residuals else None,
                                 _dp(V) if vectors else None, _ip(stats), _dp(rhos), ctypes.byref(nm))
    if rc != 0:
        raise RuntimeError("cuppen_oracle_solve rc=%d" % rc)
    k = nm.value
    return dict(lam=lam, resid=res, V=V, stats=stats[:k].copy(), rhos=rhos[:k].copy())


def merge(D, z, rho):
    """One rank-one merge (computeEigenvalues + computeNormalizationFactors)."""
    lib = load()
    D = np.array(D, dtype=np.float64); z = np.array(z, dtype=np.float64)
    m = D.size
    G = np.zeros(m, dtype=np.int32); P = np.zeros(m, dtype=np.int32)
    L = np.zeros(m); N = np.zeros(m); C = np.zeros(m); S = np.zeros(m)
    g = lib.cuppen_oracle_merge(m, _dp(D), _dp(z), float(rho), _ip(G), _dp(L), _dp(N), _ip(P), _dp(C), _dp(S))
    return dict(D=D, z=z, G=G, L=L, N=N, P=P[:g], C=C[:g], S=S[:g], numGR=g)


# ---- synthetic inputs of BASELINE.json / SURVEY.md section 8(d) -----------------------------
def rand_u(n, seed=1234):
    rng = np.random.default_rng(seed)
    d = rng.uniform(-1, 1, n); e = rng.uniform(-1, 1, n - 1)
    d[d == 0] = 0.5; e[e == 0] = 0.5
    return d, e


def goe(n, seed=7):
    """beta-Hermite (Dumitriu-Edelman) tridiagonal model of GOE, scaled so ||T|| ~ 2."""
    rng = np.random.default_rng(seed)
    d = rng.normal(0.0, np.sqrt(2.0), n)
    e = np.sqrt(rng.chisquare(np.arange(n - 1, 0, -1)))
    s = 1.0 / np.sqrt(n)
    return d * s, e * s


def wilkinson(n, norm=None):
    d = np.abs(np.arange(n) - (n - 1) / 2.0); e = np.ones(n - 1)
    if norm is not None:
        s = norm / (d.max() + 2.0)
        d, e = d * s, e * s
    return d, e


def write_mtx(path, D, E):
    """Matrix Market file the reference reader accepts (/root/reference/src/filehandling.c:76-153):
    coordinate real general, sub-diagonal entry before its super-diagonal twin."""
    n = len(D)
    with open(path, "w") as f:
        f.write("%%MatrixMarket matrix coordinate real general\n")
        f.write("%d %d %d\n" % (n, n, n + 2 * (n - 1)))
        for i in range(n):
            f.write("%d %d %.17g\n" % (i + 1, i + 1, D[i]))
            if i + 1 < n:
                f.write("%d %d %.17g\n" % (i + 2, i + 1, E[i]))
                f.write("%d %d %.17g\n" % (i + 1, i + 2, E[i]))


def ref_binary():
    return os.path.join(OUT, "cuppens_ref")


def run_reference(args, P=1, threads=1, timeout=600, stats=True, dump_dir=None):
    """Run the unmodified reference (oracle/_ref/cuppens_ref).  Returns dict(rc, stdout, out lines, stats)."""
    exe = ref_binary()
    if not os.path.exists(exe):
        raise FileNotFoundError(exe)
    env = dict(os.environ, MPISHIM_NP=str(P), OMP_NUM_THREADS=str(threads), OPENBLAS_NUM_THREADS="1")
    with tempfile.TemporaryDirectory() as td:
        st = os.path.join(td, "stats.txt")
        if stats:
            env["CUPPEN_ORACLE_STATS"] = st
        if dump_dir:
            env["CUPPEN_ORACLE_DUMP"] = dump_dir
        p = subprocess.run([exe] + list(args), env=env, capture_output=True, text=True,
                           timeout=timeout, start_new_session=True)
        rows = []
        if stats and os.path.exists(st):
            rows = [l.split() for l in open(st).read().strip().splitlines()]
    merges = sorted(((int(r[1]), int(r[0]), int(r[2]), int(r[3]), float(r[4])) for r in rows))
    return dict(rc=p.returncode, stdout=p.stdout, stderr=p.stderr,
                merges=[dict(m=m, off=o, zdefl=zd, givens=g, rho=rho) for (m, o, zd, g, rho) in merges])


def read_output(path):
    lam, res = [], []
    for line in open(path):
        t = line.split()
        if not t:
            continue
        lam.append(float(t[0])); res.append(float(t[1]) if len(t) > 1 else np.nan)
    return np.array(lam), np.array(res)
