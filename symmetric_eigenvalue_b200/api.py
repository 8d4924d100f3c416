"""ctypes binding of the C ABI in include/cuppen_b200.h (one function per entry point)."""
import ctypes
import os
from collections import namedtuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

FLAG_VECTORS = 1
FLAG_NO_RESIDUALS = 2
FLAG_SELECT = 4
NCCL_ID_BYTES = 128


class CuppenError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("cuppen error %d: %s" % (code, msg))
        self.code = code


class _MergeStat(ctypes.Structure):
    _fields_ = [("offset", ctypes.c_int), ("m", ctypes.c_int), ("n1", ctypes.c_int), ("mode", ctypes.c_int),
                ("zdefl", ctypes.c_int), ("givens", ctypes.c_int), ("k", ctypes.c_int), ("height", ctypes.c_int),
                ("rho", ctypes.c_double)]


class _Timers(ctypes.Structure):
    _fields_ = [(k, ctypes.c_double) for k in (
        "total_s", "root_finding_s", "ev_extract_s", "backtransform_s", "backtransform_ev_s", "gemm_s",
        "gemm_flop", "leaf_s", "deflation_s", "pack_s", "residual_s", "device_s", "pack_bytes", "ugen_bytes",
        "secular_root_iters")] + [("kernel_launches", ctypes.c_long), ("apply_s", ctypes.c_double), ("comm_s", ctypes.c_double),
                                       ("comm_mode", ctypes.c_long)]


class _DenseTimers(ctypes.Structure):
    _fields_ = [(k, ctypes.c_double) for k in ("tridiagonalise_s", "tridiagonal_solve_s", "tridiagonal_device_s", "backtransform_s")]


_BCAST_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int,
                             ctypes.c_int, ctypes.c_int)
_ALLRED_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.POINTER(ctypes.c_double), ctypes.c_size_t)
_ALLGATHER_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t)
_ALLRED_I32_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.POINTER(ctypes.c_int), ctypes.c_size_t)
_ALLTOALLV_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_size_t),
                                 ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_size_t))


class _Callbacks(ctypes.Structure):
    _fields_ = [("user", ctypes.c_void_p), ("group_bcast", _BCAST_FN), ("allreduce_sum_f64", _ALLRED_FN),
                ("allgather", _ALLGATHER_FN), ("allreduce_sum_i32", _ALLRED_I32_FN), ("alltoallv", _ALLTOALLV_FN)]


MergeStat = namedtuple("MergeStat", "offset m n1 mode zdefl givens k height rho")


def library_path():
    return os.path.join(_HERE, "lib", "libcuppen_b200.so")


def _declare(lib):
    dp = ctypes.POINTER(ctypes.c_double)
    ip = ctypes.POINTER(ctypes.c_int)
    H = ctypes.c_void_p
    sig = {
        "cuppen_create": [ctypes.POINTER(H), ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int],
        "cuppen_nccl_unique_id": [ctypes.c_char_p],
        "cuppen_create_nccl": [ctypes.POINTER(H), ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                               ctypes.c_int, ctypes.c_int, ctypes.c_char_p],
        "cuppen_create_callbacks": [ctypes.POINTER(H), ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                    ctypes.c_int, ctypes.c_int, ctypes.POINTER(_Callbacks)],
        "cuppen_destroy": [H],
        "cuppen_set_tridiagonal": [H, dp, dp],
        "cuppen_solve": [H],
        "cuppen_get_eigenvalues": [H, dp],
        "cuppen_get_residuals": [H, ip, ctypes.c_int, dp],
        "cuppen_get_merge_stats": [H, ctypes.POINTER(_MergeStat), ctypes.c_int, ip],
        "cuppen_get_timers": [H, ctypes.POINTER(_Timers)],
        "cuppen_local_rows": [H, ip, ip],
        "cuppen_local_row_map": [H, ip],
        "cuppen_copy_eigenvectors": [H, dp, ctypes.c_long],
        "cuppen_copy_eigenvector_columns": [H, ip, ctypes.c_int, dp, ctypes.c_long],
        "cuppen_select_eigenvectors": [H, ip, ctypes.c_int],
        "cuppen_copy_selected_eigenvectors": [H, dp, ctypes.c_long],
        "cuppen_orthogonality": [H, dp, dp],
        "cuppen_write_eigenvectors": [H, ctypes.c_char_p],
        "cuppen_dense_eigh": [ctypes.c_int, dp, ctypes.c_long, dp, dp, ctypes.c_long, ctypes.c_int, ctypes.POINTER(_DenseTimers)],
        "cuppen_scheme": [ctypes.c_int, ctypes.c_int, dp, dp],
        "cuppen_read_mtx": [ctypes.c_char_p, ctypes.POINTER(dp), ctypes.POINTER(dp), ip],
        "cuppen_read_ev_file": [ctypes.c_char_p, ctypes.c_int, ctypes.POINTER(ip), ip],
        "cuppen_write_results": [ctypes.c_char_p, ctypes.c_int, dp, dp, ctypes.c_int, ip, ctypes.c_int],
    }
    for name, args in sig.items():
        fn = getattr(lib, name)      # AttributeError here = the library does not export the ABI
        fn.argtypes = args
        fn.restype = ctypes.c_int
    lib.cuppen_last_error.argtypes = []
    lib.cuppen_last_error.restype = ctypes.c_char_p
    return lib


EXPORTED_SYMBOLS = (
    "cuppen_create", "cuppen_nccl_unique_id", "cuppen_create_nccl", "cuppen_create_callbacks", "cuppen_destroy",
    "cuppen_set_tridiagonal", "cuppen_solve", "cuppen_get_eigenvalues", "cuppen_get_residuals",
    "cuppen_get_merge_stats", "cuppen_get_timers", "cuppen_local_rows", "cuppen_local_row_map", "cuppen_copy_eigenvectors", "cuppen_copy_eigenvector_columns",
    "cuppen_select_eigenvectors", "cuppen_copy_selected_eigenvectors", "cuppen_orthogonality", "cuppen_write_eigenvectors",
    "cuppen_last_error", "cuppen_dense_eigh", "cuppen_scheme", "cuppen_read_mtx", "cuppen_read_ev_file", "cuppen_write_results",
)


def load_library(path=None):
    """Load libcuppen_b200.so (built in-tree by ``make`` / ``__graft_entry__.build()``)."""
    global _LIB
    if path is None and _LIB is not None:
        return _LIB
    p = path or library_path()
    if not os.path.exists(p):
        raise CuppenError(-10, "CUDA library %s is missing: run `make` (there is no CPU fallback)" % p)
    lib = _declare(ctypes.CDLL(p))
    if path is None:
        _LIB = lib
    return lib


def _chk(lib, rc):
    if rc != 0:
        raise CuppenError(rc, lib.cuppen_last_error().decode("utf-8", "replace"))


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def nccl_unique_id(lib=None):
    lib = lib or load_library()
    buf = ctypes.create_string_buffer(NCCL_ID_BYTES)
    _chk(lib, lib.cuppen_nccl_unique_id(buf))
    return buf.raw


# ---- test / bench only: lib/libcuppen_selftest.so (csrc/selftest.cu, csrc/cuppen_selftest.h) --------------------
_SELFTEST = None
SELFTEST_SYMBOLS = ("cuppen_measure_fp64_peak", "cuppen_measure_fp64_mix", "cuppen_selftest_gemm", "cuppen_selftest_residual",
                    "cuppen_selftest_rcp", "cuppen_selftest_last_error")


def selftest_library_path():
    return os.path.join(_HERE, "lib", "libcuppen_selftest.so")


def load_selftest_library():
    """Kernel self-tests and FP64 yardsticks; not part of the product ABI."""
    global _SELFTEST
    if _SELFTEST is None:
        p = selftest_library_path()
        if not os.path.exists(p):
            raise CuppenError(-10, "%s is missing: run `make lib`" % p)
        lib = ctypes.CDLL(p)
        dp = ctypes.POINTER(ctypes.c_double)
        i = ctypes.c_int
        for name, args in {"cuppen_measure_fp64_peak": [i, i, dp, dp], "cuppen_measure_fp64_mix": [i, dp, dp, dp, dp],
                           "cuppen_selftest_gemm": [i, i, i, i, i, i, dp, dp],
                           "cuppen_selftest_residual": [i, i, i, i, i, i, dp, dp],
                           "cuppen_selftest_rcp": [i, ctypes.c_long, dp, dp, dp]}.items():
            fn = getattr(lib, name)
            fn.argtypes = args
            fn.restype = ctypes.c_int
        lib.cuppen_selftest_last_error.argtypes = []
        lib.cuppen_selftest_last_error.restype = ctypes.c_char_p
        _SELFTEST = lib
    return _SELFTEST


def _chk_st(lib, rc):
    if rc != 0:
        raise CuppenError(rc, lib.cuppen_selftest_last_error().decode("utf-8", "replace"))


def measure_fp64_peak(device=0, ms=200):
    """(DMMA TFLOP/s, DFMA TFLOP/s) of register-resident issue loops on `device`."""
    lib = load_selftest_library()
    a, b = ctypes.c_double(0), ctypes.c_double(0)
    _chk_st(lib, lib.cuppen_measure_fp64_peak(device, ms, ctypes.byref(a), ctypes.byref(b)))
    return a.value, b.value


def measure_fp64_mix(device=0):
    """dict of TFLOP/s: DMMA / DFMA issue loops alone and sharing the SMs (do they share the FP64 units?)."""
    lib = load_selftest_library()
    v = [ctypes.c_double(0) for _ in range(4)]
    _chk_st(lib, lib.cuppen_measure_fp64_mix(device, *[ctypes.byref(x) for x in v]))
    return dict(zip(("dmma_alone", "dfma_alone", "dmma_mixed", "dfma_mixed"), (x.value for x in v)))


def selftest_gemm(variant, M, N, K, reps=3, device=0):
    """(max abs error on sampled entries, TFLOP/s) of one back-transformation GEMM kernel variant."""
    lib = load_selftest_library()
    e, t = ctypes.c_double(0), ctypes.c_double(0)
    _chk_st(lib, lib.cuppen_selftest_gemm(device, variant, M, N, K, reps, ctypes.byref(e), ctypes.byref(t)))
    return e.value, t.value


def selftest_residual(n, g0, l0, cnt, variant=0, device=0):
    """(max relative error, seconds) of residual_kernel on one slice of rows (multi-GPU layout) of random data."""
    lib = load_selftest_library()
    e, t = ctypes.c_double(0), ctypes.c_double(0)
    _chk_st(lib, lib.cuppen_selftest_residual(device, n, variant, g0, l0, cnt, ctypes.byref(e), ctypes.byref(t)))
    return e.value, t.value


def selftest_rcp(count=1 << 24, device=0):
    """(max |1 - x*seed|, max ulps off 1/x with two Newton steps, with the cubic step) of the kernels' reciprocal."""
    lib = load_selftest_library()
    v = [ctypes.c_double(0) for _ in range(3)]
    _chk_st(lib, lib.cuppen_selftest_rcp(device, count, *[ctypes.byref(x) for x in v]))
    return tuple(x.value for x in v)


class CuppenSolver:
    """One tridiagonal eigenproblem on one GPU (or one rank's share of it).

    n            matrix size
    ref_leaves   P of the reference run ``mpirun -n P cuppens`` whose divide tree, theta rule and
                 deflation thresholds are reproduced at the top log2(P) levels (1: none)
    vectors      materialise eigenvectors (the reference's ``-e``)
    select       selected-eigenvector mode (the reference's ``-eFILE``): no n x n matrix, see ``select()``
    """

    def __init__(self, n, ref_leaves=1, vectors=True, residuals=True, device=0, rank=0, world=1,
                 nccl_id=None, callbacks=None, lib=None, select=False):
        self.lib = lib or load_library()
        self.n = int(n)
        self.select_mode = bool(select)
        vectors = bool(vectors) and not self.select_mode
        self.vectors = vectors
        self._nsel = 0
        flags = (FLAG_VECTORS if vectors else 0) | (0 if residuals else FLAG_NO_RESIDUALS) | (FLAG_SELECT if select else 0)
        self._h = ctypes.c_void_p()
        self._cb = None
        if callbacks is not None:
            self._cb = callbacks                     # keep the ctypes thunks alive
            rc = self.lib.cuppen_create_callbacks(ctypes.byref(self._h), self.n, ref_leaves, flags, device, rank,
                                                  world, ctypes.byref(callbacks))
        elif world > 1:
            if nccl_id is None or len(nccl_id) != NCCL_ID_BYTES:
                raise CuppenError(-1, "world > 1 needs the 128-byte NCCL unique id of rank 0")
            rc = self.lib.cuppen_create_nccl(ctypes.byref(self._h), self.n, ref_leaves, flags, device, rank, world,
                                             nccl_id)
        else:
            rc = self.lib.cuppen_create(ctypes.byref(self._h), self.n, ref_leaves, flags, device)
        _chk(self.lib, rc)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self.lib.cuppen_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_tridiagonal(self, D, E):
        D = np.ascontiguousarray(D, dtype=np.float64)
        E = np.ascontiguousarray(E, dtype=np.float64)
        if D.size != self.n or E.size != max(self.n - 1, 0):
            raise CuppenError(-1, "D must have n and E n-1 entries")
        if E.size == 0:
            E = np.zeros(1)
        _chk(self.lib, self.lib.cuppen_set_tridiagonal(self._h, _dp(D), _dp(E)))

    def solve(self):
        _chk(self.lib, self.lib.cuppen_solve(self._h))

    def select(self, indices):
        """0-based ranks (ascending-lambda order) of the eigenvectors wanted from the next solve()."""
        idx = np.ascontiguousarray(indices, dtype=np.int32).ravel()
        _chk(self.lib, self.lib.cuppen_select_eigenvectors(self._h, idx.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), idx.size))
        self._nsel = int(idx.size)

    def selected_eigenvectors(self):
        """n x cnt matrix, column t = eigenvector of the t-th selected rank."""
        V = np.empty((self.n, self._nsel), order="F")
        if self._nsel:
            _chk(self.lib, self.lib.cuppen_copy_selected_eigenvectors(self._h, _dp(V), self.n))
        return V

    def orthogonality(self):
        """(max |V^T V - I|, device seconds) -- evaluated on the GPU, the product is never formed."""
        a, b = ctypes.c_double(0), ctypes.c_double(0)
        _chk(self.lib, self.lib.cuppen_orthogonality(self._h, ctypes.byref(a), ctypes.byref(b)))
        return a.value, b.value

    def write_eigenvectors(self, filename):
        """CUPPENV1 file; several ranks: every rank calls, rank 0 writes (filename may be None elsewhere)."""
        _chk(self.lib, self.lib.cuppen_write_eigenvectors(self._h, None if filename is None else os.fsencode(filename)))

    def eigenvalues(self):
        out = np.empty(self.n)
        _chk(self.lib, self.lib.cuppen_get_eigenvalues(self._h, _dp(out)))
        return out

    def residuals(self, indices=None):
        if indices is None:
            out = np.empty(self.n)
            _chk(self.lib, self.lib.cuppen_get_residuals(self._h, None, self.n, _dp(out)))
            return out
        idx = np.ascontiguousarray(indices, dtype=np.int32)
        out = np.empty(idx.size)
        _chk(self.lib, self.lib.cuppen_get_residuals(self._h, idx.ctypes.data_as(ctypes.POINTER(ctypes.c_int)),
                                                     idx.size, _dp(out)))
        return out

    def merge_stats(self):
        cnt = ctypes.c_int(0)
        _chk(self.lib, self.lib.cuppen_get_merge_stats(self._h, None, 0, ctypes.byref(cnt)))
        arr = (_MergeStat * max(cnt.value, 1))()
        _chk(self.lib, self.lib.cuppen_get_merge_stats(self._h, arr, cnt.value, ctypes.byref(cnt)))
        return [MergeStat(*(getattr(arr[i], f[0]) for f in _MergeStat._fields_)) for i in range(cnt.value)]

    def timers(self):
        t = _Timers()
        _chk(self.lib, self.lib.cuppen_get_timers(self._h, ctypes.byref(t)))
        return {f[0]: getattr(t, f[0]) for f in _Timers._fields_}

    def local_rows(self):
        r0, rows = ctypes.c_int(0), ctypes.c_int(0)
        _chk(self.lib, self.lib.cuppen_local_rows(self._h, ctypes.byref(r0), ctypes.byref(rows)))
        return r0.value, rows.value

    def local_row_map(self):
        """Global row index of every local row of eigenvectors()."""
        _, rows = self.local_rows()
        out = np.zeros(max(rows, 1), dtype=np.int32)
        _chk(self.lib, self.lib.cuppen_local_row_map(self._h, out.ctypes.data_as(ctypes.POINTER(ctypes.c_int))))
        return out[:rows]

    def eigenvector_columns(self, indices):
        """Local rows of the eigenvectors with the given 0-based ranks (ascending lambda): rows x len(indices)."""
        idx = np.ascontiguousarray(indices, dtype=np.int32).ravel()
        _, rows = self.local_rows()
        V = np.empty((max(rows, 1), idx.size), order="F")
        _chk(self.lib, self.lib.cuppen_copy_eigenvector_columns(self._h, idx.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), idx.size,
                                                                _dp(V), max(rows, 1)))
        return V[:rows]

    def eigenvectors(self):
        """Rows of V owned by this rank (all rows when world == 1), columns in ascending-lambda order."""
        _, rows = self.local_rows()
        V = np.empty((rows, self.n), order="F")
        _chk(self.lib, self.lib.cuppen_copy_eigenvectors(self._h, _dp(V), rows))
        return V


# ---- mirrors of the reference's helper / file functions (host side of the C ABI) --------------------
def _scheme(s, n, lib=None):
    lib = lib or load_library()
    D = np.empty(n)
    E = np.empty(max(n - 1, 1))
    _chk(lib, lib.cuppen_scheme(s, n, _dp(D), _dp(E)))
    return D, E[: n - 1]


def createMatrixScheme1(n, lib=None):
    """src/helper.h:28 -- d_i evenly spaced in [1,100], e = -1."""
    return _scheme(1, n, lib)


def createMatrixScheme2(n, lib=None):
    """src/helper.h:39 -- [-1 2 -1]."""
    return _scheme(2, n, lib)


def readSymmTriadiagonalMatrixFromSparseMTX(filename, lib=None):
    """src/filehandling.h:56 -- returns (D, E); raises CuppenError(-2) where the reference returns -1."""
    lib = lib or load_library()
    dp = ctypes.POINTER(ctypes.c_double)
    D, E, n = dp(), dp(), ctypes.c_int(0)
    _chk(lib, lib.cuppen_read_mtx(os.fsencode(filename), ctypes.byref(D), ctypes.byref(E), ctypes.byref(n)))
    libc = ctypes.CDLL(None)
    libc.free.argtypes = [ctypes.c_void_p]
    try:
        d = np.ctypeslib.as_array(D, shape=(n.value,)).copy()
        e = np.ctypeslib.as_array(E, shape=(max(n.value - 1, 1),)).copy()[: n.value - 1]
    finally:
        libc.free(D)
        libc.free(E)
    return d, e


def determineEigenvectorsToCompute(filename, n, lib=None):
    """src/filehandling.h:69 -- sorted 0-based ranks (ascending-lambda order) listed in the -eFILE file."""
    lib = lib or load_library()
    ip = ctypes.POINTER(ctypes.c_int)
    idx, cnt = ip(), ctypes.c_int(0)
    _chk(lib, lib.cuppen_read_ev_file(os.fsencode(filename), n, ctypes.byref(idx), ctypes.byref(cnt)))
    libc = ctypes.CDLL(None)
    libc.free.argtypes = [ctypes.c_void_p]
    out = np.ctypeslib.as_array(idx, shape=(max(cnt.value, 1),)).copy()[: cnt.value] if cnt.value else np.zeros(0, np.int32)
    libc.free(idx)
    return out


def writeResults(filename, lam, resid=None, indices=None, lib=None):
    """src/filehandling.h:81 (output part) -- '%20.19g %20.19g' / '%20.19g' lines in ascending lambda."""
    lib = lib or load_library()
    lam = np.ascontiguousarray(lam, dtype=np.float64)
    n = lam.size
    rp = _dp(np.ascontiguousarray(resid, dtype=np.float64)) if resid is not None else None
    if indices is None:
        rc = lib.cuppen_write_results(os.fsencode(filename), n, _dp(lam), rp, 1 if resid is not None else 0, None, 0)
    else:
        idx = np.ascontiguousarray(indices, dtype=np.int32)
        rc = lib.cuppen_write_results(os.fsencode(filename), n, _dp(lam), rp, 0,
                                      idx.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), idx.size)
    _chk(lib, rc)


def read_eigenvector_file(filename):
    """Reader of the CUPPENV1 file written by cuppen_write_eigenvectors / `cuppens -v`: (ranks, lam, V[n, ncols])."""
    with open(filename, "rb") as f:
        if f.read(8) != b"CUPPENV1":
            raise CuppenError(-2, "%s is not a CUPPENV1 eigenvector file" % filename)
        n, ncols = (int(x) for x in np.fromfile(f, dtype="<i8", count=2))
        ranks = np.fromfile(f, dtype="<i8", count=ncols)
        lam = np.fromfile(f, dtype="<f8", count=ncols)
        V = np.fromfile(f, dtype="<f8", count=n * ncols)
    if V.size != n * ncols:
        raise CuppenError(-2, "%s is truncated" % filename)
    return ranks, lam, V.reshape(ncols, n).T


def dense_eigh(A, vectors=True, device=0, lib=None):
    """Dense symmetric eigenproblem on one GPU (Householder tridiagonalisation + the tridiagonal path + back-transformation).
    Returns (w ascending, Z or None, timers dict)."""
    lib = lib or load_library()
    A = np.asfortranarray(A, dtype=np.float64)
    n = A.shape[0]
    if A.ndim != 2 or A.shape[1] != n:
        raise CuppenError(-1, "A must be square")
    w = np.empty(n)
    Z = np.empty((n, n), order="F") if vectors else None
    t = _DenseTimers()
    _chk(lib, lib.cuppen_dense_eigh(n, _dp(A), max(n, 1), _dp(w), _dp(Z) if vectors else None, max(n, 1), device, ctypes.byref(t)))
    return w, Z, {f[0]: getattr(t, f[0]) for f in _DenseTimers._fields_}


def cuppens(D, E, ref_leaves=1, vectors=True, device=0, lib=None, select=None):
    """Convenience: full decomposition.  Returns dict(lam, resid, V, stats, timers).
    select = list of 0-based ranks: selected-eigenvector mode, V is n x len(select), resid per selected rank."""
    s = CuppenSolver(len(D), ref_leaves=ref_leaves, vectors=vectors, device=device, lib=lib, select=select is not None)
    try:
        s.set_tridiagonal(D, E)
        if select is not None:
            s.select(select)
        s.solve()
        out = dict(lam=s.eigenvalues(), stats=s.merge_stats(), timers=s.timers(), resid=None, V=None)
        if select is not None:
            out["resid"] = s.residuals(select) if len(select) else np.zeros(0)
            out["V"] = s.selected_eigenvectors()
        elif vectors:
            out["resid"] = s.residuals()
            out["V"] = s.eigenvectors()
        return out
    finally:
        s.close()
