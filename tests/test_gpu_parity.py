"""Parity tests proper: the CUDA path (through the C ABI of libcuppen_b200.so) against the reference's
golden outputs, the CPU oracle on the same seeded inputs, and size-independent properties at the
BASELINE sizes.  Tolerances are BASELINE.json's: eigenvalues within 1e-12*||T||, residual and
orthogonality at or below the reference's, identical deflation counts per merge."""
import os
import subprocess

import numpy as np
import pytest

import symmetric_eigenvalue_b200 as se
from conftest import ROOT, check_against_golden, check_select_against_efile_golden, golden_cases, load_golden, norm_T, ref_stats

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib(product_lib):
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return product_lib


@pytest.mark.parametrize("name", golden_cases())
def test_matches_reference_goldens(lib, name):
    g = load_golden(name)
    vec = bool(np.isfinite(g["resid"]).any())
    out = se.cuppens(g["D"], g["E"], ref_leaves=g["P"], vectors=vec, lib=lib)
    check_against_golden(g, out, vec)
    # the same run with eigenvectors: deflation bookkeeping and eigenvalues must not depend on the mode
    out2 = se.cuppens(g["D"], g["E"], ref_leaves=g["P"], vectors=not vec, lib=lib)
    assert np.abs(out2["lam"] - out["lam"]).max() <= 1e-13 * norm_T(g["D"], g["E"])
    assert ref_stats(out2["stats"]) == ref_stats(out["stats"])


@pytest.mark.parametrize("gen,n,P", [("goe", 700, 4), ("rand_u", 512, 8), ("wilk", 600, 4), ("s1", 777, 3), ("s2", 512, 8),
                                     ("goe", 300, 2), ("s2", 96, 6)])
def test_against_oracle_seeded(lib, oracle, gen, n, P):
    D, E = {"goe": oracle.goe, "rand_u": oracle.rand_u, "wilk": lambda k: oracle.wilkinson(k, norm=64.0),
            "s1": lambda k: oracle.scheme(1, k), "s2": lambda k: oracle.scheme(2, k)}[gen](n)
    o = oracle.solve(D, E, P, vectors=True)
    out = se.cuppens(D, E, ref_leaves=P, lib=lib)
    nT = norm_T(D, E)
    assert np.abs(out["lam"] - o["lam"]).max() <= 1e-12 * nT
    assert ref_stats(out["stats"]) == sorted((int(m), int(off), int(zd), int(gv)) for (off, m, zd, gv) in o["stats"].tolist())
    assert out["resid"].max() <= o["resid"].max() * 1.05 + 4 * 2.2e-16 * nT
    mine = np.abs(out["V"].T @ out["V"] - np.eye(n)).max()
    ref = np.abs(o["V"].T @ o["V"] - np.eye(n)).max()
    assert mine <= max(ref, 1e-13)


@pytest.mark.parametrize("gen,n", [("goe", 2048), ("rand_u", 1500), ("wilk", 1001), ("s1", 4096), ("s2", 3000), ("s2", 33), ("s1", 2)])
def test_accurate_mode_against_lapack(lib, oracle, gen, n):
    from scipy.linalg import eigvalsh_tridiagonal
    D, E = {"goe": oracle.goe, "rand_u": oracle.rand_u, "wilk": oracle.wilkinson,
            "s1": lambda k: oracle.scheme(1, k), "s2": lambda k: oracle.scheme(2, k)}[gen](n)
    out = se.cuppens(D, E, ref_leaves=1, lib=lib)
    nT = norm_T(D, E)
    assert np.abs(out["lam"] - eigvalsh_tridiagonal(D, E)).max() < 3e-14 * nT
    assert out["resid"].max() < 1e-14 * nT
    V = out["V"]
    assert np.abs(V.T @ V - np.eye(n)).max() < 1e-13


@pytest.mark.parametrize("name", __import__("families").FAMILIES)
def test_matrix_families(lib, name):
    """Graded, glued, clustered, badly scaled ... inputs under the accurate rule, against LAPACK."""
    import families
    D, E = families.family(name, 777)
    families.check_accurate(se.cuppens(D, E, ref_leaves=1, lib=lib), D, E)


def test_baseline_config_properties(lib, oracle):
    """BASELINE configs[1]: -s 1 -n 4096 with eigenvectors (reference tree P=8): properties that
    need no oracle -- T V = V Lambda, V^T V = I, trace, ordering."""
    n = 4096
    D, E = se.createMatrixScheme1(n, lib=lib)
    out = se.cuppens(D, E, ref_leaves=8, lib=lib)
    lam, V = out["lam"], out["V"]
    assert (np.diff(lam) >= 0).all()
    assert abs(lam.sum() - D.sum()) < 1e-9 * abs(D.sum())
    TV = D[:, None] * V
    TV[1:] += E[:, None] * V[:-1]
    TV[:-1] += E[:, None] * V[1:]
    r = np.linalg.norm(TV - V * lam[None, :], axis=0)
    assert np.allclose(r, out["resid"], rtol=1e-6, atol=1e-12)
    assert r.max() < 2e-6          # set by the reference's 1e-6 z-deflation (golden: 1.3e-6 at n=1024)
    assert np.abs(V.T @ V - np.eye(n)).max() < 1e-12
    g = load_golden("s1_n4096_p8")
    assert np.abs(lam - g["lam"]).max() <= 1e-12 * norm_T(D, E)


def test_cli_drop_in(lib, tmp_path):
    """The `cuppens` executable: same stdout lines and output file format as the reference
    (SURVEY.md Appendix B.2 / B.5)."""
    exe = os.path.join(ROOT, "cuppens")
    out = tmp_path / "out.txt"
    r = subprocess.run([exe, "-p", "2", "-i", os.path.join(ROOT, "tests", "golden", "tinyL.mtx"), "-e", str(out)],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    for line in ("Input file: ", "Program will compute all eigenvectors", "Output file: ", "Number of MPI tasks is: 2",
                 "Start divide phase ...", "Average leaf size will be 2.0", "Apply QR algorithm on leaves ...",
                 "Start Conquer Phase ...", "Required time to compute all eigenvalues: ", "Required time for root finding: ",
                 "Required time for eigenvector extraction from U_i's: ", "Write results to file ...",
                 "Required time for backtransformation: ", "Program finished successfully!"):
        assert line in r.stdout, line
    g = load_golden("tinyL_p2")
    rows = [l.split() for l in out.read_text().splitlines()]
    assert len(rows) == 4 and all(len(x) == 2 for x in rows)
    assert np.abs(np.array([float(x[0]) for x in rows]) - g["lam"]).max() <= 4e-12
    assert max(float(x[1]) for x in rows) <= g["resid"].max() * 1.05 + 4e-15
    # eigenvalues only, scheme input, -eFILE selection
    r = subprocess.run([exe, "-p", "4", "-s", "2", "-n", "256", str(out)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "Use a matrix of scheme 2 with dimension 256" in r.stdout
    lam = np.array([float(l) for l in out.read_text().splitlines()])
    assert np.abs(lam - load_golden("s2_n256_p4")["lam"]).max() <= 4e-12
    ev = tmp_path / "ev.txt"
    ev.write_text("1\n3\nfoo\n999\n")
    r = subprocess.run([exe, "-p", "4", "-s", "2", "-n", "256", "-e" + str(ev), str(out)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.count("WARNING: Line") == 2
    rows = [l.split() for l in out.read_text().splitlines()]
    assert [len(x) for x in rows[:4]] == [2, 1, 2, 1] and all(len(x) == 1 for x in rows[4:])


def test_resolve_is_repeatable(lib, oracle):
    D, E = oracle.goe(900)
    s = se.CuppenSolver(900, ref_leaves=4, lib=lib)
    s.set_tridiagonal(D, E)
    s.solve()                                   # eager
    a = s.eigenvalues().copy(); ra = s.residuals().copy(); Va = s.eigenvectors().copy()
    st = s.merge_stats()
    for it in range(3):                         # CUDA-graph capture, then replays
        s.solve()
        assert np.array_equal(a, s.eigenvalues()) and np.array_equal(ra, s.residuals())
        assert s.merge_stats() == st
    assert np.array_equal(Va, s.eigenvectors())
    # a new matrix on the same handle invalidates the captured graph
    D2, E2 = oracle.goe(900, seed=11)
    s.set_tridiagonal(D2, E2)
    for it in range(3):
        s.solve()
    ref = se.cuppens(D2, E2, ref_leaves=4, lib=lib)
    assert np.array_equal(ref["lam"], s.eigenvalues())
    s.close()


@pytest.mark.parametrize("variant", [0, 1, 2, 3, 4, 5])
@pytest.mark.parametrize("M,N,K", [(128, 128, 16), (1, 1, 1), (300, 200, 70), (77, 129, 0), (1000, 1203, 513), (130, 4000, 33), (2048, 1024, 2048)])
def test_gemm_kernels_against_fma_reference(lib, variant, M, N, K):
    """The back-transformation GEMM kernels on full-mantissa random data with ragged M/N/K and scattered output columns:
    0 cp.async 128x128 (odd row offset), 1 TMA bulk-copy lines (even offset), 2 cp.async 64x64, 3 TMA bulk-copy lines with
    an ODD row offset (lines fetched from one row earlier), 4 / 5 TMA tensor maps with an even / odd offset.
    fp64 tolerance: K * 4 ulp."""
    from symmetric_eigenvalue_b200 import api
    err, _ = api.selftest_gemm(variant, M, N, K, reps=1)
    assert err <= max(K, 1) * 4 * 2.2e-16, (variant, M, N, K, err)


def test_tma_and_cpasync_gemm_paths_agree(lib, oracle):
    """Same decomposition with the tensor-map TMA GEMM (default), the bulk-copy TMA GEMM (CUPPEN_GEMM=bulk) and the cp.async
    GEMM (CUPPEN_GEMM=cpasync); n = 1501 with P = 4 gives reference leaves of 376 / 375 rows: odd row offsets."""
    for n in (1500, 1501):
        D, E = oracle.goe(n)
        a = se.cuppens(D, E, ref_leaves=4, lib=lib)
        for variant in ("bulk", "cpasync"):
            os.environ["CUPPEN_GEMM"] = variant
            try:
                b = se.cuppens(D, E, ref_leaves=4, lib=lib)
            finally:
                del os.environ["CUPPEN_GEMM"]
            assert np.array_equal(a["lam"], b["lam"])
            assert np.abs(a["V"] - b["V"]).max() < 1e-13
            assert np.abs(a["resid"] - b["resid"]).max() < 1e-12


@pytest.mark.parametrize("gen,n,P", [("goe", 4096, 4), ("s2", 4096, 8), ("wilk", 3000, 4)])
def test_tile_split_row_support_and_l2_hint_switches_agree(lib, oracle, gen, n, P):
    """The diagnostic switches of INTEGRATION.md section 5 against the defaults.  Half tiles for a short last GEMM wave
    (GOE n=4096 P=4: 464 tiles at the 2048 level = 3 waves + 20; s2 n=4096: 512 = 3 waves + 68) and the L2 eviction hints
    compute every element with the same K order: bit-identical eigenvectors.  The row supports only skip rows that hold
    exact zeros: the same residuals up to the summation order."""
    D, E = {"goe": oracle.goe, "s2": lambda k: oracle.scheme(2, k), "wilk": lambda k: oracle.wilkinson(k, norm=64.0)}[gen](n)
    a = se.cuppens(D, E, ref_leaves=P, lib=lib)
    for var, val in (("CUPPEN_SPLIT_TAIL", "0"), ("CUPPEN_GEMM_HINT", "1"), ("CUPPEN_SPAN", "0")):
        os.environ[var] = val
        try:
            b = se.cuppens(D, E, ref_leaves=P, lib=lib)
        finally:
            del os.environ[var]
        assert np.array_equal(a["lam"], b["lam"]), var
        assert np.array_equal(a["V"], b["V"]), var
        assert np.allclose(a["resid"], b["resid"], rtol=1e-6, atol=1e-13 * norm_T(D, E)), var


# ---- selected-eigenvector mode (-eFILE) ---------------------------------------------------------------
def sign_invariant_diff(A, B):
    return float(np.minimum(np.abs(A - B).max(axis=0), np.abs(A + B).max(axis=0)).max()) if A.size else 0.0


@pytest.mark.parametrize("gen,n,P", [("goe", 700, 4), ("s1", 1000, 8), ("s2", 512, 8), ("rand_u", 1500, 1), ("wilk", 1001, 4),
                                     ("goe", 3000, 2), ("goe", 1, 1), ("s2", 33, 1)])
def test_select_mode_matches_full_mode(lib, oracle, gen, n, P):
    """cauchy_apply_kernel + chain back-walk against the materialised back-transformation (DMMA GEMMs):
    identical eigenvalues / deflation counts, the same vectors to a few ulp, the same residuals."""
    D, E = {"goe": oracle.goe, "rand_u": oracle.rand_u, "wilk": lambda k: oracle.wilkinson(k, norm=64.0),
            "s1": lambda k: oracle.scheme(1, k), "s2": lambda k: oracle.scheme(2, k)}[gen](n)
    full = se.cuppens(D, E, ref_leaves=P, lib=lib)
    rng = np.random.default_rng(n)
    sel = [0, n - 1] + rng.integers(0, n, size=min(n, 19)).tolist()
    out = se.cuppens(D, E, ref_leaves=P, lib=lib, select=sel)
    # (the solve runs in eigenvalue-only mode: boundary rows by RowGemv instead of being read from the GEMM result)
    assert np.abs(out["lam"] - full["lam"]).max() <= 1e-13 * norm_T(D, E)
    assert ref_stats(out["stats"]) == ref_stats(full["stats"])
    # the same vectors up to sign: a z component that is zero by symmetry but rounds to +-1e-15 (live under the
    # accurate rule) decides the sign of "its" eigenvector, and the two paths feed the merges boundary rows that
    # differ in the last bits (DMMA GEMM result vs RowGemv) -- seen on the Poisson matrix
    assert sign_invariant_diff(out["V"], full["V"][:, sel]) < 1e-12
    assert np.allclose(out["resid"], full["resid"][sel], rtol=1e-3, atol=1e-13 * norm_T(D, E))


def test_select_mode_baseline_size_against_lapack(lib, oracle):
    """n=16384 (BASELINE configs[2] input, accurate rule): 24 selected eigenvectors against LAPACK's
    inverse iteration, without ever forming a 2 GB matrix."""
    from scipy.linalg import eigh_tridiagonal
    n = 16384
    D, E = oracle.goe(n)
    sel = np.unique(np.linspace(0, n - 1, 24).astype(int))
    s = se.CuppenSolver(n, ref_leaves=1, lib=lib, select=True)
    s.set_tridiagonal(D, E)
    s.select(sel)
    for _ in range(3):                              # eager, graph capture, replay: the apply phase follows each
        s.solve()
    lam, V, r = s.eigenvalues(), s.selected_eigenvectors(), s.residuals(sel)
    s.close()
    nT = norm_T(D, E)
    assert r.max() < 1e-14 * nT
    assert np.abs(np.linalg.norm(V, axis=0) - 1).max() < 1e-13
    assert np.abs(V.T @ V - np.eye(len(sel))).max() < 1e-13
    for t, i in enumerate(sel[::6]):
        w, x = eigh_tridiagonal(D, E, select="i", select_range=(int(i), int(i)))
        assert abs(w[0] - lam[i]) < 3e-14 * nT
        x = x[:, 0] * np.sign(x[:, 0] @ V[:, 6 * t])
        assert np.abs(x - V[:, 6 * t]).max() < 1e-9          # eigenvector condition ~ eps*||T||/gap, gap ~ 2/n
    # and the reference-rule tree: residuals at the level the reference's 1e-6 deflation allows
    g = load_golden("goe_n4096_p8")
    sel = list(range(0, 4096, 256))
    out = se.cuppens(g["D"], g["E"], ref_leaves=8, lib=lib, select=sel)
    check_against_golden(g, out, False)
    full = se.cuppens(g["D"], g["E"], ref_leaves=8, lib=lib)
    assert sign_invariant_diff(out["V"], full["V"][:, sel]) < 1e-12


# ---- on-GPU orthogonality check (gram_check_kernel) and eigenvector output ------------------------------
@pytest.mark.parametrize("gen,n,P", [("goe", 1001, 4), ("s1", 4096, 8), ("s2", 130, 1), ("wilk", 2049, 2), ("goe", 5, 1)])
def test_orthogonality_kernel_against_numpy(lib, oracle, gen, n, P):
    D, E = {"goe": oracle.goe, "wilk": lambda k: oracle.wilkinson(k, norm=64.0),
            "s1": lambda k: oracle.scheme(1, k), "s2": lambda k: oracle.scheme(2, k)}[gen](n)
    s = se.CuppenSolver(n, ref_leaves=P, lib=lib)
    s.set_tridiagonal(D, E)
    s.solve()
    dev, _ = s.orthogonality()                      # storage order
    V = s.eigenvectors()
    dev2, _ = s.orthogonality()                     # after the sorted gather
    want = np.abs(V.T @ V - np.eye(n)).max()
    s.close()
    assert abs(dev - want) <= 64 * 2.2e-16 and abs(dev2 - want) <= 64 * 2.2e-16, (dev, dev2, want)
    assert dev < 1e-12


def test_orthogonality_at_baseline_size(lib, oracle):
    """BASELINE configs[2]/[3] size: ||V^T V - I|| of the n=16384 decompositions without moving V off the GPU."""
    for gen in ("goe", "wilk"):
        D, E = (oracle.goe(16384) if gen == "goe" else oracle.wilkinson(16384, norm=64.0))
        s = se.CuppenSolver(16384, ref_leaves=8, lib=lib)
        s.set_tridiagonal(D, E)
        s.solve()
        dev, sec = s.orthogonality()
        r = s.residuals()
        s.close()
        assert dev < 1e-12, (gen, dev)              # the reference: 1e-8 ... 1e-10 (SURVEY.md finding 4)
        assert r.max() < 1e-5 * norm_T(D, E)


def test_cli_eigenvector_file_and_orthogonality(lib, tmp_path):
    exe = os.path.join(ROOT, "cuppens")
    out, vf, ev = tmp_path / "out.txt", tmp_path / "v.bin", tmp_path / "ev.txt"
    r = subprocess.run([exe, "-p", "4", "-s", "1", "-n", "512", "-e", "-c", "-v", str(vf), str(out)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert "Eigenvectors written to: " in r.stdout and "Orthogonality max|V^T V - I|: " in r.stdout
    dev = float(r.stdout.split("Orthogonality max|V^T V - I|: ")[1].split()[0])
    ranks, lam, V = se.read_eigenvector_file(str(vf))
    assert V.shape == (512, 512) and abs(np.abs(V.T @ V - np.eye(512)).max() - dev) < 1e-14
    rows = np.array([[float(x) for x in l.split()] for l in out.read_text().splitlines()])
    assert np.array_equal(rows[:, 0], lam)
    D, E = se.createMatrixScheme1(512, lib=lib)
    TV = D[:, None] * V
    TV[1:] += E[:, None] * V[:-1]
    TV[:-1] += E[:, None] * V[1:]
    assert np.allclose(np.linalg.norm(TV - V * lam[None, :], axis=0), rows[:, 1], rtol=1e-6, atol=1e-13)
    ev.write_text("512\n7\n")
    r = subprocess.run([exe, "-p", "4", "-s", "1", "-n", "512", "-e" + str(ev), "-v", str(vf), str(out)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    ranks2, lam2, V2 = se.read_eigenvector_file(str(vf))                     # selected-eigenvector mode: 2 <= 512/16
    assert ranks2.tolist() == [6, 511] and sign_invariant_diff(V2, V[:, [6, 511]]) < 1e-12


@pytest.mark.parametrize("n,g0,l0,cnt", [(4096, 0, 0, 4096), (4097, 0, 0, 4097), (1000, 0, 0, 1000), (5000, 1250, 0, 1250), (5000, 1251, 313, 1249),
                                         (5000, 3750, 7, 1250), (5000, 4999, 0, 1), (3, 0, 0, 3), (1, 0, 0, 1), (2049, 1024, 512, 1025), (700, 1, 1, 698)])
def test_residual_kernel_slices(lib, n, g0, l0, cnt):
    """residual_kernel (paired rows, neighbour shuffles, halo rows) on every slice shape of the multi-GPU layout:
    odd / even first rows, odd local offsets (scalar-load path), single rows, slices touching row 0 and n-1."""
    from symmetric_eigenvalue_b200 import api
    # relative error of the sum of squares against a plain per-column loop on full-mantissa random data: a row's
    # y = (d - lambda) x + e x' + e x'' may cancel, and the two sides contract their FMAs differently
    for variant in (0, 14, 83):
        assert api.selftest_residual(n, g0, l0, cnt, variant=variant)[0] < 1e-10


def test_fast_reciprocal_accuracy(lib):
    """The reciprocal of the Cauchy-like inner loops (platform.h: hardware seed + one cubic step) against the correctly
    rounded 1/x on 16 M random operands (random mantissas, both signs, binary exponents -500..500): the seed carries
    about 20 bits, so the cubic step leaves e^3 ~ 2^-60 before the final rounding -- within one ulp, like the two-Newton-
    step form it replaced."""
    from symmetric_eigenvalue_b200 import api
    seed_err, newton_ulp, cubic_ulp = api.selftest_rcp(1 << 24)
    assert seed_err < 2.0 ** -18, seed_err
    assert newton_ulp <= 1.0, newton_ulp
    assert cubic_ulp <= 1.0, cubic_ulp


@pytest.mark.parametrize("name", ["s1_n4096_p8_sel", "goe_n4096_p8_sel"])
def test_select_mode_against_reference_efile_goldens(lib, name):
    """Selected-eigenvector mode against the reference's own `-eFILE` output at n=4096, P=8."""
    g = load_golden(name)
    out = se.cuppens(g["D"], g["E"], ref_leaves=g["P"], lib=lib, select=(g["sel"] - 1).tolist())
    check_select_against_efile_golden(g, out)


def test_eigenvector_columns(lib, oracle):
    """cuppen_copy_eigenvector_columns (gather_sel_cols_kernel) against the full copy, before and after the sorted gather."""
    D, E = oracle.goe(1300)
    s = se.CuppenSolver(1300, ref_leaves=4, lib=lib)
    s.set_tridiagonal(D, E)
    s.solve()
    idx = [1299, 0, 7, 7, 650]
    a = s.eigenvector_columns(idx)
    V = s.eigenvectors()
    b = s.eigenvector_columns(idx)
    s.close()
    assert np.array_equal(a, V[:, idx]) and np.array_equal(b, V[:, idx])
