"""CPU-side checks (no GPU): the C-ABI library loads and exports every declared symbol, the host
I/O mirrors the reference, the CLI's option handling / exit codes, and -- through the TEST-ONLY host
build of the product sources -- the tree planning, deflation bookkeeping and numerics against the
oracle and the reference's golden outputs."""
import os
import re
import subprocess

import numpy as np
import pytest

import symmetric_eigenvalue_b200 as se
from symmetric_eigenvalue_b200 import api
from conftest import ROOT, check_against_golden, check_select_against_efile_golden, golden_cases, load_golden, norm_T, ref_stats


# ---- the shipped library -----------------------------------------------------------------------------
def test_library_exports_every_declared_symbol(product_lib):
    hdr = open(os.path.join(ROOT, "include", "cuppen_b200.h")).read()
    declared = set(re.findall(r"\b(cuppen_[a-z_0-9]+)\s*\(", hdr))
    declared -= {"cuppen_handle_s"}
    assert declared == set(api.EXPORTED_SYMBOLS), declared ^ set(api.EXPORTED_SYMBOLS)
    for name in declared:
        assert hasattr(product_lib, name), name


def test_no_cpu_fallback_in_product(product_lib):
    """Without a GPU the product must fail loudly, not compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(se.CuppenError) as ei:
        se.CuppenSolver(16, lib=product_lib)
    assert ei.value.code == -10
    syms = subprocess.run(["nm", "-D", "--defined-only", api.library_path()], capture_output=True, text=True).stdout
    assert "host_leaf_ql" not in syms and "gemm_host" not in syms and "oracle" not in syms


def test_schemes(product_lib):
    D, E = se.createMatrixScheme1(5, lib=product_lib)
    assert np.allclose(D, 1 + np.arange(5) * 99 / 4) and (E == -1).all()
    D, E = se.createMatrixScheme2(5, lib=product_lib)
    assert (D == 2).all() and (E == -1).all()


def test_mtx_reader(product_lib, tmp_path, capfd):
    p = tmp_path / "t.mtx"
    p.write_text("%%MatrixMarket matrix coordinate real general\n%c\n3 3 7\n1 1 2\n2 1 -1\n1 2 -1\n2 2 3\n3 2 -4\n2 3 -4\n3 3 5\n")
    D, E = se.readSymmTriadiagonalMatrixFromSparseMTX(str(p), lib=product_lib)
    assert D.tolist() == [2, 3, 5] and E.tolist() == [-1, -4]
    bad = {
        "%%MatrixMarket matrix coordinate real symmetric\n2 2 1\n1 1 1\n": "does not support",
        "%MatrixMarket matrix coordinate real general\n2 2 1\n1 1 1\n": "Could not process Matrix Market banner",
        "%%MatrixMarket matrix coordinate real general\n2 3 1\n1 1 1\n": "Matrix is not square",
        "%%MatrixMarket matrix coordinate real general\n3 3 1\n1 3 1\n": "Matrix is not tridiagonal",
        "%%MatrixMarket matrix coordinate real general\n2 2 4\n1 1 1\n2 1 5\n1 2 6\n2 2 1\n": "Matrix is not symmetric",
        "%%MatrixMarket matrix coordinate real general\n2 2 4\n1 1 1\n1 2 5\n2 1 5\n2 2 1\n": "Matrix is not symmetric",
    }
    for text, msg in bad.items():
        p.write_text(text)
        with pytest.raises(se.CuppenError) as ei:
            se.readSymmTriadiagonalMatrixFromSparseMTX(str(p), lib=product_lib)
        assert ei.value.code == -2
        assert msg in capfd.readouterr().out
    with pytest.raises(se.CuppenError):
        se.readSymmTriadiagonalMatrixFromSparseMTX(str(tmp_path / "missing.mtx"), lib=product_lib)


def test_ev_file_and_writer(product_lib, tmp_path, capfd):
    ev = tmp_path / "ev.txt"
    ev.write_text("3\n1\nfoo\n9\n1\n")
    idx = se.determineEigenvectorsToCompute(str(ev), 4, lib=product_lib)
    assert idx.tolist() == [0, 0, 2]
    assert capfd.readouterr().out.count("WARNING: Line") == 2
    out = tmp_path / "o.txt"
    lam = np.array([0.3819660112501050975, 1.38196601125010532, 2.5, 3.5])
    res = np.array([2.92423134397388715e-16, 1e-15, 2e-15, 3e-15])
    se.writeResults(str(out), lam, res, indices=idx, lib=product_lib)
    lines = out.read_text().splitlines()
    assert lines[0] == "0.3819660112501050975 2.92423134397388715e-16"
    assert lines[1] == " 1.38196601125010532"
    assert len(lines[2].split()) == 2 and len(lines[3].split()) == 1
    se.writeResults(str(out), lam, lib=product_lib)
    assert all(len(l.split()) == 1 for l in out.read_text().splitlines())


def test_cli_usage_and_exit_codes(product_lib, tmp_path):
    exe = os.path.join(ROOT, "cuppens")
    assert os.path.exists(exe)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "USAGE cuppens [options] [outputfile]" in r.stdout
    r = subprocess.run([exe, "-h"], capture_output=True, text=True)
    assert r.returncode == 0 and " -e(FILENAME)" in r.stdout
    for args, msg in ((["-s", "3"], "Invalid argument for option -s. See help."),
                      (["-n", "0"], "Invalid argument for option -n. See help."),
                      (["-x"], "Unknown option `-x'."),
                      (["a", "b"], "Invalid number of positional arguments. See help."),
                      (["-g", "0"], "Invalid argument for option -g. See help.")):
        r = subprocess.run([exe] + args, capture_output=True, text=True)
        assert r.returncode == 1 and msg in r.stderr, (args, r.returncode, r.stderr)
    r = subprocess.run([exe, "-i", str(tmp_path / "nope.mtx")], capture_output=True, text=True)
    assert r.returncode == 2 and "Could not open file" in r.stderr
    r = subprocess.run([exe, "-p", "8", "-n", "4", "-s", "2"], capture_output=True, text=True)
    assert r.returncode == 4 and "Leaf Size is too small! Reduce number of tasks." in r.stderr


# ---- product sources on the host (TEST-ONLY build): orchestration + numerics vs oracle / goldens -------
@pytest.mark.parametrize("name", [c for c in golden_cases() if "4096" not in c and "16384" not in c])
def test_hostemu_matches_reference_goldens(hostemu, name):
    g = load_golden(name)
    vec = bool(np.isfinite(g["resid"]).any())
    out = se.cuppens(g["D"], g["E"], ref_leaves=g["P"], vectors=vec, lib=hostemu)
    check_against_golden(g, out, vec)


@pytest.mark.parametrize("name", ["goe_n4096_p8", "s2_n16384_p8", "wilk64_n16384_p8", "randu_n16384_p8"])
def test_hostemu_baseline_size_goldens_eigenvalue_mode(hostemu, name):
    """BASELINE-size inputs (configs 3/4) against the reference's own output, eigenvalue-only path
    (boundary-row propagation, src/main.c:613-639): the reference is off from the true spectrum by
    1e-6 here, we must land on the reference's answer."""
    g = load_golden(name)
    out = se.cuppens(g["D"], g["E"], ref_leaves=g["P"], vectors=False, lib=hostemu)
    check_against_golden(g, out, False)


@pytest.mark.parametrize("vectors", [True, False])
def test_hostemu_eigenvalue_only_mode_agrees(hostemu, oracle, vectors):
    D, E = oracle.goe(300)
    a = se.cuppens(D, E, ref_leaves=4, vectors=vectors, lib=hostemu)
    b = se.cuppens(D, E, ref_leaves=4, vectors=True, lib=hostemu)
    assert np.abs(a["lam"] - b["lam"]).max() < 1e-14
    assert ref_stats(a["stats"]) == ref_stats(b["stats"])


@pytest.mark.parametrize("gen,n", [("goe", 500), ("rand_u", 333), ("wilkinson", 257), ("s1", 1000), ("s2", 640)])
def test_hostemu_accurate_mode(hostemu, oracle, gen, n):
    """ref_leaves=1: LAPACK-grade tolerances on every level (stands in for dsteqr)."""
    from scipy.linalg import eigvalsh_tridiagonal
    D, E = {"goe": oracle.goe, "rand_u": oracle.rand_u, "wilkinson": oracle.wilkinson,
            "s1": lambda k: oracle.scheme(1, k), "s2": lambda k: oracle.scheme(2, k)}[gen](n)
    out = se.cuppens(D, E, ref_leaves=1, lib=hostemu)
    nT = norm_T(D, E)
    assert np.abs(out["lam"] - eigvalsh_tridiagonal(D, E)).max() < 2e-14 * nT
    assert out["resid"].max() < 5e-15 * nT
    V = out["V"]
    assert np.abs(V.T @ V - np.eye(n)).max() < 5e-14


@pytest.mark.parametrize("name", __import__("families").FAMILIES)
def test_hostemu_matrix_families(hostemu, name):
    """Graded, glued, clustered, badly scaled ... inputs under the accurate rule, against LAPACK."""
    import families
    D, E = families.family(name, 200)
    families.check_accurate(se.cuppens(D, E, ref_leaves=1, lib=hostemu), D, E)


def test_hostemu_orthogonality_beats_reference(hostemu, oracle):
    D, E = oracle.goe(256)
    o = oracle.solve(D, E, 4, vectors=True)
    out = se.cuppens(D, E, ref_leaves=4, lib=hostemu)
    mine = np.abs(out["V"].T @ out["V"] - np.eye(256)).max()
    ref = np.abs(o["V"].T @ o["V"] - np.eye(256)).max()
    assert mine <= ref and mine < 1e-13
    assert out["resid"].max() <= o["resid"].max() * 1.05


def test_hostemu_edge_cases(hostemu):
    out = se.cuppens(np.array([3.0]), np.zeros(0), lib=hostemu)
    assert out["lam"].tolist() == [3.0]
    out = se.cuppens(np.array([2.0, 2.0]), np.array([-1.0]), ref_leaves=2, lib=hostemu)
    assert np.allclose(out["lam"], [1.0, 3.0], atol=1e-14)
    with pytest.raises(se.CuppenError) as ei:
        se.CuppenSolver(3, ref_leaves=4, lib=hostemu)
    assert ei.value.code == -4
    # zero off-diagonal away from the reference splits is fine (accurate rule), at a reference split it is an error
    D = np.arange(1.0, 65.0); E = np.ones(63); E[10] = 0.0
    out = se.cuppens(D, E, ref_leaves=1, lib=hostemu)
    from scipy.linalg import eigvalsh_tridiagonal
    assert np.abs(out["lam"] - eigvalsh_tridiagonal(D, E)).max() < 1e-12
    E = np.ones(63); E[31] = 0.0
    s = se.CuppenSolver(64, ref_leaves=2, lib=hostemu)
    with pytest.raises(se.CuppenError) as ei:
        s.set_tridiagonal(D, E)
    assert ei.value.code == -3


# ---- selected-eigenvector mode (-eFILE): back-application of the implicit U factors ---------------------
@pytest.mark.parametrize("gen,n,P", [("goe", 300, 4), ("s1", 500, 8), ("s2", 256, 4), ("rand_u", 333, 1), ("wilk", 257, 2),
                                     ("goe", 40, 1), ("s2", 1000, 3), ("goe", 1, 1)])
def test_hostemu_select_mode_matches_full_mode(hostemu, oracle, gen, n, P):
    """The selected columns must be the same vectors the full back-transformation produces (same formulas,
    different association order: a few ulp), with the same residuals, and the eigenvalue path must not change."""
    D, E = {"goe": oracle.goe, "rand_u": oracle.rand_u, "wilk": lambda k: oracle.wilkinson(k, norm=64.0),
            "s1": lambda k: oracle.scheme(1, k), "s2": lambda k: oracle.scheme(2, k)}[gen](n)
    full = se.cuppens(D, E, ref_leaves=P, lib=hostemu)
    rng = np.random.default_rng(n)
    sel = [0, n - 1] + rng.integers(0, n, size=min(n, 19)).tolist()          # unsorted, duplicates, > SEL_NV: three passes
    out = se.cuppens(D, E, ref_leaves=P, lib=hostemu, select=sel)
    assert np.array_equal(out["lam"], full["lam"])
    assert ref_stats(out["stats"]) == ref_stats(full["stats"])
    assert out["V"].shape == (n, len(sel))
    assert np.abs(out["V"] - full["V"][:, sel]).max() < 5e-14
    assert np.allclose(out["resid"], full["resid"][sel], rtol=1e-3, atol=1e-13 * norm_T(D, E))


def test_hostemu_select_mode_api(hostemu, oracle):
    D, E = oracle.goe(200)
    s = se.CuppenSolver(200, ref_leaves=2, lib=hostemu, select=True)
    s.set_tridiagonal(D, E)
    s.solve()                                        # nothing selected: eigenvalues only
    lam = s.eigenvalues()
    assert s.selected_eigenvectors().shape == (200, 0)
    with pytest.raises(se.CuppenError) as ei:
        s.select([200])
    assert ei.value.code == -1
    s.select([5, 7])
    with pytest.raises(se.CuppenError) as ei:       # a new selection needs a new solve
        s.eigenvalues()
    assert ei.value.code == -5
    s.solve()
    assert np.array_equal(lam, s.eigenvalues())
    r = s.residuals()
    assert np.isfinite(r[[5, 7]]).all() and np.isnan(np.delete(r, [5, 7])).all()
    assert np.array_equal(s.residuals([7, 5]), r[[7, 5]])
    with pytest.raises(se.CuppenError) as ei:
        s.residuals([6])
    assert ei.value.code == -1
    V = s.selected_eigenvectors()
    T = np.diag(D) + np.diag(E, 1) + np.diag(E, -1)
    assert np.allclose(np.linalg.norm(T @ V - V * lam[[5, 7]], axis=0), r[[5, 7]], rtol=1e-6, atol=1e-14)
    s.close()
    full = se.CuppenSolver(200, ref_leaves=2, lib=hostemu)
    with pytest.raises(se.CuppenError) as ei:       # not a select handle
        full.select([1])
    assert ei.value.code == -5
    full.close()
    h = __import__("ctypes").c_void_p()
    assert hostemu.cuppen_create(__import__("ctypes").byref(h), 64, 1, api.FLAG_VECTORS | api.FLAG_SELECT, 0) == -1


def test_hostemu_orthogonality_and_eigenvector_file(hostemu, oracle, tmp_path):
    D, E = oracle.goe(150)
    s = se.CuppenSolver(150, ref_leaves=2, lib=hostemu)
    s.set_tridiagonal(D, E)
    s.solve()
    V = s.eigenvectors()
    dev, _ = s.orthogonality()
    assert abs(dev - np.abs(V.T @ V - np.eye(150)).max()) < 1e-15 and dev < 1e-13
    f = tmp_path / "v.bin"
    s.write_eigenvectors(str(f))
    ranks, lam, W = se.read_eigenvector_file(str(f))
    assert ranks.tolist() == list(range(150)) and np.array_equal(lam, s.eigenvalues()) and np.array_equal(W, V)
    s.close()
    s = se.CuppenSolver(150, ref_leaves=2, lib=hostemu, select=True)
    s.set_tridiagonal(D, E)
    s.select([149, 3])
    s.solve()
    s.write_eigenvectors(str(f))
    ranks, lam, W = se.read_eigenvector_file(str(f))
    assert ranks.tolist() == [149, 3] and np.array_equal(lam, s.eigenvalues()[[149, 3]])
    assert np.array_equal(W, s.selected_eigenvectors()) and np.abs(W - V[:, [149, 3]]).max() < 1e-13
    with pytest.raises(se.CuppenError):
        s.orthogonality()
    s.close()


def test_hostemu_select_mode_baseline_config(hostemu):
    """BASELINE configs[1] (`-s 1 -n 4096`, reference tree P=8) with `-eFILE`-style selection: eigenvalues and
    deflation counts as the reference printed them, residuals of the selected vectors at the level the
    reference's 1e-6 z-deflation allows (golden s1_n1024_p4: 1.3e-6)."""
    g = load_golden("s1_n4096_p8")
    sel = list(range(0, 4096, 256)) + [4095]
    out = se.cuppens(g["D"], g["E"], ref_leaves=8, lib=hostemu, select=sel)
    check_against_golden(g, out, False)
    assert out["V"].shape == (4096, len(sel)) and np.abs(np.linalg.norm(out["V"], axis=0) - 1).max() < 1e-12
    assert out["resid"].max() < 2e-6
    D, E, V, lam = g["D"], g["E"], out["V"], out["lam"][sel]
    TV = D[:, None] * V
    TV[1:] += E[:, None] * V[:-1]
    TV[:-1] += E[:, None] * V[1:]
    assert np.allclose(np.linalg.norm(TV - V * lam[None, :], axis=0), out["resid"], rtol=1e-6, atol=1e-13)


def test_cli_help_lists_every_option(product_lib):
    exe = os.path.join(ROOT, "cuppens")
    r = subprocess.run([exe, "-h"], capture_output=True, text=True)
    for opt in (" -h", " -i FILENAME", " -s NUM", " -n NUM", " -e(FILENAME)", " -p NUM", " -g NUM", " -v FILENAME", " -c"):
        assert opt + "\n" in r.stdout, opt


@pytest.mark.parametrize("name", ["s1_n4096_p8_sel", "goe_n4096_p8_sel"])
def test_hostemu_select_mode_against_reference_efile_goldens(hostemu, name):
    """The reference's own `-eFILE` output at n=4096, P=8 (tests/golden/make_golden.py, ~4 s per vector there)."""
    g = load_golden(name)
    out = se.cuppens(g["D"], g["E"], ref_leaves=g["P"], lib=hostemu, select=(g["sel"] - 1).tolist())
    check_select_against_efile_golden(g, out)


@pytest.mark.parametrize("name", __import__("families").FAMILIES)
def test_hostemu_select_mode_matrix_families(hostemu, name):
    """The coefficient-space back-application on graded, glued, clustered, badly scaled ... inputs (accurate rule):
    residuals, norms and mutual orthogonality of the selected vectors at working precision."""
    import families
    D, E = families.family(name, 200)
    n = len(D)
    nT = np.abs(D).max() + 2 * np.abs(E).max()
    sel = list(range(0, n, 7)) + [n - 1]
    out = se.cuppens(D, E, ref_leaves=1, lib=hostemu, select=sel)
    from scipy.linalg import eigh_tridiagonal
    assert np.abs(out["lam"] - eigh_tridiagonal(D, E, eigvals_only=True)).max() <= 5e-14 * nT
    V = out["V"]
    assert out["resid"].max() <= 5e-14 * nT
    assert np.abs(V.T @ V - np.eye(len(sel))).max() <= 5e-13


def test_hostemu_handle_reuse_with_new_matrices(hostemu, oracle):
    """A handle keeps its divide tree (the shape depends on (n, P) only); every cuppen_set_tridiagonal re-runs only the
    divide pass (theta rule, rho, modified diagonal).  Results must equal those of a fresh handle, in any order."""
    n, P = 300, 4
    mats = [oracle.goe(n), oracle.scheme(1, n), oracle.rand_u(n), oracle.goe(n, seed=3)]
    fresh = [se.cuppens(D, E, ref_leaves=P, lib=hostemu) for D, E in mats]
    s = se.CuppenSolver(n, ref_leaves=P, lib=hostemu)
    for k in (0, 1, 2, 3, 1, 0):
        s.set_tridiagonal(*mats[k])
        s.solve()
        assert np.array_equal(s.eigenvalues(), fresh[k]["lam"])
        assert np.array_equal(s.residuals(), fresh[k]["resid"])
        assert s.merge_stats() == fresh[k]["stats"]
        t = s.timers()
        assert t["kernel_launches"] > 0 and t["total_s"] > 0
    s.close()


def test_hostemu_rejects_non_finite_input_and_failed_set_leaves_no_result(hostemu, oracle):
    """NaN/Inf entries must be refused before they reach the sort (they corrupted the index lists), and a
    failed cuppen_set_tridiagonal must not leave the previous matrix's decomposition looking valid."""
    n = 200
    D, E = oracle.goe(n)
    s = se.CuppenSolver(n, ref_leaves=2, lib=hostemu)
    s.set_tridiagonal(D, E)
    s.solve()
    lam = s.eigenvalues()
    for bad_d, bad_e in ((5, None), (None, 17)):
        D2, E2 = D.copy(), E.copy()
        if bad_d is not None:
            D2[bad_d] = np.nan
        if bad_e is not None:
            E2[bad_e] = np.inf
        with pytest.raises(se.CuppenError) as ei:
            s.set_tridiagonal(D2, E2)
        assert ei.value.code == -1
        with pytest.raises(se.CuppenError) as ei:          # no matrix: solve and results are refused
            s.solve()
        assert ei.value.code == -5
        with pytest.raises(se.CuppenError) as ei:
            s.eigenvalues()
        assert ei.value.code == -5
    E3 = E.copy(); E3[n // 2 - 1] = 0.0                      # zero at the reference split: CUPPEN_ERR_ZERO, same state rule
    with pytest.raises(se.CuppenError) as ei:
        s.set_tridiagonal(D, E3)
    assert ei.value.code == -3
    with pytest.raises(se.CuppenError):
        s.solve()
    s.set_tridiagonal(D, E)
    s.solve()
    assert np.array_equal(lam, s.eigenvalues())
    s.close()


@pytest.mark.parametrize("expo", [-300, -160, -120, 0, 150, 290])
def test_hostemu_power_of_two_scaling_is_exact(hostemu, oracle, expo):
    """Matrices of extreme norm are scaled by a power of two inside the library, as dstedc does with dlascl (ADVICE
    r01: 1e-140*T gave max|VtV-I| = 1 unscaled).  Under the accurate rule (ref_leaves=1) scaling T by 2^e must scale
    eigenvalues and residuals by exactly 2^e and leave the vectors and the deflation counts alone.  (The reference
    rule is not scale invariant by construction -- absolute thresholds, theta = 1000*beta -- and is kept in the
    caller's units: plan_divide(inv_scale), MergeDesc::dthr.)"""
    n = 200
    D, E = oracle.goe(n)
    base = se.cuppens(D, E, ref_leaves=1, lib=hostemu)
    f = 2.0 ** round(expo * np.log2(10.0))
    out = se.cuppens(D * f, E * f, ref_leaves=1, lib=hostemu)
    assert np.array_equal(out["lam"], base["lam"] * f)
    assert np.array_equal(out["V"], base["V"])
    assert np.allclose(out["resid"], base["resid"] * f, rtol=1e-12, atol=0)
    assert [(x.m, x.offset, x.zdefl, x.givens) for x in out["stats"]] == [(x.m, x.offset, x.zdefl, x.givens) for x in base["stats"]]


def test_cli_rejects_non_finite_and_missing_entries(product_lib, tmp_path):
    """an .mtx file that omits a sub-diagonal entry (the reader pre-filled E with NaN) or holds a literal nan
    must stop at the CLI's input checks (src/main.c:196-200 asserts on zero entries) instead of reaching the solver."""
    exe = os.path.join(ROOT, "cuppens")
    p = tmp_path / "gap.mtx"
    # 3 x 3, the (3,2)/(2,3) pair is missing: that entry of the matrix is zero
    p.write_text("%%MatrixMarket matrix coordinate real general\n3 3 5\n1 1 2\n2 1 -1\n1 2 -1\n2 2 2\n3 3 2\n")
    D, E = se.readSymmTriadiagonalMatrixFromSparseMTX(str(p), lib=product_lib)
    assert E.tolist() == [-1.0, 0.0]
    r = subprocess.run([exe, "-i", str(p)], capture_output=True, text=True)
    assert r.returncode != 0 and "Assertion `E[i] != 0' failed" in r.stderr
    q = tmp_path / "nan.mtx"
    q.write_text("%%MatrixMarket matrix coordinate real general\n2 2 4\n1 1 nan\n2 1 -1\n1 2 -1\n2 2 2\n")
    r = subprocess.run([exe, "-i", str(q)], capture_output=True, text=True)
    assert r.returncode == 2 and "not finite" in r.stdout


def test_hostemu_eigenvector_columns(hostemu, oracle):
    """cuppen_copy_eigenvector_columns: the listed columns (ascending-lambda ranks, any order, duplicates) of V, before and
    after the sorted copy has been materialised."""
    D, E = oracle.goe(180)
    s = se.CuppenSolver(180, ref_leaves=2, lib=hostemu)
    s.set_tridiagonal(D, E)
    s.solve()
    idx = [179, 0, 5, 5, 90]
    a = s.eigenvector_columns(idx)              # storage order + permutation
    V = s.eigenvectors()                        # materialises the sorted copy
    b = s.eigenvector_columns(idx)
    assert np.array_equal(a, V[:, idx]) and np.array_equal(b, V[:, idx])
    with pytest.raises(se.CuppenError) as ei:
        s.eigenvector_columns([180])
    assert ei.value.code == -1
    s.close()
