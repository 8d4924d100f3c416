"""The oracle is only trusted after it is pinned: tinyL goldens, the scheme-2
closed form, and outputs of the reference itself (tests/golden/*.npz, made by make_golden.py)."""
import numpy as np
import pytest

from conftest import golden_cases, load_golden, norm_T


def test_tinyL_closed_form(oracle):
    D = np.full(4, 2.0); E = np.full(3, -1.0)
    true = 2 - 2 * np.cos(np.arange(1, 5) * np.pi / 5)
    for P in (1, 2, 4):
        r = oracle.solve(D, E, P)
        assert np.abs(r["lam"] - true).max() < 1e-14
        assert r["resid"].max() < 1e-14


def test_tinyL_reference_digits(oracle):
    """the reference's own printed digits for P=2 and P=4 (bit-exact,
    both runs contain real merges; P=1 is pure dsteqr and differs in the last ulp)."""
    D = np.full(4, 2.0); E = np.full(3, -1.0)
    want = {2: ["0.3819660112501024329", "1.38196601125010532", "2.618033988749898455", "3.618033988749895791"],
            4: ["0.3819660112501079841", "1.381966011250109316", "2.618033988749899343", "3.618033988749894014"]}
    for P, digits in want.items():
        r = oracle.solve(D, E, P)
        got = ["%.19g" % x for x in r["lam"]]
        for a, b in zip(got, digits):
            assert abs(float(a) - float(b)) <= 2e-15, (P, a, b)


def test_scheme2_closed_form(oracle):
    n = 200
    D, E = oracle.scheme(2, n)
    true = np.sort(2 + 2 * np.cos(np.pi * np.arange(1, n + 1) / (n + 1)))   # src/helper.c:52-62
    r = oracle.solve(D, E, 2)
    assert np.abs(r["lam"] - true).max() < 1e-12


@pytest.mark.parametrize("name", golden_cases())
def test_oracle_matches_reference_output(oracle, name):
    g = load_golden(name)
    if len(g["D"]) > 1100:
        pytest.skip("covered by the smaller cases; keeps the CPU suite short")
    vec = bool(np.isfinite(g["resid"]).any())
    r = oracle.solve(g["D"], g["E"], g["P"], residuals=vec)
    assert np.abs(r["lam"] - g["lam"]).max() <= 1e-12 * norm_T(g["D"], g["E"])
    got = sorted((int(m), int(o), int(zd), int(gv)) for (o, m, zd, gv) in r["stats"].tolist())
    assert got == [tuple(int(x) for x in row) for row in g["merges"].tolist()]
    if vec:
        # rounding-level residuals differ by summation order; the large ones (set by the 1e-6
        # z-deflation) must agree closely
        big = g["resid"] > 1e-10
        assert np.allclose(r["resid"][big], g["resid"][big], rtol=1e-3)
        assert r["resid"].max() <= 2 * g["resid"].max() + 1e-13 and g["resid"].max() <= 2 * r["resid"].max() + 1e-13


def test_leaf_ql_against_lapack(oracle):
    from scipy.linalg import eigh_tridiagonal
    rng = np.random.default_rng(3)
    d = rng.normal(size=40); e = rng.normal(size=39)
    r = oracle.solve(d, e, 1, vectors=True)
    w = eigh_tridiagonal(d, e, eigvals_only=True)
    assert np.abs(r["lam"] - w).max() < 1e-13
    V = r["V"]
    assert np.abs(V.T @ V - np.eye(40)).max() < 1e-13


def test_leaf_size_too_small(oracle):
    with pytest.raises(RuntimeError):
        oracle.solve(np.ones(3), np.ones(2), 4)
