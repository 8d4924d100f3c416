"""Matrix families for the accuracy stress tests (accurate rule, ref_leaves=1), shared by the CPU
(host build) and GPU suites."""
import numpy as np


def family(name, n, seed=5):
    rng = np.random.default_rng(seed)
    if name == "normal":
        return rng.normal(size=n), rng.normal(size=n - 1)
    if name == "graded":
        return 10.0 ** np.linspace(0, -12, n), 10.0 ** np.linspace(0, -12, n - 1)
    if name == "graded_rev":
        return 10.0 ** np.linspace(-12, 0, n), 10.0 ** np.linspace(-12, 0, n - 1)
    if name == "huge_norm":
        return 1e8 * rng.normal(size=n), 1e8 * rng.normal(size=n - 1)
    if name == "tiny_norm":
        return 1e-8 * rng.normal(size=n), 1e-8 * rng.normal(size=n - 1)
    if name == "extreme_tiny":       # dstedc scales T to unit norm (dlascl); unscaled, z^2*rho products underflow
        return 1e-160 * rng.normal(size=n), 1e-160 * rng.normal(size=n - 1)
    if name == "extreme_huge":
        return 1e160 * rng.normal(size=n), 1e160 * rng.normal(size=n - 1)
    if name == "tiny_offdiag":
        return rng.normal(size=n), 1e-10 * rng.normal(size=n - 1)
    if name == "const_diag":
        return np.ones(n), 0.5 * np.ones(n - 1)
    if name == "zero_diag":
        return np.zeros(n), np.ones(n - 1)
    if name == "clustered":
        return 1 + 1e-9 * rng.normal(size=n), 1e-9 * rng.normal(size=n - 1)
    if name == "wilkinson":
        return np.abs(np.arange(n) - (n - 1) / 2.0), np.ones(n - 1)
    if name == "glued_wilkinson":
        k = 21
        blocks = max(n // k, 2)
        d = np.tile(np.abs(np.arange(k) - (k - 1) / 2.0), blocks)
        e = np.ones(len(d) - 1)
        e[k - 1::k] = 1e-8
        return d, e
    if name == "negative":
        return -np.abs(rng.normal(size=n)) - 1, rng.normal(size=n - 1)
    if name == "some_zero_E":
        return rng.normal(size=n), rng.normal(size=n - 1) * (rng.uniform(size=n - 1) > 0.2)
    if name == "laguerre":
        return 2 * np.arange(n) + 1.0, -np.arange(1, n) * 1.0
    raise KeyError(name)


FAMILIES = ["normal", "graded", "graded_rev", "huge_norm", "tiny_norm", "extreme_tiny", "extreme_huge", "tiny_offdiag", "const_diag", "zero_diag",
            "clustered", "wilkinson", "glued_wilkinson", "negative", "some_zero_E", "laguerre"]


def check_accurate(out, D, E):
    from scipy.linalg import eigh_tridiagonal
    n = len(D)
    w = eigh_tridiagonal(D, E, eigvals_only=True)
    nT = np.abs(D).max() + 2 * np.abs(E).max()
    V = out["V"]
    assert np.abs(out["lam"] - w).max() <= 5e-14 * nT
    assert out["resid"].max() <= 5e-14 * nT
    assert np.abs(V.T @ V - np.eye(n)).max() <= 5e-13
