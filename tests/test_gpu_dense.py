"""Dense symmetric front end: Householder tridiagonalisation + the tridiagonal path +
back-transformation, through the C ABI (cuppen_dense_eigh), against LAPACK (numpy.linalg.eigh): eigenvalues to
c*n*eps*||A||, residual ||A Z - Z W|| and orthogonality ||Z^T Z - I|| at working precision."""
import numpy as np
import pytest

import symmetric_eigenvalue_b200 as se
from symmetric_eigenvalue_b200 import api

pytestmark = pytest.mark.gpu


def _matrix(kind, n, seed=3):
    rng = np.random.default_rng(seed)
    if kind == "random":
        B = rng.normal(size=(n, n))
        return (B + B.T) / 2
    if kind == "diag":
        return np.diag(rng.normal(size=n))
    if kind == "identity":
        return np.eye(n)
    if kind == "lowrank":
        U = rng.normal(size=(n, 3))
        return U @ U.T
    if kind == "clustered":
        Q, _ = np.linalg.qr(rng.normal(size=(n, n)))
        w = np.concatenate([np.full(n // 2, 1.0), 1.0 + 1e-9 * rng.normal(size=n - n // 2)])
        return (Q * w) @ Q.T
    if kind == "laplace2d":
        m = int(round(np.sqrt(n)))
        T = 2 * np.eye(m) - np.eye(m, k=1) - np.eye(m, k=-1)
        return np.kron(np.eye(m), T) + np.kron(T, np.eye(m))
    raise KeyError(kind)


@pytest.mark.parametrize("kind,n", [("random", 1), ("random", 2), ("random", 3), ("random", 64), ("random", 65), ("random", 200),
                                    ("random", 513), ("random", 1500), ("diag", 100), ("identity", 70), ("lowrank", 300),
                                    ("clustered", 257), ("laplace2d", 1024)])
def test_dense_eigh_against_lapack(product_lib, kind, n):
    A = _matrix(kind, n)
    n = A.shape[0]
    w, Z, t = api.dense_eigh(A, lib=product_lib)
    w_ref = np.linalg.eigvalsh(A)
    nA = max(np.abs(A).sum(axis=0).max(), 1e-300)
    assert (np.diff(w) >= 0).all()
    assert np.abs(w - w_ref).max() <= 20 * n * 2.2e-16 * nA, np.abs(w - w_ref).max()
    R = A @ Z - Z * w[None, :]
    assert np.abs(R).max() <= 50 * n * 2.2e-16 * nA, np.abs(R).max()
    assert np.abs(Z.T @ Z - np.eye(n)).max() <= 50 * n * 2.2e-16
    # eigenvalues only: the same numbers without a back-transformation
    w2, Z2, _ = api.dense_eigh(A, vectors=False, lib=product_lib)
    assert Z2 is None and np.abs(w2 - w).max() <= 20 * n * 2.2e-16 * nA


def test_dense_eigh_reads_the_lower_triangle_and_rejects_bad_input(product_lib):
    A = _matrix("random", 130)
    L = np.tril(A) + np.triu(np.full_like(A, 7.0), 1)          # garbage above the diagonal
    w, Z, _ = api.dense_eigh(L, lib=product_lib)
    assert np.abs(w - np.linalg.eigvalsh(A)).max() <= 1e-11
    B = A.copy(); B[5, 3] = np.nan
    with pytest.raises(se.CuppenError) as ei:
        api.dense_eigh(B, lib=product_lib)
    assert ei.value.code == -1
