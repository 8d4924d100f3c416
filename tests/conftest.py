import ctypes
import glob
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    import oracle as orc
    orc.build(ref=os.path.isdir("/root/reference/src"))
    return orc


@pytest.fixture(scope="session")
def hostemu():
    """TEST-ONLY host build of the product sources (tests/host/Makefile)."""
    from symmetric_eigenvalue_b200 import api
    subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "host")], check=True, stdout=subprocess.DEVNULL)
    return api._declare(ctypes.CDLL(os.path.join(ROOT, "tests", "host", "_build", "libcuppen_hostemu.so")))


@pytest.fixture(scope="session")
def product_lib():
    """The shipped CUDA library; built in-tree by `make lib` (nvcc cross-compiles without a GPU)."""
    from symmetric_eigenvalue_b200 import api
    if not os.path.exists(api.library_path()):
        subprocess.run(["make", "-C", ROOT, "lib", "cuppen"], check=True, stdout=subprocess.DEVNULL)
    return api.load_library()


def golden_cases():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz")))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return dict(D=z["D"], E=z["E"], P=int(z["P"]), lam=z["lam"], resid=z["resid"], merges=z["merges"], rhos=z["rhos"],
                sel=(z["sel"] if "sel" in z.files else None))


def check_select_against_efile_golden(g, out):
    """Selected-eigenvector mode against the reference's own `-eFILE` run (golden `*_sel`): eigenvalues within
    1e-12*||T||, identical deflation counts, residuals of the selected vectors at or below the reference's."""
    check_against_golden(g, out, False)
    nT = norm_T(g["D"], g["E"])
    ref = g["resid"][g["sel"] - 1]
    assert np.isfinite(ref).all()
    assert (out["resid"] <= ref.max() * 1.05 + 4 * 2.2e-16 * nT).all(), (out["resid"], ref)


def norm_T(D, E):
    return float(np.abs(D).max() + (2 * np.abs(E).max() if len(E) else 0.0))


def ref_stats(stats):
    """(m, offset, zdefl, givens) of the merges that exist in the reference tree, reference order."""
    return sorted((s.m, s.offset, s.zdefl, s.givens) for s in stats if s.mode == 1)


def check_against_golden(g, out, vectors):
    """The parity bar of BASELINE.json: eigenvalues within 1e-12*||T||, identical deflation counts per
    merge, residuals at or below the reference's (where the reference printed them)."""
    nT = norm_T(g["D"], g["E"])
    tol = 1e-12 * nT
    assert np.abs(out["lam"] - g["lam"]).max() <= tol, (np.abs(out["lam"] - g["lam"]).max(), tol)
    assert ref_stats(out["stats"]) == [tuple(int(x) for x in r) for r in g["merges"].tolist()]
    if vectors and np.isfinite(g["resid"]).any():
        ok = np.isfinite(g["resid"])
        # "at or below the reference's": the reference's own residual column, with 4 ulp*||T|| of slack
        assert (out["resid"][ok] <= g["resid"][ok].max() * 1.05 + 4 * 2.2e-16 * nT).all()
