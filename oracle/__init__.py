"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the Cuppen hot path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package, and only as the checker.
The product (``symmetric_eigenvalue_b200``) never imports it.
"""
from .oracle import (  # noqa: F401
    build, load, solve, merge, scheme, run_reference, ref_binary, write_mtx,
    rand_u, goe, wilkinson,
)
